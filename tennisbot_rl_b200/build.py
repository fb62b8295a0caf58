"""Compile libtennisbot_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tennisbot_rl_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_PATH = PKG / "libtennisbot_b200.so"
SOURCES = [CSRC / "tb_kernels.cu"]
HEADERS = [CSRC / "tb_device.cuh", CSRC / "tb_scene_data.h", ROOT / "include" / "tennisbot_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale():
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, SOURCES)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (PKG / "build.log").write_text(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed, see %s" % (PKG / "build.log"))
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
