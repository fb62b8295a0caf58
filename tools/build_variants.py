"""Build kernel variants for A/B runs on the GPU box: build/variants/lib_<name>.so (git-ignored, travels with gpurun).
usage: build_variants.py name[:-DFLAG[,-DFLAG...]] ...      e.g.  base diag:-DTB_FF_DIAG stride4:-DTB_FF_SERVER_STRIDE=4
Load one with TB_LIB_PATH=build/variants/lib_<name>.so (tools/ab_variants.sh)."""
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tennisbot_rl_b200 import build as B

OUT = B.ROOT / "build" / "variants"


def one(spec):
    name, _, flags = spec.partition(":")
    out = OUT / f"lib_{name}.so"
    cmd = [B.find_nvcc(), *[f for f in B.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *(flags.split(",") if flags else []), "-o", str(out),
           *map(str, B.SOURCES)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, (r.stdout + r.stderr)[-2000:]


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for name, rc, log in ex.map(one, sys.argv[1:]):
            print(name, "ok" if rc == 0 else "FAILED\n" + log)
