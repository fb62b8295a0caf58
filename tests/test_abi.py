"""The C-ABI library loads and exports every symbol include/tennisbot_b200.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tennisbot_rl_b200 import _lib, build

    build.build_library()
    return _lib.load()


def header_symbols():
    text = (ROOT / "include" / "tennisbot_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from tennisbot_rl_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.EXPORTS) == syms  # the Python binding lists exactly the header's entry points
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.tb_abi_version() == 1


def test_dims_params_and_scene_constants(lib, oracle_lib):
    from tennisbot_rl_b200 import _lib

    assert (lib.tb_obs_dim(0), lib.tb_act_dim(0), lib.tb_obs_dim(1), lib.tb_act_dim(1)) == (6, 6, 12, 2)
    # the oracle and the CUDA library expose the same named parameters ...
    assert _lib.param_names() == oracle_lib.OracleEnv.param_names()
    # ... and hold the same scene constants (they are compiled from two separately generated headers)
    for name in ("urdf_margin", "ball_radius", "ball_mass", "racket_mass", "racket_com_z", "racket_half_x", "floor_hx",
                 "floor_hy", "floor_hz", "net_hx", "net_hy", "net_hz", "goal_radius", "goal_half_z", "goal_sides",
                 "racket_outline_n", "contact_threshold"):
        assert _lib.scene_constant(name) == oracle_lib.scene_constant(name), name
    for i in range(38):
        assert _lib.scene_constant("racket_outline_y", i) == oracle_lib.scene_constant("racket_outline_y", i)
        assert _lib.scene_constant("racket_outline_z", i) == oracle_lib.scene_constant("racket_outline_z", i)
    for i in range(32):
        assert _lib.scene_constant("goal_vertex_x", i) == oracle_lib.scene_constant("goal_vertex_x", i)
    for i in range(3):
        assert _lib.scene_constant("racket_inertia", i) == oracle_lib.scene_constant("racket_inertia", i)
    with pytest.raises(_lib.TennisbotLibraryError):
        _lib.scene_constant("no_such_constant")


def test_bad_arguments_fail_with_messages(lib):
    from tennisbot_rl_b200 import _lib

    h = ctypes.c_void_p()
    cfg = _lib.TbConfig(0, 0, 1, 0, 16, 0, 0, 1, 0)  # wrong struct_size
    assert lib.tb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"struct_size" in lib.tb_last_error()
    cfg = _lib.TbConfig(ctypes.sizeof(_lib.TbConfig), 7, 1, 0, 16, 0, 0, 1, 0)
    assert lib.tb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0 and b"env_kind" in lib.tb_last_error()
    cfg = _lib.TbConfig(ctypes.sizeof(_lib.TbConfig), 0, 1, 0, 0, 0, 0, 1, 0)
    assert lib.tb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0 and b"num_envs" in lib.tb_last_error()


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path refuses to run instead of computing on the CPU."""
    import torch

    from tennisbot_rl_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    cfg = _lib.TbConfig(ctypes.sizeof(_lib.TbConfig), 0, 1, 0, 16, 0, 0, 1, 0)
    assert lib.tb_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.tb_last_error() or b"CPU fallback" in lib.tb_last_error()
    from tennisbot_rl_b200.batch import TennisBatch

    with pytest.raises(_lib.TennisbotLibraryError):
        TennisBatch("SwingRacket-v0", 4)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product packages may import it."""
    files = list((ROOT / "tennisbot_rl_b200").rglob("*.py")) + list((ROOT / "tennisbot").rglob("*.py"))
    assert files
    for f in files:
        for line in f.read_text().splitlines():
            assert not re.match(r"\s*(from|import)\s+oracle\b", line), (f, line)
            assert "tb_oracle" not in line and "libtb_oracle" not in line, (f, line)
