"""The C-ABI library loads on a CPU-only box and exports every symbol include/henbun_b200.h declares
(no compute calls are made here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "henbun_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from henbun_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/henbun_b200.h but not exported"
    # and the ctypes table covers the same set, so no entry point is bound with default (int) signatures
    assert sorted(_lib.SIGNATURES) == [n for n in names if n != "hb_gp_config"]


def test_pure_host_entry_points():
    from henbun_b200 import _lib
    lib = _lib.load()
    assert lib.hb_version() >= 100
    assert lib.hb_reduce_workspace_bytes() > 0
    assert lib.hb_potrf_workspace_bytes(1000) >= 8 * 128 * 128 * 4
    cfg = _lib.GpConfig(65536, 8, 64, 1, 0, 1e-5, 0, 0)
    assert lib.hb_gp_param_count(ctypes.byref(cfg)) == 2 * 65536 + 4
    assert lib.hb_gp_elbo_workspace_bytes(ctypes.byref(cfg)) > 2 * 65536 * 65536 * 4
    assert lib.hb_set_gemm_engine(7) == 1 and lib.hb_set_gemm_engine(0) == 0       # HB_ERR_ARG on bad mode
