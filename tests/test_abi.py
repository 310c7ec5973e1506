"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/nanogicp_c.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "nanogicp_c.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ngicp_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from direct_lidar_odometry_b200 import _lib
    L = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/nanogicp_c.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert b"sm_100a" in L.ngicp_version()


def test_struct_layouts_match_header():
    from direct_lidar_odometry_b200 import _lib
    L = _lib.load()
    p = _lib.Params()
    L.ngicp_params_default(C.byref(p))
    # reference defaults: nano_gicp_impl.hpp:57-61, lsq_registration_impl.hpp:52-59
    assert p.k_correspondences == 20 and p.max_iterations == 64 and p.lm_max_iterations == 10
    assert abs(p.transformation_epsilon - 5e-4) < 1e-18 and abs(p.rotation_epsilon - 2e-3) < 1e-18
    assert p.lm_init_lambda_factor == 1e-9 and p.regularization_method == _lib.REG_PLANE
    assert p.optimizer == _lib.OPT_LEVENBERG_MARQUARDT
    assert p.max_correspondence_distance == float(C.c_float(3.4028234663852886e38).value)
    assert C.sizeof(_lib.Result) == 16 * 4 + 16 * 8 + 36 * 8 + 2 * 8 + 6 * 4
    from oracle import oracle
    assert C.sizeof(oracle.AlignResult) == C.sizeof(_lib.Result)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from direct_lidar_odometry_b200 import NanoGICP, NanoGICPError
    with pytest.raises(NanoGICPError):
        NanoGICP(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "direct_lidar_odometry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"liboracle", r"\borc_[a-z]", r"oracle/", r"oracle\."):
                    assert not re.search(pat, txt), f"{f} references the oracle ({pat})"
