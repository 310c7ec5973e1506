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


def test_imu_prior_matches_oracle_restatement():
    """ngicp_imu_prior (OdomNode::integrateIMU, odom.cc:859-919; host arithmetic, no GPU) against the oracle's numpy
    restatement: bit-identical float matrices, samples outside the frame interval ignored, unsorted input sorted,
    degenerate inputs give the identity; and the result is the rotation the gyro actually describes."""
    import numpy as np
    from direct_lidar_odometry_b200 import imu_prior
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    t0, t1 = 100.0, 100.1
    stamps = np.sort(rng.uniform(99.95, 100.15, size=60))
    av = rng.normal(0.0, 0.4, size=(60, 3)) + np.array([0.0, 0.0, 0.8])
    perm = rng.permutation(60)
    for s, a in ((stamps, av), (stamps[perm], av[perm])):
        got = imu_prior(s, a, t0, t1)
        ref = O.integrate_imu(s, a, t0, t1)
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(got[3], [0, 0, 0, 1]) and np.array_equal(got[:3, 3], [0, 0, 0])
        assert np.allclose(got[:3, :3] @ got[:3, :3].T, np.eye(3), atol=1e-6)
    # constant yaw rate: rotation about z by rate * (span of the samples inside the frame)
    s = np.linspace(t0, t1, 41)
    a = np.tile([0.0, 0.0, 0.5], (41, 1))
    T = imu_prior(s, a, t0, t1)
    assert abs(np.arctan2(T[1, 0], T[0, 0]) - 0.5 * 0.1) < 1e-5
    # nothing usable: identity
    assert np.array_equal(imu_prior(np.zeros(0), np.zeros((0, 3)), t0, t1), np.eye(4, dtype=np.float32))
    assert np.array_equal(imu_prior(np.array([100.05]), np.ones((1, 3)), t0, t1), np.eye(4, dtype=np.float32))
    assert np.array_equal(imu_prior(stamps, av, 200.0, 200.1), np.eye(4, dtype=np.float32))
