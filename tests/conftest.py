import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_knn():
    return np.load(os.path.join(ROOT, "tests", "golden", "knn_ref_nanoflann.npz"))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.load(prefer_ref=False)  # must at least have the restated build
    return oracle


@pytest.fixture(scope="session")
def scan_pair():
    """Two consecutive synthetic OS1-64 scans, crop-boxed (reference odom.cc:454-457); raw, not voxelised."""
    from direct_lidar_odometry_b200 import synth
    T0, T1 = synth.trajectory_pose(0), synth.trajectory_pose(1)
    s0 = synth.crop_box_negative(synth.os1_like(0, T0))
    s1 = synth.crop_box_negative(synth.os1_like(1, T1))
    return dict(T0=T0, T1=T1, s0=s0, s1=s1)
