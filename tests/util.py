"""Helpers shared by the CPU and GPU parity tests."""
import numpy as np


def tie_free_mask(d2_kplus1: np.ndarray) -> np.ndarray:
    """Given ascending squared distances of the k+1 nearest neighbours (n, k+1), return an (n, k) mask of
    the slots whose distance differs from both adjacent slots.  The parity rule (BASELINE.json north_star,
    SURVEY A12) compares neighbour INDICES only there; distances are always compared bit-for-bit."""
    d = d2_kplus1
    k = d.shape[1] - 1
    m = np.ones((d.shape[0], k), dtype=bool)
    m[:, 1:] &= d[:, 1:k] != d[:, 0:k - 1]
    m &= d[:, 0:k] != d[:, 1:k + 1]
    return m


def rot_angle(R: np.ndarray) -> float:
    c = (np.trace(R[:3, :3]) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def pose_delta(Ta: np.ndarray, Tb: np.ndarray):
    """(translation distance [m], rotation angle [rad]) between two 4x4 poses."""
    Ta = np.asarray(Ta, dtype=np.float64)
    Tb = np.asarray(Tb, dtype=np.float64)
    dt = float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    # small-angle robust: use the skew part
    s = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return dt, float(np.arcsin(min(1.0, np.linalg.norm(s))))


def make_small_submap(O):
    """A scaled-down S2M case: 6 keyframes 5 m apart voxelised at 0.5 m in the world frame (reference
    odom.cc:484-490), and one scan voxelised at 0.25 m as the source.  `O` = the oracle module."""
    from direct_lidar_odometry_b200 import synth
    keys = []
    for j in range(6):
        i = j * 33
        T = synth.trajectory_pose(i)
        s = synth.crop_box_negative(synth.os1_like(i, T))
        w = synth.transform_xyzi(O.voxel_filter(s, 0.25), T.astype(np.float32))
        keys.append(O.voxel_filter(w, 0.5))
    submap = np.ascontiguousarray(np.vstack(keys))
    i = 90
    T = synth.trajectory_pose(i)
    scan = O.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T)), 0.25)
    return submap, scan, T
