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
