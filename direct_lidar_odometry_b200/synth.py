"""Synthetic OS1-64-like LiDAR scans for parity tests and the bench.

There is no dataset in the reference tree (the README points at a 4.2 GB rosbag,
reference README.md:61-72) and no network here, so every test/bench input is made
by this deterministic ray-caster: a spinning LiDAR (beams x cols rays, +-22.5 deg
vertical field of view) inside a procedural street scene (ground plane z=0 plus
axis-aligned "buildings" on a 40 m lattice).  Points are emitted in the SENSOR
frame as pcl::PointXYZI records (reference include/dlo/dlo.h:50): 8 float32 per
point {x, y, z, 1.0, intensity, 0, 0, 0} = 32 bytes, row-major beam x col order.

Pure numpy, no GPU.  Used by tests/, bench.py and __graft_entry__.smoke().
"""
from __future__ import annotations

import numpy as np

SCENE_SEED = 7
LATTICE = 40.0
SENSOR_HEIGHT = 1.8

# trajectory: rounded rectangle 120 x 80 m whose edges run along lattice "streets"
LOOP_W, LOOP_H, LOOP_R = 120.0, 80.0, 8.0
SCAN_SPACING = 0.15  # 1.5 m/s at 10 Hz


def make_scene(seed: int = SCENE_SEED, lo: float = -240.0, hi: float = 360.0) -> np.ndarray:
    """Boxes as rows [xmin, ymin, zmin, xmax, ymax, zmax] (world frame)."""
    rng = np.random.default_rng(seed)
    centers = np.arange(lo + LATTICE / 2, hi, LATTICE)
    boxes = []
    for cx in centers:
        for cy in centers:
            fx, fy = rng.uniform(8.0, 30.0, size=2)
            h = rng.uniform(5.0, 25.0)
            jx, jy = rng.uniform(-2.0, 2.0, size=2)
            boxes.append([cx + jx - fx / 2, cy + jy - fy / 2, 0.0,
                          cx + jx + fx / 2, cy + jy + fy / 2, h])
    return np.asarray(boxes, dtype=np.float64)


def _loop_point(s: float):
    """Position (x, y) and heading of the rounded-rectangle loop at arc length s."""
    w, h, r = LOOP_W, LOOP_H, LOOP_R
    segs = [w - 2 * r, np.pi * r / 2, h - 2 * r, np.pi * r / 2] * 2
    total = sum(segs)
    s = s % total
    # straight segments start points / headings, arcs centres
    # order: bottom edge (+x), corner BR, right edge (+y), corner TR, top edge (-x), corner TL, left edge (-y), corner BL
    starts = [(r, 0.0, 0.0), None, (w, r, np.pi / 2), None, (w - r, h, np.pi), None, (0.0, h - r, -np.pi / 2), None]
    arcs = [None, (w - r, r, -np.pi / 2), None, (w - r, h - r, 0.0), None, (r, h - r, np.pi / 2), None, (r, r, np.pi)]
    for k, L in enumerate(segs):
        if s <= L:
            if starts[k] is not None:
                x0, y0, th = starts[k]
                return x0 + s * np.cos(th), y0 + s * np.sin(th), th
            cx, cy, a0 = arcs[k]
            a = a0 + s / r
            return cx + r * np.cos(a), cy + r * np.sin(a), a + np.pi / 2
        s -= L
    raise AssertionError


def trajectory_pose(i: int, spacing: float = SCAN_SPACING, start: float = 20.0) -> np.ndarray:
    """4x4 float64 world-from-sensor pose of scan i."""
    s = start + spacing * i
    x, y, yaw = _loop_point(s)
    z = SENSOR_HEIGHT + 0.05 * np.sin(2 * np.pi * s / 20.0)
    c, sn = np.cos(yaw), np.sin(yaw)
    T = np.eye(4)
    T[:3, :3] = [[c, -sn, 0.0], [sn, c, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = [x, y, z]
    return T


def perturb_pose(T: np.ndarray, dt=(0.2, 0.0, 0.0), yaw_deg: float = 1.0) -> np.ndarray:
    a = np.deg2rad(yaw_deg)
    D = np.eye(4)
    D[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    D[:3, 3] = dt
    return D @ T


def os1_like(scan_idx: int, pose: np.ndarray, beams: int = 64, cols: int = 1024,
             vfov_deg: float = 22.5, range_max: float = 100.0, sigma: float = 0.02,
             drop_prob: float = 0.02, boxes: np.ndarray | None = None,
             range_min: float = 0.5) -> np.ndarray:
    """Ray-cast one scan.  Returns (n, 8) float32 PointXYZI records, sensor frame."""
    if boxes is None:
        boxes = _default_scene()
    az = (np.arange(cols) / cols) * 2 * np.pi
    el = np.deg2rad(np.linspace(-vfov_deg, vfov_deg, beams))
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_s = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :],
                    np.broadcast_to(se, (beams, cols))], axis=-1).reshape(-1, 3)
    R, o = pose[:3, :3], pose[:3, 3]
    d = d_s @ R.T
    n = d.shape[0]
    t_hit = np.full(n, np.inf)
    # ground plane z = 0
    down = d[:, 2] < -1e-9
    t_hit[down] = -o[2] / d[down, 2]
    # boxes within reach
    reach = range_max + 25.0
    near = ((boxes[:, 0] < o[0] + reach) & (boxes[:, 3] > o[0] - reach) &
            (boxes[:, 1] < o[1] + reach) & (boxes[:, 4] > o[1] - reach))
    bx = boxes[near]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        for b in bx:
            t1 = (b[None, :3] - o[None, :]) * inv
            t2 = (b[None, 3:] - o[None, :]) * inv
            tn = np.nanmax(np.minimum(t1, t2), axis=1)
            tf = np.nanmin(np.maximum(t1, t2), axis=1)
            ok = (tf >= tn) & (tn > 0)
            t_hit = np.where(ok & (tn < t_hit), tn, t_hit)
    rng = np.random.default_rng(1000 + scan_idx)
    noise = rng.normal(0.0, sigma, size=n)
    keep = rng.random(n) >= drop_prob
    valid = np.isfinite(t_hit) & (t_hit <= range_max) & (t_hit >= range_min) & keep
    r = (t_hit + noise)[valid]
    p = d_s[valid] * r[:, None]
    out = np.zeros((p.shape[0], 8), dtype=np.float32)
    out[:, :3] = p.astype(np.float32)
    out[:, 3] = 1.0
    out[:, 4] = (t_hit[valid] / 100.0).astype(np.float32)
    return out


def os1_like_batch_torch(scan_indices, poses, device, beams: int = 64, cols: int = 1024, vfov_deg: float = 22.5,
                         range_max: float = 100.0, sigma: float = 0.02, drop_prob: float = 0.02, range_min: float = 0.5,
                         crop: float | None = 1.0):
    """The same ray-caster as os1_like, vectorised over a batch of scans with torch (fp64) on `device` — the bulk
    generator behind the 10 000-pair C4 and the 5 M-point C5 workloads, which numpy needs minutes for.  Same scene, same
    per-scan numpy RNG stream (noise seed 1000 + scan index), same arithmetic; the results agree with os1_like up to the
    last-bit differences of cos/sin and of the 3x3 products (they are NOT guaranteed identical: fixtures made with one
    generator must be checked with the same one).  Returns a list of (n_i, 8) float32 PointXYZI tensors on `device`,
    with the negative crop box (odom.cc:122-124) applied unless crop is None."""
    import torch
    boxes_np = _default_scene()
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    B = poses.shape[0]
    f64 = torch.float64
    az = torch.arange(cols, dtype=f64, device=device) / cols * (2 * np.pi)
    el = torch.deg2rad(torch.linspace(-vfov_deg, vfov_deg, beams, dtype=f64, device=device))
    ce, se = torch.cos(el)[:, None], torch.sin(el)[:, None]
    d_s = torch.stack([ce * torch.cos(az)[None, :], ce * torch.sin(az)[None, :], se.expand(beams, cols)], dim=-1).reshape(-1, 3)
    N = d_s.shape[0]
    R = torch.from_numpy(poses[:, :3, :3].copy()).to(device)
    o = torch.from_numpy(poses[:, :3, 3].copy()).to(device)
    d = torch.einsum("nj,bij->bni", d_s, R)                       # d_s @ R.T per scan
    t_hit = torch.full((B, N), float("inf"), dtype=f64, device=device)
    down = d[..., 2] < -1e-9
    t_hit = torch.where(down, -o[:, None, 2] / d[..., 2], t_hit)
    reach = range_max + 25.0
    on = poses[:, :3, 3]
    near = np.zeros(boxes_np.shape[0], dtype=bool)
    for b in range(B):                                             # a superset of every scan's own "near" set: same hits
        near |= ((boxes_np[:, 0] < on[b, 0] + reach) & (boxes_np[:, 3] > on[b, 0] - reach) &
                 (boxes_np[:, 1] < on[b, 1] + reach) & (boxes_np[:, 4] > on[b, 1] - reach))
    bx = torch.from_numpy(boxes_np[near]).to(device)
    inv = 1.0 / d
    for j in range(bx.shape[0]):
        t1 = (bx[j, :3][None, None, :] - o[:, None, :]) * inv
        t2 = (bx[j, 3:][None, None, :] - o[:, None, :]) * inv
        lo_, hi_ = torch.minimum(t1, t2), torch.maximum(t1, t2)
        tn = torch.fmax(torch.fmax(lo_[..., 0], lo_[..., 1]), lo_[..., 2])
        tf = torch.fmin(torch.fmin(hi_[..., 0], hi_[..., 1]), hi_[..., 2])
        ok = (tf >= tn) & (tn > 0) & (tn < t_hit)
        t_hit = torch.where(ok, tn, t_hit)
    def _noise(idx):                                               # the numpy generator's stream, drawn on host threads
        rng = np.random.default_rng(1000 + int(idx))
        return rng.normal(0.0, sigma, size=N), rng.random(N) >= drop_prob
    if B > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(B, 8)) as ex:
            draws = list(ex.map(_noise, scan_indices))
    else:
        draws = [_noise(scan_indices[0])]
    out = []
    for b, idx in enumerate(scan_indices):
        noise = torch.from_numpy(draws[b][0]).to(device)
        keep = torch.from_numpy(draws[b][1]).to(device)
        th = t_hit[b]
        valid = torch.isfinite(th) & (th <= range_max) & (th >= range_min) & keep
        r = (th + noise)[valid]
        p = (d_s[valid] * r[:, None]).to(torch.float32)
        if crop is not None:
            outside = ~(p.abs() < crop).all(dim=1)
            p = p[outside]
            inten = (th[valid] / 100.0).to(torch.float32)[outside]
        else:
            inten = (th[valid] / 100.0).to(torch.float32)
        rec = torch.zeros((p.shape[0], 8), dtype=torch.float32, device=device)
        rec[:, :3] = p
        rec[:, 3] = 1.0
        rec[:, 4] = inten
        out.append(rec)
    return out


_SCENE_CACHE: dict = {}


def _default_scene() -> np.ndarray:
    if "b" not in _SCENE_CACHE:
        _SCENE_CACHE["b"] = make_scene()
    return _SCENE_CACHE["b"]


def transform_xyzi(pts: np.ndarray, T: np.ndarray) -> np.ndarray:
    """Apply a float32 4x4 to PointXYZI records the way pcl::transformPointCloud does
    for an affine matrix (float arithmetic, w untouched)."""
    Tf = np.asarray(T, dtype=np.float32)
    out = pts.copy()
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    for r in range(3):
        out[:, r] = Tf[r, 0] * x + Tf[r, 1] * y + Tf[r, 2] * z + Tf[r, 3]
    return out


def crop_box_negative(pts: np.ndarray, size: float = 1.0) -> np.ndarray:
    """pcl::CropBox with setNegative(true): drop points inside +-size (reference odom.cc:122-124)."""
    inside = np.all(np.abs(pts[:, :3]) < size, axis=1)
    return pts[~inside]


def random_planes_cloud(n: int, seed: int = 0, extent: float = 20.0, noise: float = 0.01) -> np.ndarray:
    """Small generic test cloud: three noisy orthogonal planes + clutter. (n, 8) float32."""
    rng = np.random.default_rng(seed)
    k = n // 4
    a = rng.uniform(-extent, extent, size=(k, 2))
    p1 = np.c_[a, rng.normal(0, noise, k)]
    a = rng.uniform(-extent, extent, size=(k, 2))
    p2 = np.c_[a[:, 0], rng.normal(extent * 0.3, noise, k), np.abs(a[:, 1]) * 0.5]
    a = rng.uniform(-extent, extent, size=(k, 2))
    p3 = np.c_[rng.normal(-extent * 0.4, noise, k), a[:, 0], np.abs(a[:, 1]) * 0.5]
    p4 = rng.uniform(-extent, extent, size=(n - 3 * k, 3)) * [1, 1, 0.25]
    p = np.vstack([p1, p2, p3, p4])
    rng.shuffle(p)
    out = np.zeros((n, 8), dtype=np.float32)
    out[:, :3] = p.astype(np.float32)
    out[:, 3] = 1.0
    out[:, 4] = rng.random(n).astype(np.float32)
    return out
