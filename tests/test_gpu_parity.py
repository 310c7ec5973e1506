"""GPU parity tests: every stage of the CUDA path against the CPU oracle, called through the C ABI.

Tolerances are the ones BASELINE.json's north_star states:
  * voxel assignment and kNN indices bit-exact (indices compared wherever distances are not exactly tied),
    kNN squared distances bit-identical
  * covariances within 1e-5 relative
  * final poses within 1e-4 m / 1e-5 rad, identical iteration counts
"""
import numpy as np
import pytest

from direct_lidar_odometry_b200 import synth
from util import tie_free_mask, pose_delta

pytestmark = pytest.mark.gpu

COV_RTOL = 1e-5
POSE_T_TOL = 1e-4
POSE_R_TOL = 1e-5


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from direct_lidar_odometry_b200 import NanoGICP
    return NanoGICP


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    oracle.load(prefer_ref=True)
    return oracle


@pytest.fixture(scope="module")
def vox_pair(O, scan_pair):
    v0 = O.voxel_filter(scan_pair["s0"], 0.25)
    v1 = O.voxel_filter(scan_pair["s1"], 0.25)
    return v0, v1


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------------------- K0
@pytest.mark.parametrize("leaf", [0.25, 0.5, 2.0])
def test_voxel_filter_bit_exact(G, O, scan_pair, leaf):
    g = G()
    s0 = scan_pair["s0"]
    ref, ref_assign, rc = O.voxel_filter(s0, leaf, return_assignment=True)
    out, st = g.voxel_filter(s0, leaf, return_status=True)
    assert st == 0 and rc == 0
    assert out.shape == ref.shape
    assert np.array_equal(bits(out), bits(ref))                      # centroids, intensity, padding: bit-exact
    assert np.array_equal(g.voxel_assignment(s0.shape[0]), ref_assign)  # voxel of every input point
    # idempotence-style property at full size: re-filtering the centroids cannot add points
    again = g.voxel_filter(out, leaf)
    assert again.shape[0] <= out.shape[0]


def test_voxel_filter_edge_cases(G, O, scan_pair):
    g = G()
    s0 = scan_pair["s0"]
    assert g.voxel_filter(np.zeros((0, 8), np.float32), 0.25).shape[0] == 0
    one = g.voxel_filter(s0[:1], 0.25)
    assert one.shape[0] == 1 and np.array_equal(bits(one), bits(O.voxel_filter(s0[:1], 0.25)))
    # non-finite points are skipped (is_dense=false path of PCL)
    dirty = s0[:5000].copy()
    dirty[::7, 0] = np.nan
    dirty[3::11, 2] = np.inf
    ref, ref_assign, _ = O.voxel_filter(dirty, 0.5, return_assignment=True)
    out = g.voxel_filter(dirty, 0.5)
    assert np.array_equal(bits(out), bits(ref))
    assert np.array_equal(g.voxel_assignment(dirty.shape[0]), ref_assign)
    # all points in one voxel
    blob = s0[:300].copy()
    blob[:, :3] = blob[:, :3] * 1e-3 + 5.0
    assert np.array_equal(bits(g.voxel_filter(blob, 1.0)), bits(O.voxel_filter(blob, 1.0)))
    # int32 voxel index overflow: PCL warns and passes the input through
    far = s0[:100].copy()
    far[0, :3] = 1e6
    out, st = g.voxel_filter(far, 0.01, return_status=True)
    assert st == 1 and out.shape[0] == 100 and np.array_equal(out[:, :3], far[:, :3])
    # packed xyz input (stride 12) gives the same centroids with zero intensity
    xyz = np.ascontiguousarray(s0[:4000, :3])
    out = g.voxel_filter(xyz, 0.5)
    ref = O.voxel_filter(xyz, 0.5)
    assert np.array_equal(bits(out[:, :4]), bits(ref[:, :4]))



@pytest.mark.parametrize("path", [0, 1])
def test_voxel_filter_fused_and_multi_kernel_paths(G, O, scan_pair, path):
    """ngicp_params::voxel_path: 0 = one persistent cooperative launch (pack -> keys -> radix passes -> heads -> centroids
    between grid barriers) when the cloud fits 4096 points per SM, 1 = the multi-kernel pipeline.  Both must reproduce
    the oracle's bits and voxel assignment: raw scan, NaN / Inf, tiny and wide extents (1..4 radix passes), crop box,
    a cloud larger than the fused path holds (falls back by itself), PCL's overflow pass-through."""
    g = G()
    g.setVoxelPath(path)
    s0 = scan_pair["s0"]
    small = s0[:3000].copy(); small[:, :3] *= 0.01
    wide = s0.copy(); wide[::7, :3] *= 4.0
    dirty = s0[:20000].copy(); dirty[::7, 0] = np.nan; dirty[3::11, 2] = np.inf
    tiny = s0[:37].copy()
    for pts, leaf in [(s0, 0.25), (s0, 0.5), (small, 0.5), (small, 0.01), (wide, 0.25), (dirty, 0.5), (tiny, 0.25), (s0[:513], 1.0), (s0[:1], 0.25)]:
        ref, ref_assign, rc = O.voxel_filter(pts, leaf, return_assignment=True)
        out, st = g.voxel_filter(pts, leaf, return_status=True)
        assert st == rc and out.shape == ref.shape and np.array_equal(bits(out), bits(ref)), (path, pts.shape, leaf)
        if rc == 0:                                          # (PCL's overflow pass-through has no voxel assignment)
            assert np.array_equal(g.voxel_assignment(pts.shape[0]), ref_assign), (path, pts.shape, leaf)
    # the fused preprocess (NaN removal + negative crop box + voxel grid) through both paths
    ref = O.preprocess_points(dirty, 1.0, 0.25)
    out = g.preprocess(dirty, 1.0, 0.25)
    assert out.shape == ref.shape and np.array_equal(bits(out), bits(ref))
    # 700k points: more than 4096 per SM -> the multi-kernel pipeline whatever was asked for
    rng = np.random.default_rng(3)
    big = np.zeros((700_000, 8), np.float32)
    big[:, :3] = rng.uniform(-40, 40, size=(700_000, 3)).astype(np.float32)
    big[:, 2] *= 0.1
    big[:, 3] = 1.0
    big[:, 4] = rng.uniform(0, 1, size=700_000).astype(np.float32)
    ref = O.voxel_filter(big, 0.4)
    out = g.voxel_filter(big, 0.4)
    assert out.shape == ref.shape and np.array_equal(bits(out), bits(ref))
    half = big[:500_000]                                     # fits the fused path: 3379 points per block
    ref = O.voxel_filter(half, 0.4)
    out = g.voxel_filter(half, 0.4)
    assert out.shape == ref.shape and np.array_equal(bits(out), bits(ref))
    far = s0[:100].copy(); far[0, :3] = 1e6
    out, st = g.voxel_filter(far, 0.01, return_status=True)
    assert st == 1 and out.shape[0] == 100 and np.array_equal(out[:, :3], far[:, :3])
    assert np.array_equal(bits(g.voxel_filter(s0, 0.25)), bits(O.voxel_filter(s0, 0.25)))



@pytest.mark.parametrize("cell", [0.0, 0.7])
def test_index_fused_and_multi_kernel_paths_identical(G, O, scan_pair, vox_pair, cell):
    """ngicp_params::index_path: 0 = snapshot + search index in one persistent cooperative launch, 1 = the multi-kernel
    pipeline, 3 = one thread-block cluster per cloud.  Same grid (automatic or given cell edge), same sorted order, hence bit-identical kNN answers, covariances
    and registration — on a voxelised scan, on the raw scan (53k points), with NaNs, one point and an empty cloud."""
    v0, v1 = vox_pair
    s0 = scan_pair["s0"]
    dirty = s0[:6000].copy(); dirty[::13, 1] = np.nan
    res = []
    for path in (1, 0, 3):
        g = G()
        g.setIndexPath(path)
        if cell > 0:
            g.setGridCellSize(cell)
        g.setCorrespondenceRandomness(10); g.setMaxCorrespondenceDistance(1.0); g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
        out = {}
        for name, cloud in (("vox", v0), ("raw", s0), ("dirty", dirty)):
            g.setInputTarget(cloud)
            out[name + "_grid"] = g.grid_info(1)
            idx, d2 = g.knn(1, np.ascontiguousarray(cloud[:2000, :3]), 10)
            out[name + "_knn"] = (idx.copy(), d2.copy())
            g.calculateTargetCovariances()
            out[name + "_cov"] = g.getTargetCovariances().copy()
        g.setInputTarget(v0); g.setInputSource(v1)
        g.align()
        out["final"] = g.final_state().copy()
        out["iters"] = g.result.nr_iterations
        g.setInputTarget(s0[:1]); g.setInputSource(np.zeros((0, 8), np.float32))
        res.append(out)
    a = res[0]
    for b in res[1:]:
      for key in a:
          if key.endswith("_grid"):
              assert a[key] == b[key], key
          elif key.endswith("_knn"):
              assert np.array_equal(a[key][0], b[key][0]) and np.array_equal(bits(a[key][1]), bits(b[key][1])), key
          elif key == "iters":
              assert a[key] == b[key]
          else:
              assert np.array_equal(np.ascontiguousarray(a[key]).view(np.uint64), np.ascontiguousarray(b[key]).view(np.uint64)), key


@pytest.mark.parametrize("far", [1.0e13, 1.0e20, 3.0e38])
def test_index_survives_far_outliers(G, O, scan_pair, far):
    """DLO only strips NaN / Inf: one stray finite return of 1e20 m is a legal input.  The cell edge must grow until the
    dense table holds the grid (down to a single cell when the extent overflows float arithmetic) — never a table
    overrun — and the search must stay exact.  All three index paths."""
    pts = scan_pair["s0"][:3000].copy()
    pts[17, :3] = (far, -far * 0.5, far * 0.25)
    q = np.ascontiguousarray(pts[100:164, :3])                 # (not the outlier itself: its distances overflow float)
    ref_idx, ref_d2 = O.Cloud(pts).knn(q, 5)
    for path in (0, 1, 3):
        g = G()
        g.setIndexPath(path)
        g.setInputTarget(pts)
        info = g.grid_info(1)
        assert info["ncells"] >= 1 and info["ncells"] <= (1 << 25), info
        idx, d2 = g.knn(1, q, 5)
        fin = np.isfinite(ref_d2) & np.isfinite(d2)
        assert np.array_equal(bits(d2[fin]), bits(ref_d2[fin])), (path, far)
        assert np.array_equal(np.isfinite(d2), np.isfinite(ref_d2))
        g.setCorrespondenceRandomness(5)
        g.calculateTargetCovariances()
        assert np.isfinite(g.getTargetCovariances()[np.arange(3000) != 17]).all()


def test_degenerate_clouds_through_fused_paths(G, O, scan_pair):
    """Nothing finite, everything cropped away, one point, two identical points: the persistent voxel and index kernels
    (and their multi-kernel twins) must return the oracle's answer instead of hanging on a barrier or indexing garbage."""
    s0 = scan_pair["s0"]
    nan_cloud = s0[:2000].copy(); nan_cloud[:, 0] = np.nan
    near = s0[:2000].copy(); near[:, :3] = near[:, :3] / np.abs(near[:, :3]).max() * 0.5      # everything inside the +-1 m crop box
    twin = np.repeat(s0[:1], 2, axis=0)
    for vpath in (0, 1):
        g = G()
        g.setVoxelPath(vpath)
        assert g.voxel_filter(nan_cloud, 0.25).shape[0] == 0
        assert g.preprocess(nan_cloud, 1.0, 0.25).shape[0] == 0
        assert g.preprocess(near, 1.0, 0.25).shape[0] == 0                 # all cropped
        assert np.array_equal(bits(g.preprocess(near, None, 0.25)), bits(O.preprocess_points(near, None, 0.25)))
        assert np.array_equal(bits(g.voxel_filter(twin, 0.25)), bits(O.voxel_filter(twin, 0.25)))
        assert np.array_equal(bits(g.voxel_filter(s0, 0.25)), bits(O.voxel_filter(s0, 0.25)))     # and the handle still works
    for ipath in (0, 1, 3):
        g = G()
        g.setIndexPath(ipath)
        g.setCorrespondenceRandomness(2)
        g.setInputTarget(nan_cloud)                                         # nothing finite: a one-cell grid, no neighbours
        idx, d2 = g.knn(1, np.ascontiguousarray(s0[:4, :3]), 2)
        assert (idx == -1).all()
        g.setInputTarget(twin)
        idx, d2 = g.knn(1, np.ascontiguousarray(twin[:, :3]), 2)
        assert (d2 == 0).all() and set(idx[0].tolist()) == {0, 1}
        g.calculateTargetCovariances()
        assert np.isfinite(g.getTargetCovariances()).all()
        g.setInputTarget(s0[:3000]); g.setInputSource(s0[:3000])
        g.setCorrespondenceRandomness(10)
        g.align()
        assert np.allclose(g.final_state(), np.eye(4), atol=1e-9)           # a cloud against itself: identity


def test_preprocess_fused_crop_nan_voxel(G, O, scan_pair):
    """ngicp_preprocess = removeNaN + negative CropBox + voxel grid in one pass (odom.cc:443-465), bit-exact against
    the three steps done one after the other by the oracle."""
    g = G()
    raw = scan_pair["raw0"] if "raw0" in scan_pair else scan_pair["s0"]
    dirty = raw.copy()
    dirty[::13, 1] = np.nan
    dirty[5::17, 0] = -np.inf
    dirty[:50, :3] *= 0.01                      # a few points well inside the +-1 m box
    dirty[50, :3] = (1.0, -1.0, 1.0)            # exactly on the box: PCL counts it as inside
    for crop, leaf in ((1.0, 0.25), (1.0, 0.5), (None, 0.25), (2.5, 0.0), (None, 0.0)):
        ref = O.preprocess_points(dirty, crop, leaf)
        out = g.preprocess(dirty, crop, leaf)
        assert out.shape == ref.shape, (crop, leaf, out.shape, ref.shape)
        assert np.array_equal(bits(out), bits(ref)), (crop, leaf)
    assert g.preprocess(np.zeros((0, 8), np.float32), 1.0, 0.25).shape[0] == 0
    # everything cropped away
    assert g.preprocess(dirty[:50], 1.0, 0.25).shape[0] == 0
    # index overflow after the crop: the surviving points pass through, as PCL's VoxelGrid does with its input
    far = dirty[:200].copy()
    far[60, :3] = 1e6
    out, st = g.preprocess(far, 1.0, 0.01, return_status=True)
    ref = O.preprocess_points(far, 1.0, 0.0)
    assert st == 1 and np.array_equal(bits(out), bits(ref))


def test_voxel_filter_speculative_passes_and_fallback(G, O, scan_pair):
    """The voxel pipeline queues its radix passes with the key width the PREVIOUS call needed and checks afterwards;
    a cloud that needs more bits (or overflows PCL's index range) must fall back to the exact path, and device
    output buffers (copied by a kernel that reads the count on the device) must match host ones."""
    import torch
    g = G()
    s0 = scan_pair["s0"]
    small = s0[:3000].copy()
    small[:, :3] = small[:, :3] * 0.01                      # a few centimetres across: very few key bits
    wide = s0.copy()
    wide[::7, :3] *= 4.0                                     # four times the extent: one more radix pass
    far = s0[:200].copy()
    far[0, :3] = 1e6                                         # index overflow at leaf 0.01
    seq = [(small, 0.5), (wide, 0.25), (small, 0.01), (wide, 0.25), (s0, 0.25), (s0, 0.25), (small, 0.5)]
    for pts, leaf in seq:
        ref = O.voxel_filter(pts, leaf)
        out = g.voxel_filter(pts, leaf)
        assert out.shape == ref.shape and np.array_equal(bits(out), bits(ref)), (pts.shape, leaf)
        buf = torch.zeros((pts.shape[0], 8), dtype=torch.float32, device="cuda")
        outd = g.voxel_filter(pts, leaf, out=buf)
        assert np.array_equal(bits(outd.cpu().numpy()), bits(ref)), ("device out", pts.shape, leaf)
    out, st = g.voxel_filter(far, 0.01, return_status=True)   # after a hinted call: overflow must still pass through
    assert st == 1 and out.shape[0] == far.shape[0]
    out = g.voxel_filter(s0, 0.25)                           # and the next ordinary call is exact again
    assert np.array_equal(bits(out), bits(O.voxel_filter(s0, 0.25)))
    # device output buffer too small: an error, not a silent truncation
    tiny = torch.zeros((8, 8), dtype=torch.float32, device="cuda")
    with pytest.raises(Exception):
        g.voxel_filter(s0, 0.25, out=tiny)


@pytest.mark.parametrize("path", [1, 2])
def test_covariances_in_parts_sum_to_full(G, path):
    """ngicp_calc_source_covs_part: the slices of all parts (zeros elsewhere) add up to the full result bit for bit —
    what the all-reduce over the ranks of a sharded registration relies on; both kNN kernel families (ngicp_params
    knn_path: 1 = one warp per query, 2 = tiles).  The covariance kernel sums every neighbourhood in ascending
    (distance, index) order, so the bits do not depend on how the cloud is sliced or which kernel found the neighbours."""
    from direct_lidar_odometry_b200 import _lib
    T = synth.trajectory_pose(0)
    g = G()
    g.setKnnPath(path)
    v = g.voxel_filter(synth.crop_box_negative(synth.os1_like(0, T, cols=512)), 0.25)
    g.setCorrespondenceRandomness(20); g.setInputSource(v); g.calculateSourceCovariances(); g.sync()
    full = g.covs_device_tensor(_lib.SOURCE).cpu().numpy().copy()
    acc = np.zeros_like(full)
    nz = 0
    for p in range(3):
        g.calculateSourceCovariancesPart(p, 3); g.sync()   # the view is read on torch's stream, not the handle's
        part = g.covs_device_tensor(_lib.SOURCE).cpu().numpy()
        nz += int((np.abs(part).sum(axis=1) > 0).sum())
        acc += part
    assert nz == v.shape[0]
    assert np.array_equal(acc.view(np.uint64), full.view(np.uint64)), float(np.abs(acc - full).max())
    assert np.allclose(g.getSourceCovariances()[:, :3, :3].reshape(-1, 9)[:, [0, 1, 2, 4, 5, 8]], part)


def test_preprocess_pointcloud2_decode_fused(G, O, scan_pair):
    """ngicp_preprocess_pointcloud2 = pcl::fromROSMsg + preprocessPoints on the message bytes (odom.cc:636-637,
    443-465): bit-exact against the oracle's decode followed by its three preprocessing steps, for Ouster-like (48 B,
    organised 64 x W), Velodyne-like (22 B, unaligned, padded rows) and xyz-only messages."""
    from direct_lidar_odometry_b200 import pointcloud2 as pc2
    g = G()
    raw = scan_pair["raw0"] if "raw0" in scan_pair else scan_pair["s0"]
    n = (raw.shape[0] // 64) * 64
    dirty = raw[:n].copy()
    dirty[::13, 1] = np.nan
    dirty[5::17, 0] = -np.inf
    dirty[:50, :3] *= 0.01
    for kind, height, pad in (("ouster", 64, 0), ("ouster", 1, 0), ("velodyne", 64, 6), ("velodyne", 1, 0), ("xyz", 64, 2)):
        msg = pc2.make_pointcloud2(dirty, kind, height=height, row_pad=pad)
        decoded = O.from_ros_msg(msg)
        for crop, leaf in ((1.0, 0.25), (None, 0.5), (1.0, 0.0)):
            ref = O.preprocess_points(decoded, crop, leaf)
            out = g.preprocess_pointcloud2(msg, crop, leaf)
            assert out.shape == ref.shape, (kind, crop, leaf, out.shape, ref.shape)
            assert np.array_equal(bits(out), bits(ref)), (kind, crop, leaf)
    # the same through the plain-record entry point must agree as well
    msg = pc2.make_pointcloud2(dirty, "ouster", height=64)
    assert np.array_equal(bits(g.preprocess_pointcloud2(msg, 1.0, 0.25)), bits(g.preprocess(dirty, 1.0, 0.25)))
    # empty message, malformed layouts
    empty = pc2.make_pointcloud2(dirty[:0], "ouster")
    assert g.preprocess_pointcloud2(empty, 1.0, 0.25).shape[0] == 0
    bad = pc2.make_pointcloud2(dirty, "ouster", height=64)
    bad.fields[0] = pc2.PointField("x", 46, pc2.FLOAT32)      # runs past point_step
    with pytest.raises(Exception):
        g.preprocess_pointcloud2(bad, 1.0, 0.25)
    nox = pc2.make_pointcloud2(dirty, "ouster")
    nox.fields[0] = pc2.PointField("x", 0, pc2.FLOAT64)       # no FLOAT32 x
    with pytest.raises(Exception):
        g.preprocess_pointcloud2(nox, 1.0, 0.25)


def test_transform_voxel_filter_fused(G, O, scan_pair):
    """transformPointCloud + vf_submap in one pass (keyframes, odom.cc:484-490) == the two steps of the oracle."""
    g = G()
    s0 = scan_pair["s0"]
    T = synth.trajectory_pose(57).astype(np.float32)
    for leaf in (0.5, 0.25):
        ref = O.voxel_filter(synth.transform_xyzi(s0, T), leaf)
        out = g.transform_voxel_filter(s0, T, leaf)
        assert out.shape == ref.shape and np.array_equal(bits(out), bits(ref))
    moved = g.transform_voxel_filter(s0[:1000], T, 0.0)
    ref = synth.transform_xyzi(s0[:1000], T)
    assert np.array_equal(bits(moved[:, :3]), bits(ref[:, :3])) and np.array_equal(bits(moved[:, 4]), bits(ref[:, 4]))


# ---------------------------------------------------------------------------------------------- K1+K2
@pytest.mark.parametrize("k", [1, 5, 10, 20])
def test_knn_matches_reference_nanoflann_golden(G, golden_knn, k):
    """Vectors produced by the reference's own vendored nanoflann (tests/golden/make_golden.py)."""
    g = G()
    cloud = golden_knn["cloud"]
    g.setInputTarget(cloud)
    idx, d2 = g.knn(1, golden_knn["queries"], k)
    assert np.array_equal(bits(d2), bits(golden_knn[f"d2_k{k}"]))
    # ties: compare indices only where the distance differs from both neighbours in the (k+1)-list
    _, d2p = g.knn(1, golden_knn["queries"], k + 1)
    m = tie_free_mask(d2p)
    assert m.mean() > 0.9
    assert np.array_equal(idx[m], golden_knn[f"idx_k{k}"][m])
    # tied slots must still hold points at exactly that distance
    q = golden_knn["queries"]
    p = cloud[idx]
    dx, dy, dz = q[:, None, 0] - p[..., 0], q[:, None, 1] - p[..., 1], q[:, None, 2] - p[..., 2]
    recomputed = (dx * dx + dy * dy).astype(np.float32) + (dz * dz).astype(np.float32)
    assert np.array_equal(bits(recomputed.astype(np.float32)), bits(d2))


@pytest.mark.parametrize("cell,table", [(0.0, 1 << 23), (0.3, 1 << 23), (3.0, 1 << 23), (0.25, 4096)])
def test_knn_scan_vs_oracle_kdtree(G, O, vox_pair, cell, table):
    v0, v1 = vox_pair
    g = G()
    g.setGridCellSize(cell)
    g.setGridTableCells(table)     # a tiny table forces the cell edge to grow on the device
    g.setInputTarget(v0)
    tree = O.Cloud(v0)
    rng = np.random.default_rng(5)
    far = rng.uniform(-150, 150, size=(300, 3)).astype(np.float32)   # queries far outside the cloud's bbox too
    far[:, 2] *= 0.3
    q = np.vstack([v1[:6000, :3], far])
    for k in (1, 20):
        idx, d2 = g.knn(1, q, k + 1)
        ridx, rd2 = tree.knn(q, k + 1)
        assert np.array_equal(bits(d2), bits(rd2))
        m = tie_free_mask(rd2)
        assert np.array_equal(idx[:, :k][m], ridx[:, :k][m])


def test_knn_small_clouds(G, O):
    g = G()
    pts = synth.random_planes_cloud(7, seed=1)
    g.setInputTarget(pts)
    idx, d2 = g.knn(1, pts[:, :3].copy(), 10)     # fewer points than k: -1 padded like nanoflann's short result
    ridx, rd2 = O.Cloud(pts).knn(pts[:, :3].copy(), 10)
    assert np.array_equal(idx, ridx) and np.array_equal(bits(d2), bits(rd2))
    from direct_lidar_odometry_b200 import NanoGICPError
    g.setCorrespondenceRandomness(10)
    with pytest.raises(NanoGICPError) as e:
        g.calculateTargetCovariances()
    assert e.value.code == -3
    # duplicates and a single point
    dup = np.repeat(pts[:2], 5, axis=0)
    g2 = G()
    g2.setInputTarget(dup)
    idx, d2 = g2.knn(1, dup[:1, :3].copy(), 5)
    assert (d2 == 0).all() and set(idx[0]) <= set(range(10))


# ---------------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("k,method", [(10, 3), (20, 3), (20, 0), (10, 1), (10, 2), (10, 4)])
def test_covariances_vs_oracle(G, O, vox_pair, k, method):
    v0, _ = vox_pair
    g = G()
    g.setCorrespondenceRandomness(k)
    g.setRegularizationMethod(method)
    g.setInputTarget(v0)
    assert g.calculateTargetCovariances() is True
    got = g.getTargetCovariances()
    ref, ridx, rd2 = O.Cloud(v0).covariances(k, method=method, with_knn=True)
    assert got.shape == ref.shape
    assert np.abs(got[:, 3, :]).max() == 0 and np.abs(got[:, :, 3]).max() == 0
    # exclude points whose k-th and (k+1)-th neighbour distances tie exactly (neighbour SET is then ambiguous)
    _, d2p = O.Cloud(v0).knn(v0[:, :3].copy(), k + 1)
    ok = d2p[:, k - 1] != d2p[:, k]
    assert ok.mean() > 0.99
    num = np.linalg.norm((got - ref)[ok].reshape(-1, 16), axis=1)
    den = np.linalg.norm(ref[ok].reshape(-1, 16), axis=1)
    rel = num / den
    if method == 3:
        # PLANE output is basis dependent when the two smallest singular values (nearly) coincide — collinear
        # neighbourhoods; measure the gap from the raw covariance and require parity wherever it is resolvable
        raw = O.Cloud(v0).covariances(k, method=0)[:, :3, :3]
        w = np.linalg.eigvalsh(raw)[ok]
        resolvable = (w[:, 1] - w[:, 0]) > 1e-6 * w[:, 2]
        assert resolvable.mean() > 0.99
        assert rel[resolvable].max() < COV_RTOL
        ev = np.linalg.eigvalsh(0.5 * (got[:, :3, :3] + got[:, :3, :3].transpose(0, 2, 1)))
        assert np.allclose(ev, [1e-3, 1, 1], atol=1e-9)   # holds for every point, ambiguous or not
    else:
        assert rel.max() < COV_RTOL


@pytest.mark.parametrize("k", [33, 48, 64, 128])
def test_wide_k_knn_and_covariances_vs_oracle(G, O, vox_pair, k):
    """k above the warp-wide result set (the reference takes any k, nano_gicp.hpp:84): shared-memory result set on the
    warp-search path.  Same bars as k <= 32: distances bit-identical, indices identical off exact ties, covariances <= 1e-5."""
    v0, v1 = vox_pair
    g = G()
    g.setCorrespondenceRandomness(k)
    g.setKnnPath(2)                      # a forced tile path must not matter for k > 32
    g.setInputTarget(v0)
    tree = O.Cloud(v0)
    rng = np.random.default_rng(9)
    far = rng.uniform(-120, 120, size=(100, 3)).astype(np.float32)
    q = np.vstack([v1[:1500, :3], far])
    idx, d2 = g.knn(1, q, k)
    ridx, rd2 = tree.knn(q, k)
    assert np.array_equal(bits(d2), bits(rd2))
    lo = np.concatenate([np.full((len(q), 1), -1, np.float32), rd2[:, :-1]], axis=1)
    hi = np.concatenate([rd2[:, 1:], np.full((len(q), 1), np.inf, np.float32)], axis=1)
    distinct = (rd2 != lo) & (rd2 != hi)
    distinct[:, -1] = False              # the last place may tie with the (k+1)-th
    assert distinct.mean() > 0.9
    assert np.array_equal(idx[distinct], ridx[distinct])
    assert g.calculateTargetCovariances() is True
    got = g.getTargetCovariances()
    ref = tree.covariances(k, method=0)
    g0 = G()
    g0.setCorrespondenceRandomness(k); g0.setRegularizationMethod(0); g0.setInputTarget(v0)
    assert g0.calculateTargetCovariances() is True
    raw = g0.getTargetCovariances()
    _, d2p = tree.knn(v0[:, :3].copy(), k + 1)
    ok = d2p[:, k - 1] != d2p[:, k]
    assert ok.mean() > 0.99
    rel = np.linalg.norm((raw - ref)[ok].reshape(-1, 16), axis=1) / np.linalg.norm(ref[ok].reshape(-1, 16), axis=1)
    assert rel.max() < COV_RTOL
    ev = np.linalg.eigvalsh(0.5 * (got[:, :3, :3] + got[:, :3, :3].transpose(0, 2, 1)))
    assert np.allclose(ev, [1e-3, 1, 1], atol=1e-9)
    # and a registration with it: pose and iteration counts like the oracle's
    if k in (48, 128):
        gg, r = _align_both(G, O, v1, v0, dict(k=k, thr=1.0, max_iter=32, trans_eps=0.01), None, 0)
        res = gg.result
        assert (res.nr_iterations, res.converged, res.n_linearize, res.n_compute_error) == \
               (r.nr_iterations, r.converged, r.n_linearize, r.n_compute_error)
        dt, dr = pose_delta(gg.final_state(), r.Tx())
        assert dt < POSE_T_TOL and dr < POSE_R_TOL, (dt, dr)


def test_k_above_128_is_rejected(G):
    from direct_lidar_odometry_b200 import NanoGICPError
    g = G()
    with pytest.raises(NanoGICPError):
        g.setCorrespondenceRandomness(129)
    g.setCorrespondenceRandomness(128)   # and the handle still takes valid values afterwards
    g.setCorrespondenceRandomness(20)


# ---------------------------------------------------------------------------------------------- K4 / K5
def _setup_pair(G, O, v_src, v_tgt, k, thr, **kw):
    g = G()
    g.setCorrespondenceRandomness(k)
    g.setMaxCorrespondenceDistance(thr)
    for name, val in kw.items():
        getattr(g, name)(val)
    g.setInputTarget(v_tgt)
    g.setInputSource(v_src)
    o = O.Gicp(k=k, max_corr_dist=thr, num_threads=0)
    o.set_target(O.Cloud(v_tgt))
    o.set_source(O.Cloud(v_src))
    return g, o


@pytest.mark.parametrize("thr", [1.0, float(np.finfo(np.float32).max)])
def test_linearize_and_error_vs_oracle(G, O, vox_pair, thr):
    v0, v1 = vox_pair
    g, o = _setup_pair(G, O, v1, v0, 10, thr)
    g.calculateTargetCovariances(); g.calculateSourceCovariances()
    o.calc_target_covs(); o.calc_source_covs()
    # feed the GPU the oracle's covariances so this test isolates K4/K5
    g.setSourceCovariances(o.get_source_covs()); g.setTargetCovariances(o.get_target_covs())
    T = synth.perturb_pose(np.eye(4), (0.08, -0.03, 0.02), 0.4)
    lg = g.linearize(T, per_point=True)
    lo = o.linearize(T, per_point=True)
    matched = lo["corr"] >= 0
    assert matched.mean() > 0.5
    assert np.array_equal(lg["corr"] >= 0, matched)
    assert np.array_equal(bits(lg["sqd"][matched]), bits(lo["sqd"][matched]))
    same = lg["corr"] == lo["corr"]
    assert same.mean() > 0.9995          # the rest are exact 1-NN distance ties
    scale = np.abs(lo["H"]).max()
    if same.all():
        assert np.abs(lg["H"] - lo["H"]).max() < 1e-9 * scale
        assert np.abs(lg["b"] - lo["b"]).max() < 1e-9 * max(1.0, np.abs(lo["b"]).max())
        assert abs(lg["err"] - lo["err"]) < 1e-9 * abs(lo["err"])
    else:
        assert np.abs(lg["H"] - lo["H"]).max() < 1e-3 * scale
    mm = lo["mahalanobis"][same & matched]
    assert np.abs(lg["mahalanobis"][same & matched] - mm).max() < 1e-9 * np.abs(mm).max()
    # frozen-correspondence error at another transform
    T2 = synth.perturb_pose(T, (0.01, 0.01, 0.0), 0.05)
    eg, eo = g.compute_error(T2), o.compute_error(T2)
    assert abs(eg - eo) < (1e-9 if same.all() else 1e-3) * abs(eo)


# ---------------------------------------------------------------------------------------------- LM
CONFIGS = {
    # cfg/params.yaml:54-58 (S2S) and :63-67 (S2M); library defaults nano_gicp_impl.hpp:57-59
    "dlo_s2s": dict(k=10, thr=1.0, max_iter=32, trans_eps=0.01),
    "dlo_s2m": dict(k=20, thr=0.5, max_iter=32, trans_eps=0.01),
    "defaults": dict(k=20, thr=float(np.finfo(np.float32).max), max_iter=64, trans_eps=5e-4),
}


def _align_both(G, O, v_src, v_tgt, cfg, guess, mode, optimizer=1):
    g = G()
    g.setCorrespondenceRandomness(cfg["k"]); g.setMaxCorrespondenceDistance(cfg["thr"])
    g.setMaximumIterations(cfg["max_iter"]); g.setTransformationEpsilon(cfg["trans_eps"])
    g.setOptimizer(optimizer)
    g.setAlignMode(mode)
    g.setInputTarget(v_tgt); g.setInputSource(v_src)
    g.align(guess)
    o = O.Gicp(k=cfg["k"], max_corr_dist=cfg["thr"], max_iter=cfg["max_iter"], trans_eps=cfg["trans_eps"], optimizer=optimizer,
               num_threads=1)
    o.set_target(O.Cloud(v_tgt)); o.set_source(O.Cloud(v_src))
    r = o.align(guess)
    return g, r


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("cfg", ["dlo_s2s", "defaults"])
def test_align_s2s_pose_and_iterations(G, O, vox_pair, scan_pair, cfg, mode):
    v0, v1 = vox_pair
    for guess in (None, synth.perturb_pose(np.eye(4), (0.3, 0.1, 0.0), 2.0).astype(np.float32)):
        g, r = _align_both(G, O, v1, v0, CONFIGS[cfg], guess, mode)
        res = g.result
        assert (res.nr_iterations, res.converged, res.n_linearize, res.n_compute_error, res.lm_failed) == \
               (r.nr_iterations, r.converged, r.n_linearize, r.n_compute_error, r.lm_failed)
        dt, dr = pose_delta(g.final_state(), r.Tx())
        assert dt < POSE_T_TOL and dr < POSE_R_TOL, (dt, dr)
        dt, dr = pose_delta(g.getFinalTransformation(), r.T())
        assert dt < POSE_T_TOL and dr < POSE_R_TOL
        assert np.allclose(g.getFinalHessian(), r.H(), rtol=1e-6, atol=1e-6 * np.abs(r.H()).max())
        assert g.hasConverged() == bool(r.converged)
    truth = np.linalg.inv(scan_pair["T0"]) @ scan_pair["T1"]
    assert np.abs(g.final_state()[:3, 3] - truth[:3, 3]).max() < 2e-2


def test_align_gauss_newton(G, O, vox_pair):
    v0, v1 = vox_pair
    g, r = _align_both(G, O, v1, v0, CONFIGS["dlo_s2s"], None, 0, optimizer=0)
    assert g.result.nr_iterations == r.nr_iterations and g.result.n_compute_error == 0
    dt, dr = pose_delta(g.final_state(), r.Tx())
    assert dt < POSE_T_TOL and dr < POSE_R_TOL


@pytest.fixture(scope="module")
def small_submap(O):
    """A scaled-down S2M case: 6 keyframes 5 m apart voxelised at 0.5 m in the world frame (reference
    odom.cc:484-490), and one scan voxelised at 0.25 m as the source."""
    from util import make_small_submap
    return make_small_submap(O)


@pytest.mark.parametrize("mode", [0, 1])
def test_align_s2m_vs_oracle(G, O, small_submap, mode):
    submap, scan, T = small_submap
    guess = synth.perturb_pose(T, (0.2, 0.0, 0.0), 1.0).astype(np.float32)
    g, r = _align_both(G, O, scan, submap, CONFIGS["dlo_s2m"], guess, mode)
    res = g.result
    assert (res.nr_iterations, res.converged, res.n_linearize, res.n_compute_error) == \
           (r.nr_iterations, r.converged, r.n_linearize, r.n_compute_error)
    dt, dr = pose_delta(g.final_state(), r.Tx())
    assert dt < POSE_T_TOL and dr < POSE_R_TOL, (dt, dr)
    dt, dr = pose_delta(g.final_state(), T)
    assert dt < 0.05 and dr < 2e-3   # and it actually localises against the map


def test_align_edge_cases(G, O):
    from direct_lidar_odometry_b200 import NanoGICPError
    a = synth.random_planes_cloud(500, seed=1)
    b = a.copy(); b[:, :3] += 500.0
    for mode in (0, 1):
        g = G()
        g.setCorrespondenceRandomness(10); g.setMaxCorrespondenceDistance(1.0); g.setAlignMode(mode)
        g.setInputTarget(a); g.setInputSource(b)
        g.align()
        # no correspondence inside the threshold: H=b=0, d=0, rho=NaN -> accepted, converged at iteration 0
        assert g.result.nr_iterations == 0 and g.result.converged == 1
        assert np.array_equal(g.final_state(), np.eye(4))
    g = G()
    g.setInputSource(a)
    assert g.align() is None            # no target: PCL prints an error and returns
    g.setInputTarget(a[:5])
    g.setCorrespondenceRandomness(10)
    with pytest.raises(NanoGICPError):  # fewer than k points: UB in the reference, an error here
        g.align()
    # output cloud = source transformed by the final transformation
    g2 = G()
    g2.setCorrespondenceRandomness(10)
    g2.setInputTarget(a); g2.setInputSource(a)
    out = g2.align(want_output=True)
    assert out.shape == (500, 4) and np.allclose(out[:, :3], a[:, :3], atol=1e-4) and (out[:, 3] == 1).all()


# ---------------------------------------------------------------------------------------------- OdomNode call sequence
def test_odom_call_sequence_matches_oracle(G, O):
    """The S2S/S2M call pattern of OdomNode (reference src/dlo/odom.cc:472-528, 792-852) on 5 scans:
    two registration objects, shared source index, covariance hand-over, swapSourceAndTarget."""
    scans, poses = [], []
    for i in range(5):
        T = synth.trajectory_pose(i * 4)
        scans.append(O.voxel_filter(synth.crop_box_negative(synth.os1_like(i * 4, T)), 0.25))
        poses.append(T)
    s2s, s2m = G(), G()
    for obj, c in ((s2s, CONFIGS["dlo_s2s"]), (s2m, CONFIGS["dlo_s2m"])):
        obj.setCorrespondenceRandomness(c["k"]); obj.setMaxCorrespondenceDistance(c["thr"])
        obj.setMaximumIterations(c["max_iter"]); obj.setTransformationEpsilon(c["trans_eps"])
    o_s2s = O.Gicp(k=10, max_corr_dist=1.0, max_iter=32, trans_eps=0.01, num_threads=0)
    o_s2m = O.Gicp(k=20, max_corr_dist=0.5, max_iter=32, trans_eps=0.01, num_threads=0)

    # initializeInputTarget (odom.cc:472-507)
    s2s.setInputTarget(scans[0]); s2s.calculateTargetCovariances()
    oc0 = O.Cloud(scans[0])
    o_s2s.set_target(oc0); o_s2s.calc_target_covs()
    T0 = poses[0].astype(np.float32)
    key = s2s.voxel_filter(synth.transform_xyzi(scans[0], T0), 0.5)
    key_ref = O.voxel_filter(synth.transform_xyzi(scans[0], T0), 0.5)
    assert np.array_equal(bits(key), bits(key_ref))
    s2s.setInputSource(key); s2s.calculateSourceCovariances()
    key_covs = s2s.getSourceCovariances()
    ock = O.Cloud(key_ref)
    o_s2s.set_source(ock); o_s2s.calc_source_covs()
    key_covs_ref = o_s2s.get_source_covs()
    assert np.abs(key_covs - key_covs_ref).max() < 1e-5

    T_s2s_prev = T0.copy()
    T_s2s_prev_ref = T0.copy()
    first = True
    for i in range(1, 5):
        cur = scans[i]
        # setInputSources (odom.cc:514-528)
        s2s.setInputSource(cur)
        s2m.registerInputSource(cur)
        s2m.source_kdtree_ = s2s.source_kdtree_
        s2m.source_covs_.clear()
        occ = O.Cloud(cur)
        o_s2s.set_source(occ); o_s2m.set_source(occ)
        # getNextPose (odom.cc:792-852)
        s2s.align()
        r = o_s2s.align()
        assert (s2s.result.nr_iterations, s2s.result.n_compute_error) == (r.nr_iterations, r.n_compute_error)
        dt, dr = pose_delta(s2s.final_state(), r.Tx())
        assert dt < POSE_T_TOL and dr < POSE_R_TOL
        T_s2s = T_s2s_prev @ s2s.getFinalTransformation()          # propagateS2S, float (odom.cc:928)
        T_s2s_ref = T_s2s_prev_ref @ r.T()
        s2m.source_covs_ = s2s.source_covs_                         # odom.cc:815, stays in HBM
        o_s2m.set_source_covs(o_s2s.get_source_covs())
        s2s.swapSourceAndTarget(); o_s2s.swap()
        if first:                                                   # submap_hasChanged
            s2m.setInputTarget(key); s2m.setTargetCovariances(key_covs)
            o_s2m.set_target(ock); o_s2m.set_target_covs(key_covs_ref)
            first = False
        s2m.align(T_s2s)
        r2 = o_s2m.align(T_s2s_ref)
        assert (s2m.result.nr_iterations, s2m.result.n_compute_error) == (r2.nr_iterations, r2.n_compute_error)
        dt, dr = pose_delta(s2m.final_state(), r2.Tx())
        assert dt < 5 * POSE_T_TOL and dr < 5 * POSE_R_TOL   # the two chains start from slightly different float guesses
        T_s2s_prev = s2m.getFinalTransformation()
        T_s2s_prev_ref = r2.T()
        dt, dr = pose_delta(T_s2s_prev, poses[i])
        assert dt < 0.1 and dr < 5e-3                            # odometry stays on the true trajectory


# ---------------------------------------------------------------------------------------------- full-size properties
def test_full_size_submap_properties(G, O):
    """BASELINE config C2 size (500k-point submap, k=20): properties that do not need the oracle at full size,
    plus an oracle spot check on a sample of queries."""
    rng = np.random.default_rng(11)
    base = synth.random_planes_cloud(500_000, seed=4, extent=150.0, noise=0.02)
    g = G()
    g.setCorrespondenceRandomness(20)
    g.setInputTarget(base)
    sample = rng.permutation(base.shape[0])[:3000]
    q = np.ascontiguousarray(base[sample, :3])
    idx, d2 = g.knn(1, q, 20)
    assert (idx[:, 0] == sample).mean() > 0.999 and (d2[:, 0] == 0).all()      # self first
    assert (np.diff(d2, axis=1) >= 0).all()                                      # ascending
    ridx, rd2 = O.Cloud(base).knn(q, 21)
    assert np.array_equal(bits(d2), bits(rd2[:, :20]))
    m = tie_free_mask(rd2)
    assert np.array_equal(idx[m], ridx[:, :20][m])
    g.calculateTargetCovariances()
    covs = g.getTargetCovariances()
    ev = np.linalg.eigvalsh(0.5 * (covs[:, :3, :3] + covs[:, :3, :3].transpose(0, 2, 1)))
    assert np.allclose(ev, [1e-3, 1, 1], atol=1e-8)


# ---------------------------------------------------------------------------------------------- N1 keyframe store
def test_keyframe_store_submap_equals_host_concat(G, O, scan_pair):
    """Submap assembled on the device from stored keyframes (ngicp_kfstore_*) == OdomNode's host concatenation of
    keyframe clouds + keyframe_normals followed by setInputTarget/setTargetCovariances (odom.cc:1315-1328,830-833)."""
    from direct_lidar_odometry_b200 import KeyframeStore
    s2s, a, b = G(), G(), G()
    s2s.setCorrespondenceRandomness(10)
    for g in (a, b):
        g.setCorrespondenceRandomness(20); g.setMaxCorrespondenceDistance(0.5)
        g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
    store = KeyframeStore(0)
    host = []
    for i in range(4):
        T = synth.trajectory_pose(33 * i)
        kf = O.voxel_filter(synth.transform_xyzi(synth.crop_box_negative(synth.os1_like(33 * i, T)), T.astype(np.float32)), 0.5)
        s2s.setInputSource(kf)
        s2s.calculateSourceCovariances()
        assert store.push(s2s) == i and store.points(i) == kf.shape[0]
        host.append((kf, s2s.getSourceCovariances()))
    assert len(store) == 4
    Ts = synth.trajectory_pose(40)
    scan = O.voxel_filter(synth.crop_box_negative(synth.os1_like(40, Ts)), 0.25)
    guess = synth.perturb_pose(Ts, (0.15, 0.0, 0.0), 0.5).astype(np.float32)
    for sel in ([0, 1, 2, 3], [2, 0], [3]):
        store.set_target(a, sel)
        b.clearTarget()
        b.setInputTarget(np.ascontiguousarray(np.vstack([host[i][0] for i in sel])))
        b.setTargetCovariances(np.concatenate([host[i][1] for i in sel]))
        assert np.array_equal(a.getTargetCovariances(), b.getTargetCovariances())
        for g in (a, b):
            g.clearSource(); g.setInputSource(scan); g.calculateSourceCovariances(); g.align(guess)
        assert np.array_equal(a.final_state(), b.final_state())
        assert (a.result.nr_iterations, a.result.n_compute_error) == (b.result.nr_iterations, b.result.n_compute_error)
    with pytest.raises(Exception):
        store.set_target(a, [7])


# ---------------------------------------------------------------------------------------------- sharded submap (one GPU)
def test_sharded_partials_sum_to_unsharded(G, O, small_submap):
    """Two slabs of the target on two handles of ONE GPU (ranks emulated sequentially): the summed partial
    {H,b,err} equal the unsharded linearisation and the host-stepped sharded LM reproduces the fused align."""
    from direct_lidar_odometry_b200 import sharded
    submap, scan, T = small_submap
    thr = 0.5
    tc = O.Cloud(submap).covariances(20)
    sc = O.Cloud(scan).covariances(20)
    guess = synth.perturb_pose(T, (0.2, 0.0, 0.0), 1.0).astype(np.float32)
    world = 2
    backs = []
    for r in range(world):
        pts, covs, axis, lo, hi = sharded.shard_target(submap, tc, r, world, halo=thr + 0.01)
        assert pts.shape[0] < submap.shape[0]
        be = sharded.CudaShardBackend(0, k=20, max_corr_dist=thr)
        be.set_target(pts, covs, axis, lo, hi)
        be.set_source(scan, sc)
        backs.append(be)

    class Both:
        def linearize_partial(self, T_):
            return sum(np.asarray(b.linearize_partial(T_)) for b in backs)

        def compute_error_partial(self, T_):
            return sum(b.compute_error_partial(T_) for b in backs)

    g = G()
    g.setCorrespondenceRandomness(20); g.setMaxCorrespondenceDistance(thr)
    g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
    g.setInputTarget(submap); g.setTargetCovariances(tc)
    g.setInputSource(scan); g.setSourceCovariances(sc)
    whole = g.linearize_partial(np.asarray(guess, dtype=np.float64))
    parts = Both().linearize_partial(np.asarray(guess, dtype=np.float64))
    assert np.abs(parts - whole).max() < 1e-9 * np.abs(whole).max()
    al = sharded.ShardedSubmapAligner(Both(), max_corr_dist=thr, max_iter=32, trans_eps=0.01)
    res = al.align(guess)
    g.align(guess)
    assert (res["nr_iterations"], res["n_linearize"], res["n_compute_error"], res["converged"]) == \
           (g.result.nr_iterations, g.result.n_linearize, g.result.n_compute_error, g.result.converged)
    assert np.abs(res["final_x"] - g.final_state()).max() < 1e-9


def test_sharded_min_exchange_unbounded_distance(G, O, small_submap):
    """The library-default correspondence distance is FLT_MAX (nano_gicp_impl.hpp:59): no halo makes slabs exact, so the
    ranks exchange nearest neighbours — ngicp_nn1_packed per rank, element-wise minimum (what the min-all-reduce does),
    ngicp_linearize_won per rank, sum.  Two halves of the target (a plain partition, no overlap) on two handles of ONE
    GPU, ranks emulated sequentially: the minimum equals the whole target's nearest distances bit for bit, the summed
    {H, b, err} equal the unsharded linearisation, and the host-stepped LM reproduces the unsharded fused align and the
    CPU oracle."""
    from direct_lidar_odometry_b200 import sharded
    submap, scan, T = small_submap
    inf = 3.0e38
    tc = O.Cloud(submap).covariances(20)
    sc = O.Cloud(scan).covariances(20)
    guess = synth.perturb_pose(T, (0.2, 0.0, 0.0), 1.0).astype(np.float32)
    axis = int(np.argmax(submap[:, :3].max(0) - submap[:, :3].min(0)))
    order = np.argsort(submap[:, axis], kind="stable")
    halves = [order[:order.size // 2], order[order.size // 2:]]
    backs = []
    for r in range(2):
        be = sharded.CudaShardBackend(0, k=20, max_corr_dist=inf)
        be.set_target_part(np.ascontiguousarray(submap[halves[r]]), np.ascontiguousarray(tc[halves[r]]))
        be.set_source(scan, sc)
        backs.append(be)

    class Both:
        def nn1_packed(self, T_, rank):
            return np.minimum(backs[0].nn1_packed(T_, 0), backs[1].nn1_packed(T_, 1))

        def linearize_won(self, T_, rank, won):
            return sum(np.asarray(b.linearize_won(T_, r, won)) for r, b in enumerate(backs))

        def compute_error_partial(self, T_):
            return sum(b.compute_error_partial(T_) for b in backs)

    g = G()
    g.setCorrespondenceRandomness(20); g.setMaxCorrespondenceDistance(inf)
    g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
    g.setInputTarget(submap); g.setTargetCovariances(tc)
    g.setInputSource(scan); g.setSourceCovariances(sc)
    Td = np.asarray(guess, dtype=np.float64)
    won = Both().nn1_packed(Td, 0)
    whole_nn = g.nn1_packed(Td, 7)
    assert np.array_equal(won >> np.uint64(32), whole_nn >> np.uint64(32))            # global nearest distances, bit for bit
    assert set(np.unique(won & np.uint64(0xffffffff)).tolist()) == {0, 1}            # both ranks win some points
    lin = g.linearize(Td)
    assert np.array_equal(lin["sqd"].view(np.uint32), (whole_nn >> np.uint64(32)).astype(np.uint32))   # = update_correspondences' distances
    whole = g.linearize_partial(Td)
    parts = Both().linearize_won(Td, 0, won)
    assert np.abs(parts - whole).max() < 1e-9 * np.abs(whole).max()
    al = sharded.ShardedSubmapAligner(Both(), max_corr_dist=inf, max_iter=32, trans_eps=0.01)
    assert al.exchange == "min"
    res = al.align(guess)
    g.align(guess)
    assert (res["nr_iterations"], res["n_linearize"], res["n_compute_error"], res["converged"]) == \
           (g.result.nr_iterations, g.result.n_linearize, g.result.n_compute_error, g.result.converged)
    assert np.abs(res["final_x"] - g.final_state()).max() < 1e-9
    o = O.Gicp(k=20, max_corr_dist=inf, max_iter=32, trans_eps=0.01, num_threads=1)
    o.set_target(O.Cloud(submap)); o.set_target_covs(tc)
    o.set_source(O.Cloud(scan)); o.set_source_covs(sc)
    r = o.align(guess)
    assert (res["nr_iterations"], res["n_linearize"], res["n_compute_error"]) == (r.nr_iterations, r.n_linearize, r.n_compute_error)
    dt, dr = pose_delta(res["final_x"], r.Tx())
    assert dt < POSE_T_TOL and dr < POSE_R_TOL


def test_sharded_align_fused_exchange_one_gpu():
    """The exchange fused into the persistent LM kernel (ngicp_comm_*): two handles on ONE GPU each hold one slab of
    the target, their kernels run concurrently (grids capped so that both are co-resident) and meet in each other's
    exchange buffers.  Runs in a fresh process: in a long-lived one with dozens of streams two streams can share a
    hardware queue, which would serialise the two persistent kernels (real deployments run one process per GPU)."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, NGICP_ALIGN_MAX_BLOCKS="24", NGICP_COMM_TIMEOUT_MS="3000", CUDA_DEVICE_MAX_CONNECTIONS="32")
    p = subprocess.run([sys.executable, os.path.join(here, "sharded_twin_main.py")], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    assert r["ranks_bit_identical"] and r["counts_equal_unsharded"] and r["max_abs_dT_vs_unsharded"] < 1e-9
    assert r["timeout_reported"]
    assert r["recovered_after_reset"], r                       # ngicp_comm_reset on every rank, connections kept
