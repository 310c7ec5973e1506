"""CPU tests that pin and sanity-check the oracle.  No GPU."""
import os
import numpy as np
import pytest

from direct_lidar_odometry_b200 import synth
from util import tie_free_mask


def _libs(orc):
    libs = [("restated", orc.load(prefer_ref=False), orc.BACKEND_RESTATED)]
    if os.path.exists(orc.lib_path(True)):
        libs.append(("ref-nanoflann", orc.load(prefer_ref=True), orc.BACKEND_REF))
    return libs


def test_knn_matches_reference_golden(orc, golden_knn):
    """The restated kd-tree (and the _ref build when present) reproduce the vectors produced by the
    reference's own nanoflann bit-for-bit: distances AND indices, ties included."""
    for name, L, backend in _libs(orc):
        c = orc.Cloud(golden_knn["cloud"], backend, lib=L)
        for k in (1, 5, 10, 20):
            idx, d2 = c.knn(golden_knn["queries"], k, nthreads=2)
            assert np.array_equal(d2, golden_knn[f"d2_k{k}"]), (name, k)
            assert np.array_equal(idx, golden_knn[f"idx_k{k}"]), (name, k)


def test_knn_equals_bruteforce_float(orc, golden_knn):
    L = orc.load(prefer_ref=False)
    cloud, q = golden_knn["cloud"], golden_knn["queries"]
    tree = orc.Cloud(cloud, orc.BACKEND_RESTATED, lib=L)
    brute = orc.Cloud(cloud, orc.BACKEND_BRUTE, lib=L)
    for k in (1, 20):
        it, dt = tree.knn(q, k + 1)
        ib, db = brute.knn(q, k + 1)
        assert np.array_equal(dt, db)
        distinct = tie_free_mask(dt)
        assert distinct.mean() > 0.9
        assert np.array_equal(it[:, :k][distinct], ib[:, :k][distinct])
    # float metric: ((dx*dx)+dy*dy)+dz*dz, unfused (nanoflann_impl.hpp:441-449)
    i1, d1 = tree.knn(q[:50], 1)
    p = cloud[i1[:, 0]]
    dx, dy, dz = (q[:50, 0] - p[:, 0]), (q[:50, 1] - p[:, 1]), (q[:50, 2] - p[:, 2])
    ref = (dx * dx + dy * dy).astype(np.float32) + (dz * dz).astype(np.float32)
    assert np.array_equal(d1[:, 0], ref.astype(np.float32))


def test_knn_fewer_points_than_k(orc):
    pts = synth.random_planes_cloud(7, seed=1)
    c = orc.Cloud(pts, orc.BACKEND_RESTATED, lib=orc.load(False))
    idx, d2 = c.knn(pts[:3, :3], 10)
    assert (idx[:, :7] >= 0).all() and (idx[:, 7:] == -1).all()
    with pytest.raises(RuntimeError):
        c.covariances(10)  # reference is UB here (nano_gicp_impl.hpp:315-318); oracle rejects


def test_svd_properties(orc):
    import ctypes as C
    L = orc.load(False)
    rng = np.random.default_rng(0)
    for t in range(200):
        A = rng.normal(size=(3, 3))
        if t % 3 == 0:
            A = A @ A.T  # symmetric PSD like a covariance
        if t % 7 == 0:
            A[:, 2] = A[:, 0] * 2  # rank deficient
        U, S, V = np.zeros(9), np.zeros(3), np.zeros(9)
        Ac = np.ascontiguousarray(A.T).reshape(9)
        L.orc_svd3(orc._d(Ac), orc._d(U), orc._d(S), orc._d(V))
        U, V = U.reshape(3, 3).T, V.reshape(3, 3).T
        assert np.allclose(U @ np.diag(S) @ V.T, A, atol=1e-12 * max(1, np.abs(A).max()))
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-12) and np.allclose(V.T @ V, np.eye(3), atol=1e-12)
        assert S[0] >= S[1] >= S[2] >= 0
        assert np.allclose(S, np.linalg.svd(A, compute_uv=False), atol=1e-12 * max(1, S[0]))


def test_ldlt_so3_inverse(orc):
    L = orc.load(False)
    rng = np.random.default_rng(1)
    for _ in range(50):
        B = rng.normal(size=(6, 6))
        A = B @ B.T + 1e-3 * np.eye(6)
        b = rng.normal(size=6)
        x = np.zeros(6)
        L.orc_ldlt6_solve(orc._d(np.ascontiguousarray(A.T).reshape(36)), orc._d(b), orc._d(x))
        assert np.allclose(A @ x, b, atol=1e-9 * np.abs(b).max() * np.linalg.cond(A))
        w = rng.normal(size=3) * rng.choice([1e-7, 0.1, 2.0])
        R = np.zeros(9)
        L.orc_so3_exp(orc._d(w), orc._d(R))
        R = R.reshape(3, 3).T
        from scipy.spatial.transform import Rotation
        assert np.allclose(R, Rotation.from_rotvec(w).as_matrix(), atol=1e-12)
        M = rng.normal(size=(4, 4)) + 3 * np.eye(4)
        Mi = np.zeros(16)
        L.orc_inverse4(orc._d(np.ascontiguousarray(M.T).reshape(16)), orc._d(Mi))
        assert np.allclose(Mi.reshape(4, 4).T @ M, np.eye(4), atol=1e-10)
    # all-zero system: Eigen's LDLT solve returns 0 (pivots below tolerance are zeroed)
    x = np.ones(6)
    L.orc_ldlt6_solve(orc._d(np.zeros(36)), orc._d(np.ones(6)), orc._d(x))
    assert np.array_equal(x, np.zeros(6))


def test_voxel_filter_properties(orc, scan_pair):
    s0 = scan_pair["s0"]
    out, assign, rc = orc.voxel_filter(s0, 0.25, return_assignment=True)
    assert rc == 0 and 15000 < out.shape[0] < 30000
    assert (out[:, 3] == 1.0).all() and (out[:, 5:] == 0).all()
    # PCL formula for the voxel id of each input point; outputs ordered by ascending id
    inv = np.float32(1.0) / np.float32(0.25)
    mn = np.floor(s0[:, :3].min(0) * inv).astype(np.int64)
    mx = np.floor(s0[:, :3].max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    ijk = (np.floor(s0[:, :3] * inv) - mn.astype(np.float32)).astype(np.int64)
    vid = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uniq, inverse = np.unique(vid, return_inverse=True)
    assert out.shape[0] == uniq.shape[0]
    assert np.array_equal(assign, inverse.astype(np.int32))
    # centroid = float32 running sum in input order / count
    for v in (0, 17, out.shape[0] // 2, out.shape[0] - 1):
        members = np.nonzero(inverse == v)[0]
        acc = np.zeros(4, dtype=np.float32)
        for m in members:
            acc += s0[m, [0, 1, 2, 4]]
        acc = acc / np.float32(len(members))
        assert np.array_equal(acc, out[v, [0, 1, 2, 4]])
    # idempotent-ish: every centroid lies inside its voxel => filtering again keeps the count
    out2 = orc.voxel_filter(out, 0.25)
    assert out2.shape[0] <= out.shape[0]
    # empty and single-point inputs
    assert orc.voxel_filter(np.zeros((0, 8), np.float32), 0.25).shape[0] == 0
    one = orc.voxel_filter(s0[:1], 0.25)
    assert one.shape[0] == 1 and np.array_equal(one[0, :3], s0[0, :3])
    # index overflow => pass-through (PCL warns and copies the input)
    far = s0[:100].copy()
    far[0, :3] = 1e6
    o, _, rc = orc.voxel_filter(far, 0.01, return_assignment=True)
    assert rc == 1 and o.shape[0] == 100


def test_plane_covariances(orc, scan_pair):
    v0 = orc.voxel_filter(scan_pair["s0"], 0.25)
    c = orc.Cloud(v0, lib=orc.load(False), backend=orc.BACKEND_RESTATED)
    covs, idx, d2 = c.covariances(10, with_knn=True)
    assert (idx[:, 0] == np.arange(v0.shape[0])).all() and (d2[:, 0] == 0).all()  # self is a neighbour
    assert np.abs(covs[:, 3, :]).max() == 0 and np.abs(covs[:, :, 3]).max() == 0
    w = np.linalg.eigvalsh(0.5 * (covs[:, :3, :3] + covs[:, :3, :3].transpose(0, 2, 1)))
    assert np.allclose(w, [1e-3, 1, 1], atol=1e-9)
    # the small axis is the least-squares plane normal of the neighbourhood
    i = 1234
    nb = v0[idx[i], :3].astype(np.float64)
    C = np.cov(nb.T, bias=True)
    n_ref = np.linalg.eigh(C)[1][:, 0]
    n_got = np.linalg.eigh(covs[i, :3, :3])[1][:, 0]
    assert abs(abs(n_ref @ n_got) - 1) < 1e-8
    # NONE returns the raw covariance divided by k
    raw = c.covariances(10, method=orc.REG_NONE)
    assert np.allclose(raw[i, :3, :3], C, atol=1e-12)


def test_linearize_consistency(orc, scan_pair):
    """H/b agree with finite differences of the frozen-correspondence cost; M is the inverse of C_B + R C_A R^T."""
    v0 = orc.voxel_filter(scan_pair["s0"], 0.5)
    v1 = orc.voxel_filter(scan_pair["s1"], 0.5)
    L = orc.load(False)
    tgt, src = orc.Cloud(v0, lib=L), orc.Cloud(v1, lib=L)
    g = orc.Gicp(lib=L, k=10, max_corr_dist=1.0, num_threads=1)
    g.set_target(tgt); g.set_source(src)
    g.calc_target_covs(); g.calc_source_covs()
    T = synth.perturb_pose(np.eye(4), (0.05, -0.02, 0.01), 0.3)
    lin = g.linearize(T, per_point=True)
    assert (lin["corr"] >= 0).mean() > 0.5
    i = int(np.nonzero(lin["corr"] >= 0)[0][10])
    CA, CB = g.get_source_covs()[i], g.get_target_covs()[lin["corr"][i]]
    RCR = CB[:3, :3] + T[:3, :3] @ CA[:3, :3] @ T[:3, :3].T
    assert np.allclose(lin["mahalanobis"][i][:3, :3] @ RCR, np.eye(3), atol=1e-9)
    assert np.abs(lin["mahalanobis"][i][3]).max() == 0
    # threshold semantics: sq distance < thr^2
    assert ((lin["sqd"] < 1.0) == (lin["corr"] >= 0)).all()
    # finite differences: cost(exp(d) T) ~ y0 + 2 b.d + d.H.d  (error uses frozen M/corr)
    from scipy.spatial.transform import Rotation
    y0 = lin["err"]
    assert abs(g.compute_error(T) - y0) <= 1e-9 * abs(y0)
    eps = 1e-6
    grad = np.zeros(6)
    for a in range(6):
        d = np.zeros(6); d[a] = eps
        def cost(dd):
            D = np.eye(4); D[:3, :3] = Rotation.from_rotvec(dd[:3]).as_matrix(); D[:3, 3] = dd[3:]
            return g.compute_error(D @ T)
        grad[a] = (cost(d) - cost(-d)) / (2 * eps)
    assert np.allclose(grad, 2 * lin["b"], rtol=1e-4, atol=1e-4 * np.abs(lin["b"]).max())
    assert np.allclose(lin["H"], lin["H"].T, atol=1e-9 * np.abs(lin["H"]).max())
    assert np.all(np.linalg.eigvalsh(lin["H"]) > 0)


def test_align_recovers_motion_and_is_thread_invariant(orc, scan_pair):
    v0 = orc.voxel_filter(scan_pair["s0"], 0.25)
    v1 = orc.voxel_filter(scan_pair["s1"], 0.25)
    L = orc.load(False)
    tgt, src = orc.Cloud(v0, lib=L), orc.Cloud(v1, lib=L)
    truth = np.linalg.inv(scan_pair["T0"]) @ scan_pair["T1"]
    res = {}
    for nt in (1, 4):
        g = orc.Gicp(lib=L, k=10, max_corr_dist=1.0, max_iter=32, trans_eps=0.01, num_threads=nt)
        g.set_target(tgt); g.set_source(src)
        r = g.align()
        assert r.converged == 1 and r.lm_failed == 0
        assert r.n_linearize == r.nr_iterations + 1
        assert np.abs(r.Tx()[:3, 3] - truth[:3, 3]).max() < 5e-3
        res[nt] = r
    assert res[1].nr_iterations == res[4].nr_iterations
    assert np.allclose(res[1].Tx(), res[4].Tx(), atol=1e-9)
    # final transformation is the double state cast to float (lsq_registration_impl.hpp:113)
    assert np.array_equal(res[1].T(), res[1].Tx().astype(np.float32))
    # swap + re-align from the inverse gives roughly the inverse motion
    g.swap()
    r2 = g.align()
    assert np.abs(r2.Tx()[:3, 3] + truth[:3, 3]).max() < 2e-2


def test_align_without_correspondences_is_identity_step(orc):
    """No match within the threshold: H=b=0, d=0, rho=NaN -> accepted, converged at iteration 0."""
    L = orc.load(False)
    a = synth.random_planes_cloud(500, seed=1)
    b = a.copy(); b[:, :3] += 500.0
    g = orc.Gicp(lib=L, k=10, max_corr_dist=1.0, num_threads=1)
    g.set_target(orc.Cloud(a, lib=L)); g.set_source(orc.Cloud(b, lib=L))
    r = g.align()
    assert r.nr_iterations == 0 and r.converged == 1
    assert np.array_equal(r.Tx(), np.eye(4))


def test_preprocess_points_restatement():
    """removeNaN + negative CropBox + voxel grid (odom.cc:443-465): box membership is inclusive like pcl::CropBox."""
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    pts = np.zeros((2000, 8), np.float32)
    pts[:, :3] = rng.uniform(-4, 4, size=(2000, 3))
    pts[:, 3] = 1.0
    pts[:, 4] = rng.uniform(0, 1, 2000)
    pts[0, :3] = (1.0, 1.0, -1.0)      # on the box -> inside -> dropped
    pts[1, :3] = (1.0000001, 0.0, 0.0)  # just outside -> kept
    pts[2, 0] = np.nan
    kept = orc.preprocess_points(pts, 1.0, 0.0)
    inside = (np.abs(pts[:, :3]) <= 1.0).all(axis=1)
    finite = np.isfinite(pts[:, :3]).all(axis=1)
    assert kept.shape[0] == int((finite & ~inside).sum())
    assert not ((np.abs(kept[:, :3]) <= 1.0).all(axis=1)).any()
    assert (kept[:, :3] == pts[1, :3]).all(axis=1).any() and not (kept[:, :3] == pts[0, :3]).all(axis=1).any()
    v = orc.preprocess_points(pts, 1.0, 0.5)
    assert np.array_equal(v, orc.voxel_filter(pts[finite & ~inside], 0.5))


def test_from_ros_msg_restatement_and_field_mapping():
    """pcl::fromROSMsg for PointXYZI (odom.cc:636-637): members are matched by name + FLOAT32 + count 1; padding,
    foreign fields, unaligned records and row padding do not matter; a missing intensity stays 0."""
    from oracle import oracle as orc
    from direct_lidar_odometry_b200 import pointcloud2 as pc2
    rng = np.random.default_rng(11)
    pts = np.zeros((64 * 96, 8), np.float32)
    pts[:, :3] = rng.normal(0, 20, size=(pts.shape[0], 3))
    pts[:, 3] = 1.0
    pts[:, 4] = rng.uniform(0, 1, pts.shape[0])
    pts[7, 2] = np.nan
    for kind, height, pad in (("ouster", 64, 0), ("ouster", 1, 0), ("velodyne", 1, 0), ("velodyne", 64, 6), ("xyz", 64, 3)):
        msg = pc2.make_pointcloud2(pts, kind, height=height, row_pad=pad)
        got = orc.from_ros_msg(msg)
        ref = pts.copy()
        if kind == "xyz":
            ref[:, 4] = 0.0
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), kind
        lay = pc2.xyzi_layout(msg)
        assert (lay.offset_x, lay.offset_y, lay.offset_z) == (0, 4, 8)
        assert lay.offset_intensity == {"ouster": 16, "velodyne": 12, "xyz": -1}[kind]
        assert lay.row_step == msg.width * msg.point_step + pad and lay.width * lay.height == pts.shape[0]
    # a field with the right name but another datatype is NOT a match (PCL: "Failed to find match for field")
    msg = pc2.make_pointcloud2(pts, "ouster")
    msg.fields[3] = pc2.PointField("intensity", 16, pc2.UINT32)
    assert pc2.xyzi_layout(msg).offset_intensity == -1
    assert (orc.from_ros_msg(msg)[:, 4] == 0).all()
