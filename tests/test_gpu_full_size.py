"""GPU parity at BASELINE sizes and on the production kernels (VERDICT r1 "parity holes"):

  * the TILE kNN path (what every cloud >= 131072 points takes, i.e. the bench's dominant kernel) against the oracle's
    kd-tree — neighbour lists, neighbour sets and covariances — on a 170k-point voxelised submap and on the bench's own
    500k-point C2 submap; the WARP path on the same clouds; both paths against each other bit for bit;
  * covariances BIT-IDENTICAL to the oracle wherever the k+1 nearest distances are distinct (summation in ascending
    distance order, no FMA contraction — reference nano_gicp_impl.hpp:315-321, CMakeLists.txt:13-14);
  * the full C2 bench workload registered by the GPU and by the oracle: same iteration / linearize / compute_error
    counts, pose within 1e-4 m / 1e-5 rad;
  * two handles computing the covariances of two large clouds concurrently (the round-1 tile kernel could deadlock);
  * C3: OdomNode's replay over 300 scans, every align compared with the oracle on identical inputs: identical counts;
  * C4: a seeded sample of 200 distinct scan pairs of the 10 001-pose trajectory against the oracle.
"""
import os
import sys
import threading

import numpy as np
import pytest

from direct_lidar_odometry_b200 import synth, _lib
from util import tie_free_mask, pose_delta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COV_RTOL = 1e-5
POSE_T_TOL = 1e-4
POSE_R_TOL = 1e-5
S2S = dict(k=10, thr=1.0, max_iter=32, trans_eps=0.01)      # cfg/params.yaml:54-58
S2M = dict(k=20, thr=0.5, max_iter=32, trans_eps=0.01)      # cfg/params.yaml:63-67


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


def configure(g, cfg):
    g.setCorrespondenceRandomness(cfg["k"]); g.setMaxCorrespondenceDistance(cfg["thr"])
    g.setMaximumIterations(cfg["max_iter"]); g.setTransformationEpsilon(cfg["trans_eps"])


@pytest.fixture(scope="module")
def c2_workload(G):
    """bench.py's own C2 inputs (500 000-point submap, ~22k-point scans, guesses)."""
    sys.path.insert(0, ROOT)
    import bench
    g = G()
    return bench.make_workload(lambda p, leaf: g.voxel_filter(p, leaf))


@pytest.fixture(scope="module")
def submap_170k(c2_workload):
    """The first keyframes of that submap: a voxelised world-frame submap above the 131072-point switch."""
    return np.ascontiguousarray(c2_workload["submap"][:170_000])


def upper(c):
    """(n, 6) upper triangle of (n, 4, 4) covariance records."""
    return np.stack([c[:, 0, 0], c[:, 0, 1], c[:, 0, 2], c[:, 1, 1], c[:, 1, 2], c[:, 2, 2]], axis=1)


def oracle_covs(O, cloud, k):
    tree = O.Cloud(cloud)
    ref, ridx, rd2 = tree.covariances(k, with_knn=True)
    _, d2p = tree.knn(np.ascontiguousarray(cloud[:, :3]), k + 1)
    raw = tree.covariances(k, method=0)[:, :3, :3]
    return dict(ref=ref, idx=ridx, d2=rd2, d2p=d2p, raw=raw)


def check_covs_against_oracle(g, cloud, k, path, oc):
    """One kNN path against the oracle on one cloud; returns the GPU covariances (upper triangles) for cross-path checks."""
    g.setKnnPath(path)
    g.setCorrespondenceRandomness(k)
    g.clearTarget()
    g.setInputTarget(cloud)
    assert g.calculateTargetCovariances() is True
    got = g.getTargetCovariances()
    nidx, nd2 = g.cov_neighbors(_lib.TARGET)
    ref, ridx, rd2, d2p = oc["ref"], oc["idx"], oc["d2"], oc["d2p"]
    n = cloud.shape[0]
    # 1. the sorted squared distances of the k neighbours the covariance was summed over: bit-identical to nanoflann's
    assert np.array_equal(nd2.view(np.uint32), rd2.view(np.uint32))
    # 2. neighbour SETS identical wherever the k-th and (k+1)-th distances differ (north star: "excluding exact distance ties")
    no_boundary_tie = d2p[:, k - 1] != d2p[:, k]
    assert no_boundary_tie.mean() > 0.99
    assert np.array_equal(np.sort(nidx[no_boundary_tie], axis=1), np.sort(ridx[no_boundary_tie], axis=1))
    # 3. neighbour ORDER identical on tie-free slots
    m = tie_free_mask(d2p)
    assert np.array_equal(nidx[m], ridx[m])
    # 4. covariances within 1e-5 relative wherever the PLANE regularisation is resolvable (see test_covariances_vs_oracle)
    num = np.linalg.norm((got - ref).reshape(n, 16), axis=1)
    den = np.linalg.norm(ref.reshape(n, 16), axis=1)
    w = np.linalg.eigvalsh(oc["raw"])
    resolvable = no_boundary_tie & ((w[:, 1] - w[:, 0]) > 1e-6 * w[:, 2])
    assert resolvable.mean() > 0.98
    assert (num / den)[resolvable].max() < COV_RTOL
    ev = np.linalg.eigvalsh(0.5 * (got[:, :3, :3] + got[:, :3, :3].transpose(0, 2, 1)))
    assert np.allclose(ev, [1e-3, 1, 1], atol=1e-8)
    # 5. BIT-identical covariances wherever all k+1 nearest distances are distinct: same neighbours, same summation
    #    order, same unfused fp64 operations as the oracle
    distinct = (np.diff(d2p, axis=1) > 0).all(axis=1)
    assert distinct.mean() > 0.9
    gu, ru = upper(got), upper(ref)
    same = (gu.view(np.uint64) == ru.view(np.uint64)).all(axis=1)
    bad = int((distinct & ~same).sum())
    assert bad == 0, f"{bad} of {int(distinct.sum())} tie-free points differ from the oracle in the last bits (max {np.abs(gu - ru)[distinct].max():.3e})"
    return gu


@pytest.mark.parametrize("k", [20, 10])
def test_tile_and_warp_knn_covariances_vs_oracle_170k(G, O, submap_170k, k):
    g = G()
    oc = oracle_covs(O, submap_170k, k)
    tile = check_covs_against_oracle(g, submap_170k, k, _lib.KNN_TILE, oc)
    warp = check_covs_against_oracle(g, submap_170k, k, _lib.KNN_WARP, oc)
    # the two kernel families agree bit for bit on every point without a distance tie among its k+1 nearest
    distinct = (np.diff(oc["d2p"], axis=1) > 0).all(axis=1)
    assert np.array_equal(tile[distinct].view(np.uint64), warp[distinct].view(np.uint64))
    # AUTO picks the tiles at this size: same bits as the forced tile path everywhere
    g.setKnnPath(_lib.KNN_AUTO)
    g.clearTarget(); g.setInputTarget(submap_170k); g.calculateTargetCovariances()
    assert np.array_equal(upper(g.getTargetCovariances()).view(np.uint64), tile.view(np.uint64))


def test_tile_knn_covariances_vs_oracle_c2_submap_500k(G, O, c2_workload):
    """The bench's own dominant launch: k=20 covariances over the 500 000-point C2 submap, default (tile) path."""
    submap = c2_workload["submap"]
    assert submap.shape[0] == 500_000
    oc = oracle_covs(O, submap, 20)
    check_covs_against_oracle(G(), submap, 20, _lib.KNN_AUTO, oc)


def test_scan_covariances_bit_identical_to_oracle(G, O, scan_pair):
    """Voxelised scans (warp path, DLO's S2S k=10 and S2M k=20)."""
    g = G()
    v0 = O.voxel_filter(scan_pair["s0"], 0.25)
    for k in (10, 20):
        oc = oracle_covs(O, v0, k)
        check_covs_against_oracle(g, v0, k, _lib.KNN_AUTO, oc)


def test_c2_full_size_align_parity(G, O, c2_workload):
    """bench.py's step on the GPU and on the oracle: identical counts, pose within the north-star tolerance."""
    submap = c2_workload["submap"]
    for r in (0, 3):
        scan, guess = c2_workload[f"scan_{r}"], c2_workload["guesses"][r]
        g = G()
        configure(g, S2M)
        g.setInputTarget(submap); g.calculateTargetCovariances()
        g.setInputSource(scan); g.calculateSourceCovariances()
        g.align(guess)
        o = O.Gicp(k=S2M["k"], max_corr_dist=S2M["thr"], max_iter=S2M["max_iter"], trans_eps=S2M["trans_eps"], num_threads=0)
        o.set_target(O.Cloud(submap)); o.set_source(O.Cloud(scan))
        ro = o.align(guess)
        res = g.result
        assert (res.nr_iterations, res.converged, res.n_linearize, res.n_compute_error, res.lm_failed) == \
               (ro.nr_iterations, ro.converged, ro.n_linearize, ro.n_compute_error, ro.lm_failed)
        dt, dr = pose_delta(g.final_state(), ro.Tx())
        assert dt < POSE_T_TOL and dr < POSE_R_TOL, (dt, dr)
        dt, dr = pose_delta(g.final_state(), c2_workload["truths"][r])
        assert dt < 0.05 and dr < 2e-3


@pytest.mark.timeout(600)
def test_concurrent_large_covariances_two_handles(G):
    """Two handles, two host threads, two streams: covariances of two clouds above the tile switch at the same time,
    repeatedly — completes and reproduces the serial bits (round 1's persistent tile kernel waited for all of its own
    warps and could deadlock when two such launches shared the GPU)."""
    clouds = [synth.random_planes_cloud(200_000, seed=21, extent=80.0, noise=0.02),
              synth.random_planes_cloud(260_000, seed=22, extent=100.0, noise=0.02)]
    serial = []
    for c in clouds:
        g = G()
        g.setKnnPath(_lib.KNN_TILE)
        g.setInputTarget(c); g.calculateTargetCovariances()
        serial.append(upper(g.getTargetCovariances()))
    handles = [G(), G()]
    for h in handles:
        h.setKnnPath(_lib.KNN_TILE)
    start = threading.Barrier(2)
    out, errs = [None, None], []

    def work(i):
        try:
            h, c = handles[i], clouds[i]
            start.wait()
            for _ in range(6):
                h.clearTarget()
                h.setInputTarget(c)
                h.calculateTargetCovariances()
            out[i] = upper(h.getTargetCovariances())
        except Exception as e:  # pragma: no cover
            errs.append(e)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in ts), "concurrent covariance launches did not finish"
    assert not errs, errs
    for i in range(2):
        assert np.array_equal(out[i].view(np.uint64), serial[i].view(np.uint64))


@pytest.mark.timeout(1800)
def test_c3_replay_identical_counts_300_scans(G, O):
    """OdomNode's per-scan sequence (S2S align, submap selection, S2M align, keyframe update) over 300 scans on the
    GPU and on the oracle; after every scan the GPU replay continues from the oracle's pose so that each of the 598
    aligns is compared on identical inputs.  Identical iteration counts everywhere, poses within tolerance."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import configs as bc
    n = 300
    scans = bc.gen_scans(list(range(n)))
    vox = G()

    def make_gpu(cfg):
        g = G()
        configure(g, cfg)
        return g
    threads = os.cpu_count()
    rc = bc.Replay(lambda cfg: bc.OracleGicp(O, cfg, threads), lambda p, l: O.voxel_filter(p, l), None)
    rg = bc.Replay(make_gpu, lambda p, l: vox.voxel_filter(p, l), None)
    mism, worst = [], (0.0, 0.0)
    for i in range(n):
        T_true, raw = scans[i]
        sc = O.voxel_filter(raw, 0.25)
        sg = vox.voxel_filter(raw, 0.25)
        assert np.array_equal(sc.view(np.uint32), sg.view(np.uint32))
        if i == 0:
            rc.first(sc, T_true); rg.first(sg, T_true)
            continue
        ic = rc.step(sc)
        ig = rg.step(sg, force_T=rc.T)
        if tuple(ic) != tuple(ig):
            mism.append((i, tuple(ic), tuple(ig)))
        dt, dr = bc.pose_err(rc.T, rg.T_result)
        worst = (max(worst[0], dt), max(worst[1], dr))
    assert not mism, f"iteration counts differ on {len(mism)} of {n - 1} scans: {mism[:5]}"
    assert worst[0] < POSE_T_TOL and worst[1] < POSE_R_TOL, worst
    assert len(rg.keyframes) == len(rc.keyframes) and len(rg.keyframes) >= 5


@pytest.mark.timeout(1800)
def test_c4_distinct_pairs_sample_vs_oracle(G, O):
    """C4: pair i registers scan i+1 against scan i of the 10 001-pose trajectory.  A seeded sample
    of 200 of the 10 000 distinct pairs against the oracle: identical counts, pose within tolerance."""
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import configs as bc
    rng = np.random.default_rng(2026)
    pairs = np.sort(rng.choice(10_000, size=200, replace=False))
    need = sorted(set(pairs.tolist()) | set((pairs + 1).tolist()))
    scans = bc.gen_scans(need)
    g = G()
    configure(g, S2S)
    threads = os.cpu_count()
    bad = []
    for i in pairs.tolist():
        a = g.voxel_filter(scans[i][1], 0.25)
        b = g.voxel_filter(scans[i + 1][1], 0.25)
        g.clearSource(); g.clearTarget()
        g.setInputTarget(a); g.setInputSource(b)
        g.align()
        o = O.Gicp(k=S2S["k"], max_corr_dist=S2S["thr"], max_iter=S2S["max_iter"], trans_eps=S2S["trans_eps"], num_threads=threads)
        o.set_target(O.Cloud(a)); o.set_source(O.Cloud(b))
        r = o.align()
        res = g.result
        dt, dr = pose_delta(g.final_state(), r.Tx())
        ok = (res.nr_iterations, res.converged, res.n_linearize, res.n_compute_error) == \
             (r.nr_iterations, r.converged, r.n_linearize, r.n_compute_error) and dt < POSE_T_TOL and dr < POSE_R_TOL
        if not ok:
            bad.append((i, res.nr_iterations, r.nr_iterations, dt, dr))
    assert not bad, f"{len(bad)} of {len(pairs)} pairs differ: {bad[:5]}"


def test_align_batch_bit_identical_to_single_calls(G, O, scan_pair):
    """ngicp_align_batch: B pairs in one launch (one thread-block cluster per pair) == B single ngicp_align calls, bit
    for bit (state, Hessian, counts), for scan-to-scan pairs (one lane per point), a scan-to-map pair (two lanes per
    point), different parameters per handle, a guess, and a pair without any correspondence."""
    from direct_lidar_odometry_b200 import align_batch
    from util import make_small_submap
    scans = []
    for i in (0, 1, 40, 41, 500, 501):
        T = synth.trajectory_pose(i)
        scans.append(O.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T)), 0.25))
    submap, scan_m, Tm = make_small_submap(O)
    far = scans[0].copy(); far[:, :3] += 500.0
    cases = [  # (source, target, config, guess)
        (scans[1], scans[0], S2S, None),
        (scans[3], scans[2], S2S, synth.perturb_pose(np.eye(4), (0.1, 0.05, 0.0), 0.5).astype(np.float32)),
        (scans[5], scans[4], dict(k=20, thr=float(np.finfo(np.float32).max), max_iter=64, trans_eps=5e-4), None),
        (scan_m, submap, S2M, synth.perturb_pose(Tm, (0.2, 0.0, 0.0), 1.0).astype(np.float32)),
        (far, scans[0], S2S, None),
        (scans[0], scans[1], S2S, None),
    ]

    def setup(case):
        src, tgt, cfg, _ = case
        g = G()
        configure(g, cfg)
        g.setInputTarget(tgt); g.setInputSource(src)
        return g
    single = [setup(c) for c in cases]
    for g, c in zip(single, cases):
        g.align(c[3])
    batch = [setup(c) for c in cases]
    res = align_batch(batch, [c[3] for c in cases])
    assert len(res) == len(cases)
    for gs, gb in zip(single, batch):
        a, b = gs.result, gb.result
        assert (a.nr_iterations, a.converged, a.n_linearize, a.n_compute_error, a.lm_failed) == \
               (b.nr_iterations, b.converged, b.n_linearize, b.n_compute_error, b.lm_failed)
        assert np.array_equal(gs.final_state().view(np.uint64), gb.final_state().view(np.uint64))
        assert np.array_equal(gs.getFinalHessian().view(np.uint64), gb.getFinalHessian().view(np.uint64))
        assert np.array_equal(gs.getFinalTransformation(), gb.getFinalTransformation()) and gs.hasConverged() == gb.hasConverged()
    assert single[0].result.nr_iterations >= 1 and single[4].result.nr_iterations == 0
    # a second batch on the same handles (covariances now present, other guesses) still matches fresh single calls
    g2 = [synth.perturb_pose(np.eye(4), (0.05, 0.0, 0.0), 0.2).astype(np.float32)] * 2
    align_batch(batch[:2], g2)
    for gs, gb, gg in zip(single[:2], batch[:2], g2):
        gs.align(gg)
        assert np.array_equal(gs.final_state().view(np.uint64), gb.final_state().view(np.uint64))
    # error paths: the same handle twice, a handle without target
    with pytest.raises(Exception):
        align_batch([batch[0], batch[0]])
    with pytest.raises(Exception):
        align_batch([batch[0], G()])
