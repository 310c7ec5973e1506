"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in direct_lidar_odometry_b200/sharded.py:
pair partitioning, slab sharding with halo, the 43-double all-reduce and the host LM loop.  The per-rank compute
is injected (the CPU oracle stands in for the CUDA backend, which needs a GPU); the GPU twin of this test is
tests/test_gpu_parity.py::test_sharded_partials_sum_to_unsharded."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload():
    from oracle import oracle as O
    from direct_lidar_odometry_b200 import synth
    keys = []
    for j in range(3):
        i = j * 33
        T = synth.trajectory_pose(i)
        s = synth.crop_box_negative(synth.os1_like(i, T, cols=512))
        keys.append(O.voxel_filter(synth.transform_xyzi(O.voxel_filter(s, 0.25), T.astype(np.float32)), 0.5))
    submap = np.ascontiguousarray(np.vstack(keys))
    i = 40
    T = synth.trajectory_pose(i)
    scan = O.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T, cols=512)), 0.4)
    guess = synth.perturb_pose(T, (0.15, 0.05, 0.0), 0.8).astype(np.float32)
    tc = O.Cloud(submap).covariances(10)
    sc = O.Cloud(scan).covariances(10)
    return submap, tc, scan, sc, guess, T


class OracleShardBackend:
    """Test double for CudaShardBackend: same contract, CPU oracle inside."""

    def __init__(self, max_corr_dist):
        from oracle import oracle as O
        self.O, self.thr = O, max_corr_dist
        self.sub = None

    def set_target(self, pts, covs, axis, lo, hi):
        self.tgt, self.tcovs, self.axis, self.lo, self.hi = self.O.Cloud(pts), covs, axis, lo, hi

    def set_source(self, pts, covs):
        self.src, self.scovs = pts, covs

    def _owned(self, T):
        Tf = np.asarray(T, dtype=np.float32)
        p = self.src[:, :3]
        q = (Tf[self.axis, 0] * p[:, 0] + Tf[self.axis, 1] * p[:, 1]) + (Tf[self.axis, 2] * p[:, 2] + Tf[self.axis, 3])
        return (q >= np.float32(max(self.lo, -3e38))) & (q < np.float32(min(self.hi, 3e38)))

    def linearize_partial(self, T):
        m = self._owned(T)
        out = np.zeros(43)
        self.sub = None
        if m.any():
            g = self.O.Gicp(k=10, max_corr_dist=self.thr, num_threads=1)
            g.set_target(self.tgt); g.set_target_covs(self.tcovs)
            g.set_source(self.O.Cloud(self.src[m], build_index=False)); g.set_source_covs(self.scovs[m])
            lin = g.linearize(T)
            out[:36] = lin["H"].T.reshape(36); out[36:42] = lin["b"]; out[42] = lin["err"]
            self.sub = g
        return out

    def compute_error_partial(self, T):
        return self.sub.compute_error(T) if self.sub is not None else 0.0

    # -- min-exchange mode: this rank's nearest neighbour of every source point, then the terms of the points it won
    def set_target_part(self, pts, covs):
        self.tgt_pts, self.tgt, self.tcovs = pts, self.O.Cloud(pts), covs

    def nn1_packed(self, T, rank):
        Tf = np.asarray(T, dtype=np.float32)
        p = self.src[:, :3]
        q = np.stack([(Tf[r, 0] * p[:, 0] + Tf[r, 1] * p[:, 1]) + (Tf[r, 2] * p[:, 2] + Tf[r, 3]) for r in range(3)], axis=1)
        t = self.tgt_pts[:, :3]
        out = np.empty(q.shape[0], dtype=np.uint64)
        for i0 in range(0, q.shape[0], 512):
            d = q[i0:i0 + 512, None, :] - t[None, :, :]                       # float32, nanoflann's metric order
            d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
            best = d2.min(axis=1).astype(np.float32)
            out[i0:i0 + 512] = (best.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.uint64(rank)
        thr2 = np.float64(self.thr) ** 2
        none = np.uint64(0x7f800000ffffffff)
        d2f = (out >> np.uint64(32)).astype(np.uint32).view(np.float32).astype(np.float64)
        out[~(d2f < thr2)] = none
        return out

    def linearize_won(self, T, rank, packed_min):
        m = (packed_min & np.uint64(0xffffffff)) == np.uint64(rank)
        m &= packed_min != np.uint64(0x7f800000ffffffff)
        out = np.zeros(43)
        self.sub = None
        if m.any():
            g = self.O.Gicp(k=10, max_corr_dist=self.thr, num_threads=1)
            g.set_target(self.tgt); g.set_target_covs(self.tcovs)
            g.set_source(self.O.Cloud(self.src[m], build_index=False)); g.set_source_covs(self.scovs[m])
            lin = g.linearize(T)
            out[:36] = lin["H"].T.reshape(36); out[36:42] = lin["b"]; out[42] = lin["err"]
            self.sub = g
        return out


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from direct_lidar_odometry_b200 import sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # pair partition: every pair exactly once
        mine = list(sharded.partition_pairs(10_001, rank, world))
        import torch
        cnt = torch.tensor([len(mine), mine[0], mine[-1]], dtype=torch.int64)
        allc = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allc, cnt)
        assert sum(int(c[0]) for c in allc) == 10_001
        assert all(int(allc[r][2]) + 1 == int(allc[r + 1][1]) for r in range(world - 1))
        # sharded submap
        submap, tc, scan, sc, guess, truth = _workload()
        thr = 0.5
        pts, covs, axis, lo, hi = sharded.shard_target(submap, tc, rank, world, halo=thr + 0.01)
        assert pts.shape[0] < submap.shape[0]
        be = OracleShardBackend(thr)
        be.set_target(pts, covs, axis, lo, hi)
        be.set_source(scan, sc)
        al = sharded.ShardedSubmapAligner(be, max_corr_dist=thr, max_iter=32, trans_eps=0.01)
        res = al.align(guess)
        # unsharded oracle on the whole submap
        from oracle import oracle as O
        g = O.Gicp(k=10, max_corr_dist=thr, max_iter=32, trans_eps=0.01, num_threads=1)
        g.set_target(O.Cloud(submap)); g.set_target_covs(tc)
        g.set_source(O.Cloud(scan)); g.set_source_covs(sc)
        ref = g.align(guess)
        assert (res["nr_iterations"], res["converged"], res["n_linearize"], res["n_compute_error"]) == \
               (ref.nr_iterations, ref.converged, ref.n_linearize, ref.n_compute_error)
        assert np.abs(res["final_x"] - ref.Tx()).max() < 1e-8
        assert np.abs(res["final_x"][:3, 3] - truth[:3, 3]).max() < 0.05
        # one linearisation: the all-reduced partials equal the unsharded sums
        H, b, e = al.linearize(np.asarray(guess, dtype=np.float64))
        lin = g.linearize(np.asarray(guess, dtype=np.float64))
        assert np.abs(H - lin["H"]).max() < 1e-9 * np.abs(lin["H"]).max() and abs(e - lin["err"]) < 1e-9 * abs(lin["err"])
        # ---- unbounded correspondence distance (the library default): no halo is wide enough, the ranks exchange the
        #      nearest neighbours (min-all-reduce of packed (distance, rank) words), then the sums ----
        order = np.argsort(submap[:, axis], kind="stable")
        half = order[:order.size // 2] if rank == 0 else order[order.size // 2:]       # a plain partition: no overlap, no halo
        inf = 1e30
        be2 = OracleShardBackend(inf)
        be2.set_target_part(np.ascontiguousarray(submap[half]), np.ascontiguousarray(tc[half]))
        be2.set_source(scan, sc)
        al2 = sharded.ShardedSubmapAligner(be2, max_corr_dist=inf, max_iter=32, trans_eps=0.01, rank=rank)
        assert al2.exchange == "min"
        res2 = al2.align(guess)
        g2 = O.Gicp(k=10, max_corr_dist=inf, max_iter=32, trans_eps=0.01, num_threads=1)
        g2.set_target(O.Cloud(submap)); g2.set_target_covs(tc)
        g2.set_source(O.Cloud(scan)); g2.set_source_covs(sc)
        ref2 = g2.align(guess)
        assert (res2["nr_iterations"], res2["converged"], res2["n_linearize"], res2["n_compute_error"]) == \
               (ref2.nr_iterations, ref2.converged, ref2.n_linearize, ref2.n_compute_error)
        assert np.abs(res2["final_x"] - ref2.Tx()).max() < 1e-8
        H2, b2, e2 = al2.linearize(np.asarray(guess, dtype=np.float64))
        lin2 = g2.linearize(np.asarray(guess, dtype=np.float64))
        assert np.abs(H2 - lin2["H"]).max() < 1e-9 * np.abs(lin2["H"]).max() and abs(e2 - lin2["err"]) < 1e-9 * abs(lin2["err"])
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_submap_and_pairs():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_partition_pairs_covers_everything():
    from direct_lidar_odometry_b200 import sharded
    for n, w in ((10, 3), (10_000, 8), (5, 8), (0, 2)):
        got = [i for r in range(w) for i in sharded.partition_pairs(n, r, w)]
        assert got == list(range(n))
        sizes = [len(sharded.partition_pairs(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_slab_bounds_balance_target_or_queries():
    """Cut planes: equal target counts by default; with the scan's expected positions and weight 1 equal QUERY counts
    (the per-scan work), a mixture in between."""
    from direct_lidar_odometry_b200 import sharded
    rng = np.random.default_rng(2)
    target = np.zeros((20000, 3), np.float32)
    target[:, 0] = rng.uniform(-300, 300, 20000); target[:, 1] = rng.uniform(-50, 50, 20000)
    queries = np.zeros((4000, 3), np.float32)
    queries[:, 0] = rng.normal(10, 15, 4000)           # crowded around the sensor
    axis, b0 = sharded.slab_bounds(target, 4)
    assert axis == 0 and b0[0] == -np.inf and b0[-1] == np.inf
    cnt = np.histogram(target[:, 0], bins=np.concatenate([[-1e9], b0[1:-1], [1e9]]))[0]
    assert cnt.max() - cnt.min() <= 2
    _, b1 = sharded.slab_bounds(target, 4, queries, 1.0)
    qc = np.histogram(queries[:, 0], bins=np.concatenate([[-1e9], b1[1:-1], [1e9]]))[0]
    assert qc.max() - qc.min() <= 2
    _, bh = sharded.slab_bounds(target, 4, queries, 0.5)
    assert np.all(np.diff(bh[1:-1]) > 0) and abs(bh[2] - b1[2]) < abs(b0[2] - b1[2]) + 1e-9
    qh = np.histogram(queries[:, 0], bins=np.concatenate([[-1e9], bh[1:-1], [1e9]]))[0]
    qz = np.histogram(queries[:, 0], bins=np.concatenate([[-1e9], b0[1:-1], [1e9]]))[0]
    assert qh.max() < qz.max()                          # less crowded than with target-only cuts
    assert sharded.slab_bounds(target, 1)[1].tolist() == [-np.inf, np.inf]
