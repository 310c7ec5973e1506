"""Multi-GPU host logic for the two ways the NanoGICP path shards.

1. Independent scan pairs: `partition_pairs` gives every rank a contiguous block of pairs; no collective.
2. One big submap sharded over ranks (`ShardedSubmapAligner`): the target is cut into slabs along its longest axis
   (equal point counts), every rank holds its slab plus a halo of the max-correspondence distance.  Because a match
   farther than that distance is discarded anyway (reference include/nano_gicp/impl/nano_gicp_impl.hpp:195), the
   rank whose slab contains a transformed source point finds exactly the nearest neighbour the unsharded search
   finds.  Per linearisation every rank contributes the partial {H (36), b (6), err} of the source points it owns,
   the 43 doubles are all-reduced (one tiny collective per phase: NCCL over NVLink on GPUs, gloo in the CPU tests)
   and every rank takes the same Levenberg-Marquardt decision (reference impl/lsq_registration_impl.hpp:160-208)
   from the same numbers.

The per-rank compute is behind a small backend interface; the product backend is `CudaShardBackend` (the C ABI).
Tests may inject another backend to exercise this host logic on machines without a GPU.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib


def partition_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous block partition of `n_pairs` independent registrations; sizes differ by at most one."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def slab_bounds(points: np.ndarray, world: int, queries: np.ndarray | None = None, query_weight: float = 0.0):
    """Axis with the largest extent of the target and `world+1` boundaries; outer bounds are +-inf.

    Without `queries` the cuts give every rank the same number of TARGET points (balances the one-off index +
    covariance build).  The per-scan work, however, follows the SOURCE points: a scan crowds around the sensor, so the
    rank that owns the sensor's slab does most of the correspondence search.  With `queries` (where the scan's points
    are expected to land, e.g. the scan under the initial guess) the cuts are the quantiles of the mixture
    (1 - query_weight) x target distribution + query_weight x query distribution: query_weight = 1 balances the
    align alone, values in between trade it against the build and the memory per rank."""
    xyz = np.asarray(points)[:, :3]
    axis = int(np.argmax(xyz.max(0) - xyz.min(0)))
    if world <= 1:
        return axis, np.array([-np.inf, np.inf])
    t = xyz[:, axis].astype(np.float64)
    levels = np.linspace(0, 1, world + 1)[1:-1]
    if queries is None or query_weight <= 0.0 or len(queries) == 0:
        q = np.quantile(t, levels)
    else:
        s = np.asarray(queries)[:, axis].astype(np.float64)
        s = s[np.isfinite(s)]
        v = np.concatenate([t, s])
        w = np.concatenate([np.full(t.size, (1.0 - query_weight) / t.size), np.full(s.size, query_weight / max(s.size, 1))])
        o = np.argsort(v, kind="stable")
        cw = np.cumsum(w[o])
        q = v[o][np.minimum(np.searchsorted(cw, levels * cw[-1]), v.size - 1)]
    bounds = np.concatenate([[-np.inf], q.astype(np.float32).astype(np.float64), [np.inf]])
    return axis, bounds


def shard_target(points: np.ndarray, covs: np.ndarray, rank: int, world: int, halo: float):
    """This rank's part of the submap: points (and their covariances) inside its slab widened by `halo`."""
    axis, bounds = slab_bounds(points, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    c = np.asarray(points)[:, axis]
    keep = (c >= lo - halo) & (c < hi + halo)
    return np.ascontiguousarray(points[keep]), np.ascontiguousarray(np.asarray(covs)[keep]), axis, float(lo), float(hi)


class CudaShardBackend:
    """One NanoGICP handle on this rank's GPU holding one slab of the target."""

    def __init__(self, device: int, **params):
        from .nanogicp import NanoGICP
        self.g = NanoGICP(device)
        self.g.setCorrespondenceRandomness(params.get("k", 20))
        self.g.setMaxCorrespondenceDistance(params["max_corr_dist"])

    def set_target(self, pts, covs, axis, lo, hi):
        self.g.clearTarget()
        self.g.setInputTarget(pts)
        self.g.setTargetCovariances(covs)
        f32 = np.finfo(np.float32)
        self.g.set_owner_slab(axis, float(np.clip(lo, f32.min, f32.max)), float(np.clip(hi, f32.min, f32.max)))

    def set_source(self, pts, covs=None):
        self.g.clearSource()
        if covs is None:
            self.g.setInputSource(pts)            # index + covariances on this GPU (every rank holds the whole scan)
            self.g.calculateSourceCovariances()
        else:
            self.g.registerInputSource(pts)
            self.g.setSourceCovariances(covs)

    def build_target_shard(self, points: np.ndarray, rank: int, world: int, halo: float, k: int, cov_halo: float = 2.0,
                           chunk: int = 1 << 18, queries: np.ndarray | None = None, query_weight: float = 0.0) -> dict:
        """This rank's slab of `points` with covariances computed HERE, on the slab widened by a halo that is grown
        until the result is exact: a point the rank can be asked to match (inside the slab widened by `halo`, the
        max-correspondence distance) must have its k-th neighbour closer than the nearest cut plane, otherwise its
        neighbourhood may continue on the other side of the cut.  The check reads back the k-th distances of those
        points through the kNN entry of the C ABI, in chunks."""
        from . import _lib
        axis, bounds = slab_bounds(points, world, queries, query_weight)
        lo, hi = float(bounds[rank]), float(bounds[rank + 1])
        c = np.asarray(points)[:, axis].astype(np.float64)
        H, rounds = max(float(cov_halo), float(halo)), 0
        self.g.setCorrespondenceRandomness(k)
        while True:
            rounds += 1
            keep = (c >= lo - H) & (c < hi + H)
            sub = np.ascontiguousarray(points[keep])
            self.g.clearTarget()
            self.g.setInputTarget(sub)
            self.g.calculateTargetCovariances()
            if keep.all():
                break
            cs = c[keep]
            cut_lo = lo - H if np.isfinite(lo) and (c < lo - H).any() else -np.inf
            cut_hi = hi + H if np.isfinite(hi) and (c >= hi + H).any() else np.inf
            margin = np.minimum(cs - cut_lo, cut_hi - cs)
            band = np.nonzero((cs >= lo - halo) & (cs < hi + halo) & np.isfinite(margin))[0]
            worst = 0.0
            bad = 0
            for i in range(0, band.size, chunk):
                sel = band[i:i + chunk]
                _, d2 = self.g.knn(_lib.TARGET, np.ascontiguousarray(sub[sel, :3]), k)
                kth = np.sqrt(d2[:, k - 1].astype(np.float64))
                bad += int((kth >= margin[sel]).sum())
                worst = max(worst, float((kth - margin[sel]).max()))
            if bad == 0:
                break
            H *= 2.0
        f32 = np.finfo(np.float32)
        self.g.set_owner_slab(axis, float(np.clip(lo, f32.min, f32.max)), float(np.clip(hi, f32.min, f32.max)))
        return dict(axis=axis, lo=lo, hi=hi, cov_halo=H, rounds=rounds, shard_points=int(keep.sum()), keep=keep)

    def set_source_sharded(self, pts, rank: int, world: int, group=None):
        """The scan is replicated on every rank (each owns a slab of the TARGET), but its covariances need not be
        computed `world` times: every rank builds the same index (deterministic), computes the covariances of its slice
        of the cell-sorted order (zeros elsewhere) and one NCCL all-reduce(sum) over NVLink — 48 B per point, exact
        because every entry is x + 0 + ... + 0 — gives all ranks the complete, bit-identical set."""
        import torch
        import torch.distributed as dist
        from . import _lib
        self.g.clearSource()
        self.g.setInputSource(pts)
        if world <= 1:
            self.g.calculateSourceCovariances()
            return
        self.g.calculateSourceCovariancesPart(rank, world)
        covs = self.g.covs_device_tensor(_lib.SOURCE)
        stream = torch.cuda.ExternalStream(self.g._L.ngicp_get_stream(self.g._h), device=covs.device)
        with torch.cuda.stream(stream):        # the collective is ordered after the covariance kernels on the handle's stream
            dist.all_reduce(covs, op=dist.ReduceOp.SUM, group=group)

    def linearize_partial(self, T) -> np.ndarray:
        return self.g.linearize_partial(T)

    # -- min-exchange mode (unbounded correspondence distance): the ranks exchange nearest neighbours -------------
    def set_target_part(self, pts, covs):
        """This rank's PART of the target (any partition: no halo, no ownership by position)."""
        self.g.clearTarget()
        self.g.setInputTarget(pts)
        self.g.setTargetCovariances(covs)
        self.g.set_owner_slab(-1)

    def nn1_packed(self, T, rank: int) -> np.ndarray:
        return self.g.nn1_packed(T, rank)

    def linearize_won(self, T, rank: int, packed_min: np.ndarray) -> np.ndarray:
        return self.g.linearize_won(T, rank, packed_min)

    # -- exchange fused into the persistent LM kernel (NVLink peer memory; include/nanogicp_c.h ngicp_comm_*) -----
    def connect_fused(self, rank: int, world: int, group=None):
        """One process per GPU: swap CUDA IPC handles of the exchange buffers over the process group, map them."""
        import torch.distributed as dist
        mine = self.g.comm_export()
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        self.g.comm_connect(rank, world, handles)
        dist.barrier(group=group)

    def set_align_params(self, max_iter: int, trans_eps: float, rot_eps: float = 2e-3):
        self.g.setMaximumIterations(max_iter)
        self.g.setTransformationEpsilon(trans_eps)
        self.g.setRotationEpsilon(rot_eps)

    def align_fused(self, guess=None) -> dict:
        """Every rank calls this with the same guess; the LM loop runs on the GPUs, partial sums meet in peer memory."""
        self.g.align(guess)
        r = self.g.result
        return dict(final_x=self.g.final_state(), final_transformation=self.g.getFinalTransformation(),
                    final_hessian=np.array(r.final_hessian).reshape(6, 6).T.copy(), nr_iterations=int(r.nr_iterations),
                    converged=int(r.converged), n_linearize=int(r.n_linearize), n_compute_error=int(r.n_compute_error),
                    lm_failed=int(r.lm_failed), last_error=float(r.last_error), lm_lambda=float(r.lm_lambda))

    def compute_error_partial(self, T) -> float:
        return float(self.g.compute_error_partial(T)[0])


class ShardedSubmapAligner:
    """Host-stepped LM over a target sharded across the ranks of a torch.distributed process group."""

    def __init__(self, backend, max_corr_dist: float, max_iter: int = 64, trans_eps: float = 5e-4, rot_eps: float = 2e-3,
                 lm_max_iter: int = 10, lm_init_lambda_factor: float = 1e-9, group=None, device=None, exchange: str = "auto",
                 rank: int = 0):
        """exchange: "slab" — ownership by slab + halo of the max-correspondence distance, one sum-all-reduce of 43 doubles per
        linearisation (exact for a FINITE distance); "min" — the ranks hold any partition of the target and exchange the
        nearest neighbours themselves: one min-all-reduce of n_source packed (distance, rank) words, then the sum (exact for
        any distance, the only exact way for the library default FLT_MAX, reference nano_gicp_impl.hpp:59); "auto" takes
        "min" when the distance is unbounded."""
        self.be = backend
        if exchange == "auto":
            exchange = "min" if not np.isfinite(max_corr_dist) or max_corr_dist >= 1e18 else "slab"
        self.exchange, self.rank = exchange, int(rank)
        self.max_corr_dist = max_corr_dist
        self.max_iter, self.trans_eps, self.rot_eps = max_iter, trans_eps, rot_eps
        self.lm_max_iter, self.lm_init_lambda_factor = lm_max_iter, lm_init_lambda_factor
        self.group = group
        self.device = device
        self._L = _lib.load()   # host-only LM helpers; the library loads without a GPU

    # -- collective: sum of a few doubles over ranks -------------------------------------------------
    def _allreduce(self, v: np.ndarray) -> np.ndarray:
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return v
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
        if self.device is not None:
            t = t.to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def _trial(self, H, b, lam, x0):
        dp = C.POINTER(C.c_double)
        Hc = np.ascontiguousarray(H.T).reshape(36)
        x0c = np.ascontiguousarray(x0.T).reshape(16)
        d, delta, xi = np.zeros(6), np.zeros(16), np.zeros(16)
        self._L.ngicp_lm_trial(Hc.ctypes.data_as(dp), np.ascontiguousarray(b).ctypes.data_as(dp), C.c_double(lam),
                               x0c.ctypes.data_as(dp), d.ctypes.data_as(dp), delta.ctypes.data_as(dp), xi.ctypes.data_as(dp))
        return d, delta.reshape(4, 4).T.copy(), xi.reshape(4, 4).T.copy()

    def _converged(self, delta) -> bool:
        dc = np.ascontiguousarray(delta.T).reshape(16)
        return bool(self._L.ngicp_lm_is_converged(dc.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(self.rot_eps),
                                                   C.c_double(self.trans_eps)))

    def _allreduce_min_u64(self, v: np.ndarray) -> np.ndarray:
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return v
        # every packed word is below 2^63 (the distance is a non-negative float): int64 orders them like uint64
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint64).view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t.cpu().numpy().view(np.uint64)

    def linearize(self, T):
        if self.exchange == "min":
            won = self._allreduce_min_u64(self.be.nn1_packed(T, self.rank))
            tot = self._allreduce(np.asarray(self.be.linearize_won(T, self.rank, won), dtype=np.float64))
            return tot[:36].reshape(6, 6).T.copy(), tot[36:42].copy(), float(tot[42])
        tot = self._allreduce(np.asarray(self.be.linearize_partial(T), dtype=np.float64))
        return tot[:36].reshape(6, 6).T.copy(), tot[36:42].copy(), float(tot[42])

    def compute_error(self, T) -> float:
        return float(self._allreduce(np.array([self.be.compute_error_partial(T)], dtype=np.float64))[0])

    def align(self, guess=None) -> dict:
        """LsqRegistration::computeTransformation with step_lm (lsq_registration_impl.hpp:89-115,160-208)."""
        x0 = np.eye(4) if guess is None else np.asarray(guess, dtype=np.float32).astype(np.float64)
        lam, converged, nr_iterations, n_lin, n_err, lm_failed, y0 = -1.0, False, 0, 0, 0, 0, 0.0
        final_H = np.eye(6)
        for it in range(self.max_iter):
            if converged:
                break
            nr_iterations = it
            H, b, y0 = self.linearize(x0)
            n_lin += 1
            if lam < 0.0:
                lam = self.lm_init_lambda_factor * np.abs(np.diag(H)).max()
            nu, ok, delta = 2.0, False, np.eye(4)
            for _ in range(self.lm_max_iter):
                d, delta, xi = self._trial(H, b, lam, x0)
                yi = self.compute_error(xi)
                n_err += 1
                with np.errstate(invalid="ignore", divide="ignore"):
                    rho = (y0 - yi) / float(d @ (lam * d - b))
                if rho < 0:
                    if self._converged(delta):
                        ok = True
                        break
                    lam, nu = nu * lam, 2 * nu
                    continue
                x0 = xi
                w3 = 2.0 * rho - 1.0
                v = 1.0 - w3 * w3 * w3
                lam = lam * (v if 1.0 / 3.0 < v else 1.0 / 3.0)
                final_H = H
                ok = True
                break
            if not ok:
                lm_failed = 1
                break
            converged = self._converged(delta)
        return dict(final_x=x0, final_transformation=x0.astype(np.float32), final_hessian=final_H, nr_iterations=nr_iterations,
                    converged=int(converged), n_linearize=n_lin, n_compute_error=n_err, lm_failed=lm_failed, last_error=y0,
                    lm_lambda=lam)
