"""Two NanoGICP handles on one GPU run the sharded align with the exchange fused into their persistent kernels
(tests/test_gpu_parity.py::test_sharded_align_fused_exchange_one_gpu runs this in a fresh process).  Prints one JSON line."""
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from direct_lidar_odometry_b200 import NanoGICP, NanoGICPError, sharded, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (test infrastructure: CPU covariances as common inputs)
from util import make_small_submap  # noqa: E402


def main():
    O.load(prefer_ref=True)
    submap, scan, T = make_small_submap(O)
    thr = 0.5
    tc = O.Cloud(submap).covariances(20)
    sc = O.Cloud(scan).covariances(20)
    guess = synth.perturb_pose(T, (0.2, 0.0, 0.0), 1.0).astype(np.float32)
    world = 2
    backs = []
    for r in range(world):
        pts, covs, axis, lo, hi = sharded.shard_target(submap, tc, r, world, halo=thr + 0.01)
        be = sharded.CudaShardBackend(0, k=20, max_corr_dist=thr)
        be.set_align_params(32, 0.01)
        be.set_target(pts, covs, axis, lo, hi)
        be.set_source(scan, sc)
        backs.append(be)
    for r, be in enumerate(backs):
        be.g.comm_connect_local(r, [b.g for b in backs])
    out, err = [None] * world, [None] * world

    def run(r):
        try:
            out[r] = backs[r].align_fused(guess)
        except Exception as e:  # noqa: BLE001
            err[r] = e

    identical = True
    for rep in range(3):   # the exchange counters carry over from one align to the next
        th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        if err != [None, None]:
            raise SystemExit(f"sharded align failed: {err}")
        identical = identical and bool(np.array_equal(out[0]["final_x"], out[1]["final_x"])) and \
            out[0]["nr_iterations"] == out[1]["nr_iterations"]
    g = NanoGICP(0)
    g.setCorrespondenceRandomness(20); g.setMaxCorrespondenceDistance(thr)
    g.setMaximumIterations(32); g.setTransformationEpsilon(0.01)
    g.setInputTarget(submap); g.setTargetCovariances(tc)
    g.setInputSource(scan); g.setSourceCovariances(sc)
    g.align(guess)
    res = out[0]
    counts = (res["nr_iterations"], res["n_linearize"], res["n_compute_error"], res["converged"]) == \
             (g.result.nr_iterations, g.result.n_linearize, g.result.n_compute_error, g.result.converged)
    dT = float(np.abs(res["final_x"] - g.final_state()).max())
    # a rank that never shows up must not hang the GPU: the waiting rank gets NGICP_E_COMM after the timeout
    os.environ["NGICP_COMM_TIMEOUT_MS"] = "200"
    for r, be in enumerate(backs):
        be.g.comm_connect_local(r, [b.g for b in backs])
    timed_out = False
    try:
        backs[0].align_fused(guess)
    except NanoGICPError as e:
        timed_out = "peer rank" in str(e)
    # recovery: the failed exchange left the sequence numbers out of step and the error flag set; after every rank's
    # ngicp_comm_reset (connections kept) the next sharded align works again and gives the same bits as before
    os.environ["NGICP_COMM_TIMEOUT_MS"] = "2000"
    for be in backs:
        be.g.comm_reset()
    out2, err2 = [None] * world, [None] * world

    def run2(r):
        try:
            out2[r] = backs[r].align_fused(guess)
        except Exception as e:  # noqa: BLE001
            err2[r] = e
    th = [threading.Thread(target=run2, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    recovered = err2 == [None, None] and bool(np.array_equal(out2[0]["final_x"], res["final_x"])) and \
        bool(np.array_equal(out2[1]["final_x"], res["final_x"]))
    for be in backs:
        be.g.comm_close()
    print(json.dumps({"ranks_bit_identical": identical, "counts_equal_unsharded": bool(counts), "max_abs_dT_vs_unsharded": dT,
                      "timeout_reported": timed_out, "recovered_after_reset": bool(recovered), "errors_after_reset": [str(e) for e in err2 if e]}))


if __name__ == "__main__":
    main()
