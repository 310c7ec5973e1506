"""Generates tests/golden/c5_oracle.json: the CPU oracle's answer for BASELINE config C5 (dense 128-beam scan, no voxel
filter, registered against the 5 M-point submap of dense keyframes; DLO's S2M parameters), so that the sharded and the
unsharded GPU registrations of bench.py / benchmarks/configs.py c5 can be compared with the reference path on identical
inputs without re-running a 5 M-point kd-tree on every box.  The inputs come from benchmarks/configs.c5_workload (numpy
ray-caster, deterministic); their SHA-1 is stored so that a consumer can tell whether it regenerated the same bytes.

    python tests/golden/make_c5_golden.py            # ~2 min on 8 cores (oracle/_ref when built, else the restated tree)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import configs as bc  # noqa: E402
from oracle import oracle as O  # noqa: E402


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def main():
    n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    L = O.load(prefer_ref=True)
    wl = bc.c5_workload(n_target, lambda *a: print(*a, file=sys.stderr))
    target, source, guess = wl["target"], wl["source"], wl["guess"]
    cfg = bc.S2M
    t0 = time.time()
    g = O.Gicp(k=cfg["k"], max_corr_dist=cfg["thr"], max_iter=cfg["max_iter"], trans_eps=cfg["trans_eps"], num_threads=os.cpu_count())
    g.set_target(O.Cloud(target)); g.set_source(O.Cloud(source))
    r = g.align(guess)
    out = {"config": "C5 oracle: dense scan vs dense-keyframe submap, S2M parameters", "params": cfg,
           "target_points": int(target.shape[0]), "source_points": int(source.shape[0]),
           "target_sha1": sha(target), "source_sha1": sha(source), "guess": np.asarray(guess, dtype=np.float64).reshape(-1).tolist(),
           "final_x": r.Tx().reshape(-1).tolist(), "nr_iterations": int(r.nr_iterations), "n_linearize": int(r.n_linearize),
           "n_compute_error": int(r.n_compute_error), "converged": int(r.converged),
           "oracle": "reference nanoflann + restated GICP math" if L.orc_has_ref_nanoflann() else "restated kd-tree + restated GICP math",
           "oracle_seconds": time.time() - t0}
    dst = os.path.join(ROOT, "tests", "golden", f"c5_oracle_{n_target}.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("target_points", "source_points", "nr_iterations", "n_compute_error", "oracle_seconds")}))


if __name__ == "__main__":
    main()
