"""Generates tests/golden/knn_ref_nanoflann.npz from the REFERENCE's own kd-tree.

Run here (container with /root/reference) after `make -C oracle`:
    python tests/golden/make_golden.py
The kNN answers come from oracle/_ref/liboracle_ref.so, i.e. the reference's vendored
include/nano_gicp/impl/nanoflann_impl.hpp compiled unmodified and instantiated exactly like
include/nano_gicp/nanoflann.hpp:100-117 (SO3_Adaptor<float>, DIM 3, int index, leaf 100).
Inputs are stored in the fixture so that it does not depend on numpy's RNG stream.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from direct_lidar_odometry_b200 import synth  # noqa: E402


def main():
    L = orc.load(prefer_ref=True)
    assert L.orc_has_ref_nanoflann() == 1, "needs the _ref build (reference nanoflann)"
    rng = np.random.default_rng(20261018)
    planes = synth.random_planes_cloud(2500, seed=3)[:, :3]
    # a voxelised synthetic scan slice, plus duplicated points and a tight cluster to exercise ties
    scan = synth.crop_box_negative(synth.os1_like(5, synth.trajectory_pose(5)))
    vox = orc.voxel_filter(scan, 0.5, lib=L)[:, :3]
    vox = vox[rng.permutation(vox.shape[0])[:2500]]
    lattice = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(3), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.5 + 30.0
    cloud = np.vstack([planes, vox, lattice, planes[:20]]).astype(np.float32)
    queries = np.vstack([cloud[rng.permutation(cloud.shape[0])[:400]],
                         rng.uniform(-60, 60, size=(200, 3)).astype(np.float32) * np.float32([1, 1, 0.2]),
                         lattice[:30] + np.float32(0.25)]).astype(np.float32)
    c = orc.Cloud(cloud, orc.BACKEND_REF, lib=L)
    out = dict(cloud=cloud, queries=queries)
    for k in (1, 5, 10, 20):
        idx, d2 = c.knn(queries, k, nthreads=1)
        out[f"idx_k{k}"] = idx
        out[f"d2_k{k}"] = d2
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "knn_ref_nanoflann.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
