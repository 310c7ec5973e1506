"""The C++ drop-in facade (include/nano_gicp/nano_gicp.hpp) driven by OdomNode's own call sequence
(tests/cpp/odom_sequence.cpp), compared with the same sequence through the Python mirror and the oracle."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from direct_lidar_odometry_b200 import synth
from util import pose_delta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "odom_sequence")


def test_facade_runs_odom_sequence(tmp_path):
    assert os.path.exists(BIN), "build it with __graft_entry__.build() (make -C tests/cpp)"
    from direct_lidar_odometry_b200 import NanoGICP
    from oracle import oracle as O
    vox = NanoGICP(0)
    scans, poses = [], []
    for i in range(5):
        T = synth.trajectory_pose(i * 4)
        scans.append(vox.voxel_filter(synth.crop_box_negative(synth.os1_like(i * 4, T)), 0.25))
        poses.append(T)
    T0 = poses[0].astype(np.float32)
    path = tmp_path / "scans.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(scans)))
        f.write(np.ascontiguousarray(T0.T).tobytes())
        for s in scans:
            f.write(struct.pack("i", s.shape[0]))
            f.write(np.ascontiguousarray(s).tobytes())
    out = subprocess.run([BIN, str(path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(rows) == 4
    # the additive device-resident keyframe store (nano_gicp::KeyframeStore) must give the same output, bit for bit
    out2 = subprocess.run([BIN, str(path), "store"], capture_output=True, text=True, timeout=120)
    assert out2.returncode == 0, out2.stderr
    assert [l for l in out2.stdout.splitlines() if l.startswith("{")] == [l for l in out.stdout.splitlines() if l.startswith("{")]

    # the same sequence through the Python mirror (same library, same inputs)
    s2s, s2m = NanoGICP(0), NanoGICP(0)
    for obj, (k, thr) in ((s2s, (10, 1.0)), (s2m, (20, 0.5))):
        obj.setCorrespondenceRandomness(k); obj.setMaxCorrespondenceDistance(thr)
        obj.setMaximumIterations(32); obj.setTransformationEpsilon(0.01)
    s2s.setInputTarget(scans[0]); s2s.calculateTargetCovariances()
    key = s2s.voxel_filter(synth.transform_xyzi(scans[0], T0), 0.5)
    s2s.setInputSource(key); s2s.calculateSourceCovariances()
    key_covs = s2s.getSourceCovariances()
    # oracle objects for the tolerance check
    o_s2s = O.Gicp(k=10, max_corr_dist=1.0, max_iter=32, trans_eps=0.01, num_threads=0)
    o_s2m = O.Gicp(k=20, max_corr_dist=0.5, max_iter=32, trans_eps=0.01, num_threads=0)
    o_s2s.set_target(O.Cloud(scans[0])); o_s2s.calc_target_covs()
    ock = O.Cloud(key)
    o_s2s.set_source(ock); o_s2s.calc_source_covs()
    key_covs_ref = o_s2s.get_source_covs()
    prev = T0.copy()
    prev_ref = T0.copy()
    first = True
    for i in range(1, 5):
        s2s.setInputSource(scans[i]); s2m.registerInputSource(scans[i])
        s2m.source_kdtree_ = s2s.source_kdtree_
        s2m.source_covs_.clear()
        s2s.align()
        T_S2S = s2s.getFinalTransformation()
        T_s2s = prev @ T_S2S
        s2m.source_covs_ = s2s.source_covs_
        s2s.swapSourceAndTarget()
        if first:
            s2m.setInputTarget(key); s2m.setTargetCovariances(key_covs); first = False
        s2m.align(T_s2s)
        prev = s2m.getFinalTransformation()
        row = rows[i - 1]
        got_s2s = np.array(row["T_S2S"], dtype=np.float32).reshape(4, 4).T
        got = np.array(row["T"], dtype=np.float32).reshape(4, 4).T
        assert np.array_equal(got_s2s, T_S2S)
        assert np.allclose(got, prev, rtol=0, atol=2e-6)   # the float guess T_s2s_prev * T_S2S is rounded differently by numpy
        assert row["s2s_iterations"] == s2s.result.nr_iterations and row["s2m_iterations"] == s2m.result.nr_iterations
        assert row["aligned_points"] == scans[i].shape[0] and row["s2m_converged"] == 1
        # oracle on the same call sequence
        occ = O.Cloud(scans[i])
        o_s2s.set_source(occ); o_s2m.set_source(occ)
        r = o_s2s.align()
        T_s2s_ref = prev_ref @ r.T()
        o_s2m.set_source_covs(o_s2s.get_source_covs())
        o_s2s.swap()
        if i == 1:
            o_s2m.set_target(ock); o_s2m.set_target_covs(key_covs_ref)
        r2 = o_s2m.align(T_s2s_ref)
        prev_ref = r2.T()
        assert row["s2s_iterations"] == r.nr_iterations and row["s2m_iterations"] == r2.nr_iterations
        dt, dr = pose_delta(got, r2.T())
        assert dt < 5e-4 and dr < 5e-5
        dt, dr = pose_delta(got, poses[i])
        assert dt < 0.1 and dr < 5e-3


def test_reference_text_compiles_and_runs_against_facade(tmp_path):
    """tests/cpp/odom_extract = the reference's OWN bodies of initializeInputTarget / setInputSources / getNextPose /
    updateKeyframes / pushSubmapIndices / getSubmapKeyframes (cut out of src/dlo/odom.cc at build time, never committed)
    compiled verbatim against include/nano_gicp/nano_gicp.hpp.  It must run OdomNode's loop — including new keyframes
    (updateKeyframes: setInputSource(keyframe) + calculateSourceCovariances + getSourceCovariances) and the
    submap_cloud / submap_normals concatenation handed to setInputTarget / setTargetCovariances — and land on the poses
    the oracle gets with the same sequence."""
    exe = os.path.join(ROOT, "tests", "cpp", "odom_extract")
    if not os.path.exists(exe):
        pytest.skip("odom_extract is built only where the reference tree exists (make -C tests/cpp)")
    from direct_lidar_odometry_b200 import NanoGICP
    from oracle import oracle as O
    vox = NanoGICP(0)
    idx = [0, 3, 6, 9, 12, 15, 18, 21, 24]             # 0.45 m apart (inside the S2S correspondence distance); threshD = 1 m below
    scans, poses = [], []
    for i in idx:
        T = synth.trajectory_pose(i)
        scans.append(vox.voxel_filter(synth.crop_box_negative(synth.os1_like(i, T)), 0.25))
        poses.append(T)
    T0 = poses[0].astype(np.float32)
    path = tmp_path / "scans.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(scans)))
        f.write(np.ascontiguousarray(T0.T).tobytes())
        for s in scans:
            f.write(struct.pack("i", s.shape[0]))
            f.write(np.ascontiguousarray(s).tobytes())
    out = subprocess.run([exe, str(path), "1.0"], capture_output=True, text=True, timeout=300)   # keyframe_thresh_dist_ = 1 m
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(rows) == len(scans) - 1
    summary = [(r["scan"], r["keyframes"], r["s2s_iterations"], r["s2m_iterations"], [round(v, 3) for v in r["T"][12:15]]) for r in rows]
    assert rows[-1]["keyframes"] >= 3, summary                 # updateKeyframes added keyframes through the facade
    assert rows[-1]["submap_keyframes"] == rows[-1]["keyframes"]   # every keyframe is in the submap (<= 10 of them)
    assert all(r["submap_points"] == r["submap_normals"] > 0 for r in rows)
    # the oracle through the same sequence (same keyframe rule, every keyframe in the submap)
    import sys
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    import configs as bc
    bc.SELECTION = "knn"
    rc = bc.Replay(lambda cfg: bc.OracleGicp(O, cfg, os.cpu_count()), lambda p, l: O.voxel_filter(p, l), None, thresh_d=1.0, knn=10)
    bc.SELECTION = "hull"
    for i, s in enumerate(scans):
        if i == 0:
            rc.first(s, poses[0])
            continue
        its = rc.step(s)
        row = rows[i - 1]
        got = np.array(row["T"], dtype=np.float32).reshape(4, 4).T
        assert (row["s2s_iterations"], row["s2m_iterations"]) == tuple(int(v) for v in its)
        dt, dr = pose_delta(got, rc.T)
        assert dt < 5e-4 and dr < 5e-5, (i, dt, dr)
        assert row["keyframes"] == len(rc.keyframes)
        dt, dr = pose_delta(got, poses[i])
        assert dt < 0.1 and dr < 5e-3
