"""Host mirror of OdomNode::getSubmapKeyframes: qhull through scipy, PCL's alpha filter restated."""
import itertools

import numpy as np
import pytest

from direct_lidar_odometry_b200 import submap_select as ss

pytest.importorskip("scipy.spatial")


def test_push_submap_indices_kth_rule_and_ties():
    out = []
    ss.push_submap_indices([5.0, 1.0, 3.0, 3.0, 9.0], 3, [10, 11, 12, 13, 14], out)
    assert out == [11, 12, 13]                      # 3rd smallest is 3.0; both 3.0 entries are <= it
    out = []
    ss.push_submap_indices([5.0, 1.0, 3.0, 3.0, 9.0], 2, [10, 11, 12, 13, 14], out)
    assert out == [11, 12, 13]                      # a tie at the k-th distance lets both in
    out = []
    ss.push_submap_indices([2.0, 1.0], 10, [7, 8], out)
    assert out == [7, 8]                            # fewer than k: all
    out = []
    ss.push_submap_indices([], 10, [], out)
    assert out == []


def test_convex_hull_vertices_against_bruteforce():
    rng = np.random.default_rng(5)
    corners = np.array(list(itertools.product([-10.0, 10.0], repeat=3)), dtype=np.float32)
    inner = rng.uniform(-9, 9, size=(40, 3)).astype(np.float32)
    pts = np.vstack([inner[:20], corners, inner[20:]])
    assert ss.convex_hull_vertices(pts) == list(range(20, 28))
    # random cloud: a point is a hull vertex iff some direction has it as the unique maximiser; check both ways
    pts = rng.normal(size=(60, 3)).astype(np.float32)
    hv = set(ss.convex_hull_vertices(pts))
    dirs = rng.normal(size=(20000, 3))
    extreme = set(np.argmax(pts.astype(np.float64) @ dirs.T, axis=0).tolist())
    assert extreme <= hv
    from scipy.optimize import linprog
    for i in range(pts.shape[0]):          # interior points are convex combinations of the others
        others = np.delete(pts.astype(np.float64), i, axis=0)
        res = linprog(np.zeros(others.shape[0]), A_eq=np.vstack([others.T, np.ones(others.shape[0])]),
                      b_eq=np.append(pts[i].astype(np.float64), 1.0), bounds=(0, None), method="highs")
        assert (res.status == 0) == (i not in hv), i
    # flat input: qhull refuses, the hull stays empty (PCL prints an error and returns nothing)
    flat = np.zeros((10, 3), np.float32); flat[:, :2] = rng.normal(size=(10, 2))
    assert ss.convex_hull_vertices(flat) == []
    assert ss.convex_hull_vertices(pts[:3]) == []


def test_concave_hull_alpha_filter():
    rng = np.random.default_rng(9)
    # a dense ball with unit spacing: with alpha ~ 1.2 spacings the alpha shape is its surface layer
    g = np.array(list(itertools.product(range(-4, 5), repeat=3)), dtype=np.float32)
    g = g[(g ** 2).sum(axis=1) <= 16.5] + rng.normal(0, 0.02, size=(g[(g ** 2).sum(axis=1) <= 16.5].shape[0], 3)).astype(np.float32)
    cv = set(ss.concave_hull_vertices(g, 1.2))
    r = np.linalg.norm(g, axis=1)
    assert cv and all(r[i] > 2.5 for i in cv)                 # no deep interior point
    assert sum(1 for i in range(g.shape[0]) if r[i] > 3.7 and i in cv) > 0.8 * (r > 3.7).sum()
    # alpha larger than everything: the alpha shape is the boundary of the whole triangulation = convex hull vertices
    pts = rng.normal(size=(50, 3)).astype(np.float32)
    assert set(ss.concave_hull_vertices(pts, 1e6)) == set(ss.convex_hull_vertices(pts))
    # alpha smaller than any triangle: nothing
    assert ss.concave_hull_vertices(pts, 1e-4) == []
    # two well separated clusters with alpha below the gap: no triangle bridges them, both surfaces are found
    a = rng.normal(size=(40, 3)).astype(np.float32); b = a + np.float32(100.0)
    cv = ss.concave_hull_vertices(np.vstack([a, b]), 3.0)
    assert any(i < 40 for i in cv) and any(i >= 40 for i in cv)


def test_selector_sequence_matches_reference_rules():
    # keyframes along a loop with a small z ripple, as the replay produces them
    t = np.linspace(0, 2 * np.pi, 41)[:-1]
    pos = np.stack([60 * np.cos(t), 40 * np.sin(t), 0.05 * np.sin(7 * t)], axis=1).astype(np.float32)
    sel = ss.SubmapSelector(knn=3, kcv=2, kcc=2, alpha=5.0)
    idx, changed = sel.select(pos[:3], pos[2])
    assert idx == [0, 1, 2] and changed and sel.keyframe_convex == [] and sel.keyframe_concave == []
    idx2, changed2 = sel.select(pos[:3], pos[2] + np.float32(0.01))
    assert idx2 == idx and not changed2
    idx, changed = sel.select(pos, pos[10])
    assert changed and {9, 10, 11} <= set(idx) and idx == sorted(set(idx))
    d = np.linalg.norm(pos - pos[10], axis=1)
    hull = sel.keyframe_convex
    assert len(hull) >= 4
    near_hull = sorted(hull, key=lambda i: d[i])[:2]
    assert set(near_hull) <= set(idx)


# ---- the product implementation (C++ behind the C ABI, csrc/submap_select.cpp) against the qhull-based checker above ----
def test_native_push_indices_matches():
    rng = np.random.default_rng(1)
    for n, k in [(5, 3), (5, 2), (2, 10), (0, 10), (40, 10), (40, 1)]:
        d = np.round(rng.uniform(0, 5, size=n), 1).astype(np.float32)      # rounded: ties at the k-th distance occur
        frames = rng.permutation(100)[:n]
        ref = []
        ss.push_submap_indices(d, k, frames, ref)
        assert ss.native_push_submap_indices(d, k, frames) == ref


def test_native_convex_hull_equals_qhull():
    rng = np.random.default_rng(11)
    for n in (4, 5, 8, 30, 100, 400):
        for trial in range(4):
            pts = (rng.normal(size=(n, 3)) * rng.uniform(1, 50)).astype(np.float32)
            assert ss.native_convex_hull_vertices(pts) == ss.convex_hull_vertices(pts), (n, trial)
    corners = np.array(list(itertools.product([-10.0, 10.0], repeat=3)), dtype=np.float32)
    inner = rng.uniform(-9, 9, size=(40, 3)).astype(np.float32)
    assert ss.native_convex_hull_vertices(np.vstack([inner[:20], corners, inner[20:]])) == list(range(20, 28))
    flat = np.zeros((10, 3), np.float32); flat[:, :2] = rng.normal(size=(10, 2))
    assert ss.native_convex_hull_vertices(flat) == []                       # qhull refuses flat input; PCL returns nothing
    assert ss.native_convex_hull_vertices(flat[:3]) == []


def test_native_concave_hull_equals_qhull_alpha_filter():
    rng = np.random.default_rng(13)
    for n in (5, 12, 40, 120, 300):
        pts = (rng.normal(size=(n, 3)) * 4.0).astype(np.float32)
        for alpha in (0.5, 1.5, 3.0, 8.0, 1e6):
            assert ss.native_concave_hull_vertices(pts, alpha) == ss.concave_hull_vertices(pts, alpha), (n, alpha)
    assert ss.native_concave_hull_vertices(pts, 1e-4) == []
    a = rng.normal(size=(40, 3)).astype(np.float32)
    two = np.vstack([a, a + np.float32(100.0)])
    assert ss.native_concave_hull_vertices(two, 3.0) == ss.concave_hull_vertices(two, 3.0)
    assert ss.native_concave_hull_vertices(pts[:4], 3.0) == []


def test_native_selector_equals_python_selector_on_a_replay_like_loop():
    # keyframes as the C3 replay produces them: a loop with a small z ripple plus pose noise (general position)
    rng = np.random.default_rng(17)
    t = np.linspace(0, 2 * np.pi, 61)[:-1]
    pos = np.stack([60 * np.cos(t), 40 * np.sin(t), 0.05 * np.sin(7 * t)], axis=1) + rng.normal(0, 0.02, size=(60, 3))
    pos = pos.astype(np.float32)
    a, b = ss.SubmapSelector(10, 10, 10, 5.0), ss.NativeSubmapSelector(10, 10, 10, 5.0)
    for nkf in range(1, 61):
        for cur in (pos[nkf - 1], pos[nkf - 1] + np.float32(2.0), pos[max(nkf - 3, 0)]):
            ra, rb = a.select(pos[:nkf], cur), b.select(pos[:nkf], cur)
            assert ra == rb, nkf
        assert a.keyframe_convex == b.keyframe_convex and a.keyframe_concave == b.keyframe_concave, nkf


def test_native_keyframe_decision():
    ident = np.array([[1, 0, 0, 0]], np.float32)
    kf = np.array([[0, 0, 0]], np.float32)
    assert not ss.native_keyframe_wanted(kf, ident, [4.9, 0, 0], ident[0], 5.0, 45.0)
    assert ss.native_keyframe_wanted(kf, ident, [5.1, 0, 0], ident[0], 5.0, 45.0)
    h = np.deg2rad(50.0) / 2
    turned = np.array([np.cos(h), 0, 0, np.sin(h)], np.float32)
    assert ss.native_keyframe_wanted(kf, ident, [1.0, 0, 0], turned, 5.0, 45.0)        # rotated, only one keyframe nearby
    kf2 = np.array([[0, 0, 0], [2, 0, 0]], np.float32)
    assert not ss.native_keyframe_wanted(kf2, np.vstack([ident, ident]), [1.0, 0, 0], turned, 5.0, 45.0)   # two nearby


def test_native_hulls_on_flat_and_nearly_flat_keyframes():
    """Keyframes of a robot on a level floor: exactly flat positions (qhull's 3-D convex hull refuses them: empty; its
    joggled Delaunay triangulates them as slivers: every triangle with a small circumcircle is on the alpha shape) and
    positions with a centimetre ripple over tens of metres."""
    rng = np.random.default_rng(0)
    for trial in range(6):
        flat = np.zeros((12 + 5 * trial, 3), np.float32)
        flat[:, :2] = rng.normal(size=(flat.shape[0], 2)) * 3
        assert ss.native_convex_hull_vertices(flat) == [] == ss.convex_hull_vertices(flat)
        for alpha in (2.0, 5.0):
            assert ss.native_concave_hull_vertices(flat, alpha) == ss.concave_hull_vertices(flat, alpha), (trial, alpha)
    for trial in range(6):
        n = 20 + 7 * trial
        pts = np.zeros((n, 3), np.float32)
        pts[:, :2] = rng.normal(size=(n, 2)) * 30
        pts[:, 2] = rng.normal(size=n) * 0.02
        for alpha in (5.0, 20.0):
            assert ss.native_concave_hull_vertices(pts, alpha) == ss.concave_hull_vertices(pts, alpha), (trial, alpha)
    line = np.zeros((10, 3), np.float32)
    line[:, 0] = np.arange(10)
    assert ss.native_convex_hull_vertices(line) == [] and ss.native_concave_hull_vertices(line, 5.0) == ss.concave_hull_vertices(line, 5.0)
