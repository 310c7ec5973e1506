"""Host mirror of OdomNode::getSubmapKeyframes (reference src/dlo/odom.cc:1240-1293, pushSubmapIndices :1210-1233,
computeConvexHull :1017-1050, computeConcaveHull :1057-1090) — SURVEY section 8f row N3.

Which keyframes form the scan-to-map target: the `knn` keyframes nearest to the current pose, the `kcv` nearest among
the vertices of the 3-D convex hull of all keyframe positions, and the `kcc` nearest among the vertices of their 3-D
alpha shape ("concave hull", alpha = keyframe threshD at construction, odom.cc:95-98).  A few dozen points per call:
host work in the reference and here.  The reference gets both hulls from PCL, which calls qhull; this mirror calls
the same library through scipy.spatial (convex hull: qhull's default options as pcl::ConvexHull's 3-D path; Delaunay
tetrahedra with "QJ" as pcl::ConcaveHull's "qhull d QJ") and restates PCL's alpha filter on top of it.  It is used by
the replay benchmark and the tests; a DLO build keeps OdomNode's own code for this step (only NanoGICP is replaced).
"""
from __future__ import annotations

import numpy as np


def push_submap_indices(dists, k: int, frames, out: list) -> None:
    """pushSubmapIndices (odom.cc:1210-1233): every frame whose distance is <= the k-th smallest distance (so ties at
    the k-th distance all get in); with fewer than k candidates all of them; with none nothing (the reference reads
    the top of an empty heap there, but then loops over zero elements)."""
    d = np.asarray(dists, dtype=np.float32)
    if d.size == 0:
        return
    kth = np.sort(d)[min(k, d.size) - 1] if k > 0 else np.float32(-np.inf)
    for i in range(d.size):
        if d[i] <= kth:
            out.append(int(frames[i]))


def convex_hull_vertices(pos: np.ndarray) -> list:
    """pcl::ConvexHull<PointXYZI> with setDimension(3) + getHullPointIndices: the input points that are vertices of
    the 3-D convex hull.  qhull refuses flat input (all keyframes coplanar); PCL then reports an error and the hull
    stays empty."""
    from scipy.spatial import ConvexHull, QhullError
    p = np.asarray(pos, dtype=np.float32).astype(np.float64)
    if p.shape[0] < 4:
        return []
    try:
        return sorted(int(v) for v in ConvexHull(p).vertices)
    except (QhullError, ValueError):
        return []


def _circumcircle_radius(a, b, c) -> float:
    """pcl::getCircumcircleRadius: Heron's formula on the three side lengths (float vectors, double arithmetic)."""
    a, b, c = (np.asarray(v, dtype=np.float32) for v in (a, b, c))
    l1 = float(np.linalg.norm(b - a)); l2 = float(np.linalg.norm(c - b)); l3 = float(np.linalg.norm(a - c))
    s = (l1 + l2 + l3) / 2.0
    area2 = s * (s - l1) * (s - l2) * (s - l3)
    area = np.sqrt(area2) if area2 > 0 else 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(l1 * l2 * l3) / np.float64(4.0 * area))


def _circumsphere_radius(t: np.ndarray) -> float:
    """distance from a vertex of the tetrahedron to its Voronoi centre (qh_pointdist(vertex, facet->center))"""
    A = 2.0 * (t[1:] - t[0])
    b = (t[1:] ** 2).sum(axis=1) - (t[0] ** 2).sum()
    try:
        c = np.linalg.solve(A, b)
    except np.linalg.LinAlgError:
        return float("inf")
    return float(np.linalg.norm(c - t[0]))


def concave_hull_vertices(pos: np.ndarray, alpha: float) -> list:
    """pcl::ConcaveHull<PointXYZI> with setDimension(3), setAlpha(alpha), setKeepInformation(true) +
    getHullPointIndices (PCL surface/concave_hull.hpp, 3-D branch): Delaunay tetrahedra; a tetrahedron is good when
    its circumsphere radius <= alpha; candidate triangles are the faces of good tetrahedra plus any face whose own
    circumcircle radius <= alpha; a candidate is part of the alpha shape unless both tetrahedra behind it are good
    (faces on the outside of the triangulation count as having a not-good neighbour).  The result is the set of input
    points used by those triangles."""
    from scipy.spatial import Delaunay, QhullError
    pf = np.asarray(pos, dtype=np.float32)
    p = pf.astype(np.float64)
    if p.shape[0] < 5:
        return []
    try:
        tri = Delaunay(p, qhull_options="QJ")
    except (QhullError, ValueError):
        return []
    tets, nbrs = tri.simplices, tri.neighbors
    good = np.array([_circumsphere_radius(p[t]) <= alpha for t in tets], dtype=bool)
    used = set()
    for ti, t in enumerate(tets):
        for f in range(4):
            nb = int(nbrs[ti, f])
            if nb >= 0 and nb < ti:
                continue                                   # each inner face once
            face = [int(t[j]) for j in range(4) if j != f]
            nb_good = nb >= 0 and bool(good[nb])
            candidate = bool(good[ti]) or nb_good or _circumcircle_radius(pf[face[0]], pf[face[1]], pf[face[2]]) <= alpha
            boundary = nb < 0 or not good[ti] or not nb_good
            if candidate and boundary:
                used.update(face)
    return sorted(used)


class SubmapSelector:
    """State OdomNode keeps between scans for this step: keyframe_convex / keyframe_concave survive calls in which
    the hull is not recomputed (fewer than 4 / 5 keyframes), submap_kf_idx_prev decides submap_hasChanged."""

    def __init__(self, knn: int = 10, kcv: int = 10, kcc: int = 10, alpha: float = 5.0):
        # defaults: cfg/params.yaml submap.keyframe.{knn,kcv,kcc} = 10, keyframe.threshD = 5.0 (alpha, odom.cc:97)
        self.knn, self.kcv, self.kcc, self.alpha = knn, kcv, kcc, float(alpha)
        self.keyframe_convex: list = []
        self.keyframe_concave: list = []
        self.prev: list | None = None

    def select(self, keyframe_positions, current_position):
        """-> (sorted unique keyframe indices, submap_hasChanged)"""
        pos = np.asarray(keyframe_positions, dtype=np.float32).reshape(-1, 3)
        cur = np.asarray(current_position, dtype=np.float32).reshape(3)
        # float differences, pow(.,2) and sqrt in double, stored as float (odom.cc:1255-1259)
        diff = (cur[None, :] - pos).astype(np.float64)
        ds = np.sqrt((diff ** 2).sum(axis=1)).astype(np.float32)
        out: list = []
        push_submap_indices(ds, self.knn, list(range(pos.shape[0])), out)
        if pos.shape[0] >= 4:
            self.keyframe_convex = convex_hull_vertices(pos)
        push_submap_indices([ds[c] for c in self.keyframe_convex], self.kcv, self.keyframe_convex, out)
        if pos.shape[0] >= 5:
            self.keyframe_concave = concave_hull_vertices(pos, self.alpha)
        push_submap_indices([ds[c] for c in self.keyframe_concave], self.kcc, self.keyframe_concave, out)
        cur_idx = sorted(set(out))
        changed = cur_idx != self.prev
        if changed:
            self.prev = cur_idx
        return cur_idx, changed
