"""Host mirror of OdomNode::getSubmapKeyframes (reference src/dlo/odom.cc:1240-1293, pushSubmapIndices :1210-1233,
computeConvexHull :1017-1050, computeConcaveHull :1057-1090) — SURVEY section 8f row N3.

Which keyframes form the scan-to-map target: the `knn` keyframes nearest to the current pose, the `kcv` nearest among
the vertices of the 3-D convex hull of all keyframe positions, and the `kcc` nearest among the vertices of their 3-D
alpha shape ("concave hull", alpha = keyframe threshD at construction, odom.cc:95-98).  A few dozen points per call:
host work in the reference and here.  The reference gets both hulls from PCL, which calls qhull; this mirror calls
the same library through scipy.spatial (convex hull: qhull's default options as pcl::ConvexHull's 3-D path; Delaunay
tetrahedra with "QJ" as pcl::ConcaveHull's "qhull d QJ") and restates PCL's alpha filter on top of it.

The functions in this file are the qhull-based CHECKER of the product implementation, which is C++ behind the C ABI
(csrc/submap_select.cpp: ngicp_submap_*, own incremental hull and Bowyer-Watson Delaunay, no qhull); NativeSubmapSelector
below is its ctypes wrapper — what the replay benchmark uses — and tests/test_submap_select.py compares the two.
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
    kk = min(k, d.size)
    kth = np.partition(d, kk - 1)[kk - 1] if k > 0 else np.float32(-np.inf)
    out.extend(np.asarray(frames, dtype=np.int64)[d <= kth].tolist())


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
    # all tetrahedra at once: circumsphere radius = |Voronoi centre - vertex 0| (qh_pointdist(vertex, facet->center))
    P = p[tets]                                                   # (T, 4, 3)
    A = 2.0 * (P[:, 1:, :] - P[:, :1, :])
    b = (P[:, 1:, :] ** 2).sum(axis=2) - (P[:, :1, :] ** 2).sum(axis=2)
    det = np.linalg.det(A)
    ok = np.abs(det) > 0
    centre = np.zeros((tets.shape[0], 3))
    if ok.any():
        centre[ok] = np.linalg.solve(A[ok], b[ok][..., None])[..., 0]
    r_tet = np.where(ok, np.linalg.norm(centre - P[:, 0, :], axis=1), np.inf)
    good = r_tet <= alpha
    # all faces at once: face f is opposite vertex f
    face_idx = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    faces = tets[:, face_idx]                                     # (T, 4, 3) point indices
    F = pf[faces]                                                 # float coordinates, as pcl::getCircumcircleRadius sees them
    l1 = np.linalg.norm(F[:, :, 1] - F[:, :, 0], axis=2).astype(np.float64)
    l2 = np.linalg.norm(F[:, :, 2] - F[:, :, 1], axis=2).astype(np.float64)
    l3 = np.linalg.norm(F[:, :, 0] - F[:, :, 2], axis=2).astype(np.float64)
    sp = (l1 + l2 + l3) / 2.0
    area2 = sp * (sp - l1) * (sp - l2) * (sp - l3)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_face = (l1 * l2 * l3) / (4.0 * np.sqrt(np.where(area2 > 0, area2, 0.0)))
    nb_good = np.where(nbrs >= 0, good[np.clip(nbrs, 0, None)], False)
    candidate = good[:, None] | nb_good | (r_face <= alpha)
    boundary = (nbrs < 0) | ~good[:, None] | ~nb_good
    keep = candidate & boundary
    return sorted(set(int(v) for v in faces[keep].reshape(-1)))


class SubmapSelector:
    """State OdomNode keeps between scans for this step: keyframe_convex / keyframe_concave survive calls in which
    the hull is not recomputed (fewer than 4 / 5 keyframes), submap_kf_idx_prev decides submap_hasChanged."""

    def __init__(self, knn: int = 10, kcv: int = 10, kcc: int = 10, alpha: float = 5.0):
        # defaults: cfg/params.yaml submap.keyframe.{knn,kcv,kcc} = 10, keyframe.threshD = 5.0 (alpha, odom.cc:97)
        import scipy.spatial  # noqa: F401  (pay the import here, not inside the first scan that needs a hull)
        self.knn, self.kcv, self.kcc, self.alpha = knn, kcv, kcc, float(alpha)
        self.keyframe_convex: list = []
        self.keyframe_concave: list = []
        self.prev: list | None = None
        self._hulls_of = None     # keyframe positions the hulls were computed for (they only change with a new keyframe;
                                  # the reference recomputes them every scan, with the same result)

    def select(self, keyframe_positions, current_position):
        """-> (sorted unique keyframe indices, submap_hasChanged)"""
        pos = np.asarray(keyframe_positions, dtype=np.float32).reshape(-1, 3)
        cur = np.asarray(current_position, dtype=np.float32).reshape(3)
        # float differences, pow(.,2) and sqrt in double, stored as float (odom.cc:1255-1259)
        diff = (cur[None, :] - pos).astype(np.float64)
        ds = np.sqrt((diff ** 2).sum(axis=1)).astype(np.float32)
        out: list = []
        push_submap_indices(ds, self.knn, np.arange(pos.shape[0]), out)
        fresh = self._hulls_of is None or self._hulls_of.shape != pos.shape or not np.array_equal(self._hulls_of, pos)
        if pos.shape[0] >= 4 and fresh:
            self.keyframe_convex = convex_hull_vertices(pos)
        push_submap_indices(ds[self.keyframe_convex], self.kcv, self.keyframe_convex, out)
        if pos.shape[0] >= 5 and fresh:
            self.keyframe_concave = concave_hull_vertices(pos, self.alpha)
        self._hulls_of = pos.copy()
        push_submap_indices(ds[self.keyframe_concave], self.kcc, self.keyframe_concave, out)
        cur_idx = sorted(set(out))
        changed = cur_idx != self.prev
        if changed:
            self.prev = cur_idx
        return cur_idx, changed


# ---------------------------------------------------------------------------------------------------------------------
# the product implementation: C++ behind the C ABI (csrc/submap_select.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def _f32(a, cols):
    import ctypes as C
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1, cols))
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def native_push_submap_indices(dists, k: int, frames) -> list:
    import ctypes as C
    from . import _lib
    L = _lib.load()
    d = np.ascontiguousarray(np.asarray(dists, dtype=np.float32))
    f = np.ascontiguousarray(np.asarray(frames, dtype=np.int32))
    out = np.zeros(max(d.size, 1), dtype=np.int32)
    m = L.ngicp_submap_push_indices(d.ctypes.data_as(C.POINTER(C.c_float)), f.ctypes.data_as(C.POINTER(C.c_int)), d.size, int(k),
                                    out.ctypes.data_as(C.POINTER(C.c_int)), out.size)
    if m < 0:
        raise RuntimeError(f"ngicp_submap_push_indices: {m}")
    return out[:m].tolist()


def native_convex_hull_vertices(pos) -> list:
    import ctypes as C
    from . import _lib
    L = _lib.load()
    a, ap = _f32(pos, 3)
    out = np.zeros(max(a.shape[0], 1), dtype=np.int32)
    m = L.ngicp_submap_convex_hull(ap, a.shape[0], out.ctypes.data_as(C.POINTER(C.c_int)), out.size)
    if m < 0:
        raise RuntimeError(f"ngicp_submap_convex_hull: {m}")
    return out[:m].tolist()


def native_concave_hull_vertices(pos, alpha: float) -> list:
    import ctypes as C
    from . import _lib
    L = _lib.load()
    a, ap = _f32(pos, 3)
    out = np.zeros(max(a.shape[0], 1), dtype=np.int32)
    m = L.ngicp_submap_concave_hull(ap, a.shape[0], float(alpha), out.ctypes.data_as(C.POINTER(C.c_int)), out.size)
    if m < 0:
        raise RuntimeError(f"ngicp_submap_concave_hull: {m}")
    return out[:m].tolist()


def native_keyframe_wanted(kf_pos, kf_quat_wxyz, cur_pos, cur_quat_wxyz, thresh_dist: float, thresh_rot_deg: float) -> bool:
    from . import _lib
    L = _lib.load()
    a, ap = _f32(kf_pos, 3)
    q, qp = _f32(kf_quat_wxyz, 4)
    c, cp = _f32(cur_pos, 3)
    cq, cqp = _f32(cur_quat_wxyz, 4)
    r = L.ngicp_keyframe_wanted(ap, qp, a.shape[0], cp, cqp, float(thresh_dist), float(thresh_rot_deg))
    if r < 0:
        raise RuntimeError(f"ngicp_keyframe_wanted: {r}")
    return bool(r)


class NativeSubmapSelector:
    """ngicp_submap_select: same interface and state as SubmapSelector, computed by the C++ library."""

    def __init__(self, knn: int = 10, kcv: int = 10, kcc: int = 10, alpha: float = 5.0):
        import ctypes as C
        from . import _lib
        self._L = _lib.load()
        self._h = C.c_void_p()
        rc = self._L.ngicp_submap_selector_create(int(knn), int(kcv), int(kcc), float(alpha), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"ngicp_submap_selector_create: {rc}")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ngicp_submap_selector_destroy(self._h)
            self._h = None

    def _hull(self, which):
        import ctypes as C
        out = np.zeros(1 << 16, dtype=np.int32)
        m = self._L.ngicp_submap_selector_hulls(self._h, which, out.ctypes.data_as(C.POINTER(C.c_int)), out.size)
        return out[:max(m, 0)].tolist()

    @property
    def keyframe_convex(self):
        return self._hull(0)

    @property
    def keyframe_concave(self):
        return self._hull(1)

    def select(self, keyframe_positions, current_position):
        import ctypes as C
        a, ap = _f32(keyframe_positions, 3)
        c, cp = _f32(current_position, 3)
        out = np.zeros(max(a.shape[0], 1), dtype=np.int32)
        changed = C.c_int(0)
        m = self._L.ngicp_submap_select(self._h, ap, a.shape[0], cp, out.ctypes.data_as(C.POINTER(C.c_int)), out.size, C.byref(changed))
        if m < 0:
            raise RuntimeError(f"ngicp_submap_select: {m}")
        return out[:m].tolist(), bool(changed.value)
