// submap_select.cpp — SURVEY §8f row N3 in the host language: which keyframes form the scan-to-map target.
//   OdomNode::pushSubmapIndices    reference src/dlo/odom.cc:1210-1233
//   OdomNode::computeConvexHull    :1017-1050   (pcl::ConvexHull<PointXYZI>, 3-D  -> qhull)
//   OdomNode::computeConcaveHull   :1057-1090   (pcl::ConcaveHull<PointXYZI>, 3-D, alpha = keyframe threshD -> qhull "d QJ")
//   OdomNode::getSubmapKeyframes   :1240-1293   (selection part; the cloud / covariance concatenation is the keyframe store)
//   OdomNode::updateKeyframes      :1102-1153   (the new-keyframe decision)
// A few dozen to a few hundred keyframe positions per call: host work in the reference and here (plain C++, no CUDA,
// no PCL, no qhull).  The reference reaches qhull through PCL; neither exists in this image, so both hulls are built
// here from their definitions:
//   convex hull   incremental 3-D hull (visible facets / horizon), orientation tests in long double.  Flat input (all
//                 points coplanar) yields no hull, as qhull refuses it ("initial simplex is flat") and PCL returns nothing.
//   alpha shape   Delaunay tetrahedra by Bowyer-Watson insertion (in-sphere tests in long double, input joggled by
//                 ~1e-9 of its extent like qhull's QJ, but from a fixed hash of the point index: deterministic), then PCL's
//                 3-D alpha filter (surface/include/pcl/surface/impl/concave_hull.hpp): a tetrahedron is good when its
//                 circumsphere radius <= alpha; candidate triangles are the faces of good tetrahedra plus any face whose
//                 own circumcircle radius <= alpha; a candidate belongs to the shape unless the tetrahedra on BOTH of
//                 its sides are good.  The result is the set of input points those triangles use.
// On inputs in general position — and on exactly flat or nearly flat ones, a robot on a level floor — the vertex sets
// equal qhull's (tests/test_submap_select.py compares with scipy's qhull on random clouds, flat sets and a replay-like loop); on degenerate input (exactly cospherical / coplanar subsets) the
// reference itself is not deterministic (QJ joggles with a random seed).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <queue>
#include <vector>

#include "../../include/nanogicp_c.h"

namespace {

typedef long double R;
struct P3 { R x, y, z; };

inline P3 sub(const P3& a, const P3& b) { return P3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 cross(const P3& a, const P3& b) { return P3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline R dot(const P3& a, const P3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// > 0 when d lies on the side of plane (a,b,c) its normal (b-a)x(c-a) points to
inline R orient3d(const P3& a, const P3& b, const P3& c, const P3& d) { return dot(cross(sub(b, a), sub(c, a)), sub(d, a)); }

// ---------------------------------------------------------------------------------------------------------------
// convex hull
// ---------------------------------------------------------------------------------------------------------------
struct Face { int v[3]; bool alive; };

void convex_hull_vertices(const std::vector<P3>& p, std::vector<int>& out) {
  out.clear();
  const int n = (int)p.size();
  if (n < 4) return;
  // initial simplex: extreme in x, farthest from it, farthest from the line, farthest from the plane
  int i0 = 0;
  for (int i = 1; i < n; i++) if (p[i].x < p[i0].x) i0 = i;
  int i1 = -1; R best = 0;
  for (int i = 0; i < n; i++) { const P3 d = sub(p[i], p[i0]); const R v = dot(d, d); if (v > best) { best = v; i1 = i; } }
  if (i1 < 0) return;
  const R scale2 = best;                                  // squared extent
  int i2 = -1; best = 0;
  for (int i = 0; i < n; i++) { const P3 c = cross(sub(p[i1], p[i0]), sub(p[i], p[i0])); const R v = dot(c, c); if (v > best) { best = v; i2 = i; } }
  if (i2 < 0 || best <= scale2 * scale2 * 1e-24L) return;  // collinear
  int i3 = -1; best = 0;
  const P3 nrm = cross(sub(p[i1], p[i0]), sub(p[i2], p[i0]));
  for (int i = 0; i < n; i++) { const R v = fabsl(dot(nrm, sub(p[i], p[i0]))); if (v > best) { best = v; i3 = i; } }
  // flat: the farthest point is within 1e-12 of the extent of the plane (qhull: "initial simplex is flat")
  if (i3 < 0 || best <= sqrtl(dot(nrm, nrm)) * sqrtl(scale2) * 1e-12L) return;
  if (orient3d(p[i0], p[i1], p[i2], p[i3]) > 0) std::swap(i1, i2);     // i3 below plane (i0,i1,i2): all faces outward
  std::vector<Face> faces;
  auto add = [&](int a, int b, int c) { Face f; f.v[0] = a; f.v[1] = b; f.v[2] = c; f.alive = true; faces.push_back(f); };
  add(i0, i1, i2); add(i0, i3, i1); add(i1, i3, i2); add(i2, i3, i0);
  std::vector<char> visible;
  std::map<std::pair<int, int>, int> edge_face;           // directed edge of a visible face -> 1
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1 || i == i2 || i == i3) continue;
    visible.assign(faces.size(), 0);
    bool any = false;
    for (size_t f = 0; f < faces.size(); f++) {
      if (!faces[f].alive) continue;
      if (orient3d(p[faces[f].v[0]], p[faces[f].v[1]], p[faces[f].v[2]], p[i]) > 0) { visible[f] = 1; any = true; }
    }
    if (!any) continue;                                    // inside (or on) the current hull
    edge_face.clear();
    for (size_t f = 0; f < faces.size(); f++)
      if (visible[f])
        for (int e = 0; e < 3; e++) edge_face[std::make_pair(faces[f].v[e], faces[f].v[(e + 1) % 3])] = 1;
    const size_t nf = faces.size();
    for (size_t f = 0; f < nf; f++) {
      if (!visible[f]) continue;
      for (int e = 0; e < 3; e++) {
        const int a = faces[f].v[e], b = faces[f].v[(e + 1) % 3];
        if (edge_face.find(std::make_pair(b, a)) == edge_face.end()) add(a, b, i);   // horizon edge: twin is not visible
      }
      faces[f].alive = false;
    }
  }
  std::vector<char> used(n, 0);
  for (const Face& f : faces) if (f.alive) for (int e = 0; e < 3; e++) used[f.v[e]] = 1;
  for (int i = 0; i < n; i++) if (used[i]) out.push_back(i);
}

// ---------------------------------------------------------------------------------------------------------------
// Delaunay tetrahedra (Bowyer-Watson) + PCL's alpha filter
// ---------------------------------------------------------------------------------------------------------------
struct Tet { int v[4]; bool alive; };

// > 0 when e is inside the circumsphere of the positively oriented tetrahedron (a,b,c,d) [orient3d(a,b,c,d) > 0]
R insphere(const P3& a, const P3& b, const P3& c, const P3& d, const P3& e) {
  const P3 A = sub(a, e), B = sub(b, e), C = sub(c, e), D = sub(d, e);
  const R a2 = dot(A, A), b2 = dot(B, B), c2 = dot(C, C), d2 = dot(D, D);
  // 4x4 determinant | A a2 ; B b2 ; C c2 ; D d2 |, expanded along the last column
  auto det3 = [](const P3& u, const P3& v, const P3& w) { return dot(u, cross(v, w)); };
  const R det = -a2 * det3(B, C, D) + b2 * det3(A, C, D) - c2 * det3(A, B, D) + d2 * det3(A, B, C);
  return -det;                                              // sign convention fixed by the self-check in delaunay()
}

inline uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

struct FaceKey {
  int a, b, c;
  bool operator<(const FaceKey& o) const { return a != o.a ? a < o.a : (b != o.b ? b < o.b : c < o.c); }
};
inline FaceKey face_key(int a, int b, int c) {
  if (a > b) std::swap(a, b);
  if (b > c) std::swap(b, c);
  if (a > b) std::swap(a, b);
  return FaceKey{a, b, c};
}

// tets over points p[0..n) (joggled copies q); returns false when the insertion broke down (degenerate beyond the joggle)
bool delaunay(const std::vector<P3>& q, int n, R extent, const P3& centre, std::vector<Tet>& tets) {
  std::vector<P3> pts(q);
  // enclosing tetrahedron far outside the data (vertices n..n+3)
  const R L = extent * 4096.0L;
  pts.push_back(P3{centre.x - L, centre.y - L, centre.z - L});
  pts.push_back(P3{centre.x + L, centre.y + L, centre.z - L});
  pts.push_back(P3{centre.x + L, centre.y - L, centre.z + L});
  pts.push_back(P3{centre.x - L, centre.y + L, centre.z + L});
  tets.clear();
  {
    Tet t; t.v[0] = n; t.v[1] = n + 1; t.v[2] = n + 2; t.v[3] = n + 3; t.alive = true;
    if (orient3d(pts[t.v[0]], pts[t.v[1]], pts[t.v[2]], pts[t.v[3]]) < 0) std::swap(t.v[0], t.v[1]);
    tets.push_back(t);
  }
  // sign self-check of insphere: the centre of the data is inside the enclosing tetrahedron's circumsphere
  const R sgn = insphere(pts[tets[0].v[0]], pts[tets[0].v[1]], pts[tets[0].v[2]], pts[tets[0].v[3]], centre) > 0 ? 1.0L : -1.0L;
  std::map<FaceKey, int> count;
  std::vector<int> bad;
  for (int i = 0; i < n; i++) {
    bad.clear();
    for (size_t t = 0; t < tets.size(); t++) {
      if (!tets[t].alive) continue;
      const Tet& T = tets[t];
      if (sgn * insphere(pts[T.v[0]], pts[T.v[1]], pts[T.v[2]], pts[T.v[3]], pts[i]) > 0) bad.push_back((int)t);
    }
    if (bad.empty()) return false;
    count.clear();
    static const int F[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};   // face opposite vertex f, outward for a positive tet
    for (int t : bad)
      for (int f = 0; f < 4; f++) count[face_key(tets[t].v[F[f][0]], tets[t].v[F[f][1]], tets[t].v[F[f][2]])]++;
    const size_t nt = tets.size();
    for (int t : bad) {
      for (int f = 0; f < 4; f++) {
        const int a = tets[t].v[F[f][0]], b = tets[t].v[F[f][1]], c = tets[t].v[F[f][2]];
        if (count[face_key(a, b, c)] != 1) continue;       // shared by two removed tetrahedra: inside the cavity
        Tet nt_; nt_.v[0] = a; nt_.v[1] = b; nt_.v[2] = c; nt_.v[3] = i; nt_.alive = true;
        const R o = orient3d(pts[a], pts[b], pts[c], pts[i]);
        if (o == 0) return false;                          // the new point lies in the plane of a cavity face
        if (o < 0) std::swap(nt_.v[0], nt_.v[1]);
        tets.push_back(nt_);
      }
    }
    for (int t : bad) if ((size_t)t < nt) tets[t].alive = false;
  }
  return true;
}

void concave_hull_vertices(const std::vector<P3>& p, const float* xyz_f, double alpha, std::vector<int>& out) {
  out.clear();
  const int n = (int)p.size();
  if (n < 5) return;
  P3 lo = p[0], hi = p[0];
  for (const P3& v : p) {
    lo.x = std::min(lo.x, v.x); lo.y = std::min(lo.y, v.y); lo.z = std::min(lo.z, v.z);
    hi.x = std::max(hi.x, v.x); hi.y = std::max(hi.y, v.y); hi.z = std::max(hi.z, v.z);
  }
  const R extent = std::max(std::max(hi.x - lo.x, hi.y - lo.y), std::max(hi.z - lo.z, (R)1e-30L));
  const P3 centre{(lo.x + hi.x) / 2, (lo.y + hi.y) / 2, (lo.z + hi.z) / 2};
  std::vector<Tet> tets;
  std::vector<P3> q(n);
  bool ok = false;
  for (int attempt = 0; attempt < 8 && !ok; attempt++) {
    // joggle: +-(1e-9 * 10^attempt) of the extent per coordinate, from a hash of (index, axis, attempt) — like qhull's QJ,
    // which also retries with a ten times larger joggle until the triangulation succeeds (exactly flat or cospherical
    // input needs a visible perturbation: keyframes of a robot on a perfectly level floor)
    const R amp = extent * 1e-9L * powl(10.0L, (R)attempt);
    for (int i = 0; i < n; i++) {
      const uint32_t h = (uint32_t)(i * 3 + attempt * 0x9e3779b9u);
      q[i].x = p[i].x + amp * ((R)hash32(h) / 4294967296.0L * 2 - 1);
      q[i].y = p[i].y + amp * ((R)hash32(h + 1) / 4294967296.0L * 2 - 1);
      q[i].z = p[i].z + amp * ((R)hash32(h + 2) / 4294967296.0L * 2 - 1);
    }
    ok = delaunay(q, n, extent, centre, tets);
  }
  if (!ok) return;
  // tetrahedra of data points only
  std::vector<Tet> real;
  for (const Tet& t : tets) if (t.alive && t.v[0] < n && t.v[1] < n && t.v[2] < n && t.v[3] < n) real.push_back(t);
  const int T = (int)real.size();
  std::vector<char> good(T, 0);
  for (int t = 0; t < T; t++) {
    // circumcentre c: 2 (p_j - p_0) . c = |p_j|^2 - |p_0|^2  (qhull's Voronoi centre of the facet); radius = |c - p_0|
    const P3& a = q[real[t].v[0]];
    const P3 r1 = sub(q[real[t].v[1]], a), r2 = sub(q[real[t].v[2]], a), r3 = sub(q[real[t].v[3]], a);
    const R det = dot(r1, cross(r2, r3));
    if (det == 0) continue;
    const R b1 = dot(r1, r1) / 2, b2 = dot(r2, r2) / 2, b3 = dot(r3, r3) / 2;
    const P3 c23 = cross(r2, r3), c31 = cross(r3, r1), c12 = cross(r1, r2);
    const P3 c{(b1 * c23.x + b2 * c31.x + b3 * c12.x) / det, (b1 * c23.y + b2 * c31.y + b3 * c12.y) / det, (b1 * c23.z + b2 * c31.z + b3 * c12.z) / det};
    good[t] = sqrtl(dot(c, c)) <= (R)alpha ? 1 : 0;
  }
  // faces -> the (at most two) tetrahedra behind them
  std::map<FaceKey, std::pair<int, int>> behind;
  static const int F[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};
  for (int t = 0; t < T; t++)
    for (int f = 0; f < 4; f++) {
      const FaceKey k = face_key(real[t].v[F[f][0]], real[t].v[F[f][1]], real[t].v[F[f][2]]);
      auto it = behind.find(k);
      if (it == behind.end()) behind[k] = std::make_pair(t, -1);
      else it->second.second = t;
    }
  // Triangles on the outside of the triangulation are also the real face of a tetrahedron with ONE vertex of the enclosing
  // tetrahedron.  For (nearly) flat input — keyframes of a robot on a level floor — ALL triangles are of that kind: the
  // slivers qhull would report have circumspheres so large that they swallow the far vertices and never form.  Such a
  // sliver is never "good", so the triangle counts as a face with no good tetrahedron behind it.
  for (const Tet& t : tets) {
    if (!t.alive) continue;
    int far = 0, far_at = -1;
    for (int j = 0; j < 4; j++) if (t.v[j] >= n) { far++; far_at = j; }
    if (far != 1) continue;
    const FaceKey k = face_key(t.v[(far_at + 1) & 3], t.v[(far_at + 2) & 3], t.v[(far_at + 3) & 3]);
    if (behind.find(k) == behind.end()) behind[k] = std::make_pair(-1, -1);
  }
  std::vector<char> used(n, 0);
  for (const auto& kv : behind) {
    const int t0 = kv.second.first, t1 = kv.second.second;
    const bool g0 = t0 >= 0 && good[t0] != 0, g1 = t1 >= 0 && good[t1] != 0;
    if (g0 && g1) continue;                                // interior of the shape
    bool candidate = g0 || g1;
    if (!candidate) {
      // pcl::getCircumcircleRadius: Heron's formula on the three side lengths of the float points (lengths in float)
      const int idx[3] = {kv.first.a, kv.first.b, kv.first.c};
      float l[3];
      for (int e = 0; e < 3; e++) {
        const float* u = xyz_f + 3 * idx[e];
        const float* v = xyz_f + 3 * idx[(e + 1) % 3];
        const float dx = v[0] - u[0], dy = v[1] - u[1], dz = v[2] - u[2];
        l[e] = std::sqrt(dx * dx + dy * dy + dz * dz);
      }
      const double l1 = l[0], l2 = l[1], l3 = l[2];
      const double s = (l1 + l2 + l3) / 2.0;
      const double area2 = s * (s - l1) * (s - l2) * (s - l3);
      const double area = area2 > 0 ? std::sqrt(area2) : 0.0;
      const double r = (l1 * l2 * l3) / (4.0 * area);      // +inf for a degenerate triangle
      candidate = r <= alpha;
    }
    if (candidate) { used[kv.first.a] = used[kv.first.b] = used[kv.first.c] = 1; }
  }
  for (int i = 0; i < n; i++) if (used[i]) out.push_back(i);
}

std::vector<P3> to_points(const float* xyz, int n) {
  std::vector<P3> p((size_t)n);
  for (int i = 0; i < n; i++) p[i] = P3{(R)xyz[3 * i], (R)xyz[3 * i + 1], (R)xyz[3 * i + 2]};
  return p;
}

int copy_out(const std::vector<int>& v, int* out, int cap) {
  if ((int)v.size() > cap) return NGICP_E_INVALID;
  for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
  return (int)v.size();
}

}  // namespace

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

int ngicp_submap_push_indices(const float* dists, const int* frames, int n, int k, int* out, int out_cap) {
  if (n < 0 || (n > 0 && (!dists || !frames)) || !out) return NGICP_E_INVALID;
  // a max-heap of at most k distances; its top is the k-th smallest (odom.cc:1213-1224)
  std::priority_queue<float> pq;
  for (int i = 0; i < n; i++) {
    const float d = dists[i];
    if ((int)pq.size() >= k && !pq.empty() && pq.top() > d) { pq.push(d); pq.pop(); }
    else if ((int)pq.size() < k) pq.push(d);
  }
  if (pq.empty()) return 0;     // the reference reads the top of an empty heap here and then loops over nothing useful
  const float kth = pq.top();
  int m = 0;
  for (int i = 0; i < n; i++)
    if (dists[i] <= kth) { if (m >= out_cap) return NGICP_E_INVALID; out[m++] = frames[i]; }
  return m;
}

int ngicp_submap_convex_hull(const float* xyz, int n, int* out, int out_cap) {
  if (n < 0 || (n > 0 && !xyz) || !out) return NGICP_E_INVALID;
  std::vector<int> v;
  convex_hull_vertices(to_points(xyz, n), v);
  return copy_out(v, out, out_cap);
}

int ngicp_submap_concave_hull(const float* xyz, int n, double alpha, int* out, int out_cap) {
  if (n < 0 || (n > 0 && !xyz) || !out) return NGICP_E_INVALID;
  std::vector<int> v;
  concave_hull_vertices(to_points(xyz, n), xyz, alpha, v);
  return copy_out(v, out, out_cap);
}

struct ngicp_submap_selector {
  int knn, kcv, kcc;
  double alpha;
  std::vector<int> convex, concave, prev;
  bool has_prev;
  std::vector<float> hulls_of;     // the keyframe positions the hulls were computed for
};

int ngicp_submap_selector_create(int knn, int kcv, int kcc, double alpha, ngicp_submap_selector_t** out) {
  if (!out) return NGICP_E_INVALID;
  ngicp_submap_selector* s = new (std::nothrow) ngicp_submap_selector();
  if (!s) return NGICP_E_CUDA;
  s->knn = knn; s->kcv = kcv; s->kcc = kcc; s->alpha = alpha; s->has_prev = false;
  *out = s;
  return NGICP_OK;
}
void ngicp_submap_selector_destroy(ngicp_submap_selector_t* s) { delete s; }

int ngicp_submap_select(ngicp_submap_selector_t* s, const float* kf_xyz, int n, const float* cur_xyz, int* out, int out_cap, int* changed) {
  if (!s || n < 0 || (n > 0 && !kf_xyz) || !cur_xyz || !out) return NGICP_E_INVALID;
  // float differences, pow(., 2) and sqrt in double, stored as float (odom.cc:1255-1259)
  std::vector<float> ds((size_t)n);
  std::vector<int> all((size_t)n);
  for (int i = 0; i < n; i++) {
    const double dx = (double)(cur_xyz[0] - kf_xyz[3 * i]), dy = (double)(cur_xyz[1] - kf_xyz[3 * i + 1]), dz = (double)(cur_xyz[2] - kf_xyz[3 * i + 2]);
    ds[i] = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
    all[i] = i;
  }
  std::vector<int> cur, tmp((size_t)n + 1);
  auto push = [&](const std::vector<float>& d, const std::vector<int>& frames, int k) {
    const int m = ngicp_submap_push_indices(d.data(), frames.data(), (int)d.size(), k, tmp.data(), (int)tmp.size());
    for (int i = 0; i < m; i++) cur.push_back(tmp[i]);
  };
  push(ds, all, s->knn);
  // the hulls only change with the keyframe set (the reference recomputes them every scan, with the same result);
  // they are kept from the last call when there are too few keyframes (computeConvexHull / computeConcaveHull return early)
  const bool fresh = s->hulls_of.size() != (size_t)n * 3 || (n > 0 && std::memcmp(s->hulls_of.data(), kf_xyz, sizeof(float) * 3 * (size_t)n) != 0);
  if (fresh) {
    const std::vector<P3> p = to_points(kf_xyz, n);
    if (n >= 4) convex_hull_vertices(p, s->convex);
    if (n >= 5) concave_hull_vertices(p, kf_xyz, s->alpha, s->concave);
    s->hulls_of.assign(kf_xyz, kf_xyz + 3 * (size_t)n);
  }
  std::vector<float> hd;
  for (int c : s->convex) hd.push_back(c < n ? ds[c] : 0.f);
  push(hd, s->convex, s->kcv);
  hd.clear();
  for (int c : s->concave) hd.push_back(c < n ? ds[c] : 0.f);
  push(hd, s->concave, s->kcc);
  std::sort(cur.begin(), cur.end());
  cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
  const bool ch = !s->has_prev || cur != s->prev;
  if (ch) { s->prev = cur; s->has_prev = true; }
  if (changed) *changed = ch ? 1 : 0;
  return copy_out(cur, out, out_cap);
}

int ngicp_submap_selector_hulls(ngicp_submap_selector_t* s, int which, int* out, int out_cap) {
  if (!s || !out) return NGICP_E_INVALID;
  return copy_out(which == 0 ? s->convex : s->concave, out, out_cap);
}

int ngicp_keyframe_wanted(const float* kf_xyz, const float* kf_quat_wxyz, int n, const float* cur_xyz, const float* cur_quat_wxyz,
                          double thresh_dist, double thresh_rot_deg) {
  if (n <= 0 || !kf_xyz || !kf_quat_wxyz || !cur_xyz || !cur_quat_wxyz) return NGICP_E_INVALID;
  // odom.cc:1102-1153
  float closest_d = INFINITY;
  int closest = 0, num_nearby = 0;
  for (int i = 0; i < n; i++) {
    const float dd = (float)std::sqrt(std::pow((double)(cur_xyz[0] - kf_xyz[3 * i]), 2) + std::pow((double)(cur_xyz[1] - kf_xyz[3 * i + 1]), 2) +
                                      std::pow((double)(cur_xyz[2] - kf_xyz[3 * i + 2]), 2));
    if (dd <= thresh_dist * 1.5) ++num_nearby;
    if (dd < closest_d) { closest_d = dd; closest = i; }
  }
  const float dd = closest_d;
  // dq = rotq * closest_r^-1 (float quaternions)
  const float* c = kf_quat_wxyz + 4 * closest;
  const float n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3];
  const float iw = c[0] / n2, ix = -c[1] / n2, iy = -c[2] / n2, iz = -c[3] / n2;
  const float* a = cur_quat_wxyz;
  const float qw = a[0] * iw - a[1] * ix - a[2] * iy - a[3] * iz;
  const float qx = a[0] * ix + a[1] * iw + a[2] * iz - a[3] * iy;
  const float qy = a[0] * iy + a[2] * iw + a[3] * ix - a[1] * iz;
  const float qz = a[0] * iz + a[3] * iw + a[1] * iy - a[2] * ix;
  const float theta_rad = (float)(2. * std::atan2(std::sqrt(std::pow((double)qx, 2) + std::pow((double)qy, 2) + std::pow((double)qz, 2)), (double)qw));
  const float theta_deg = (float)(theta_rad * (180.0 / M_PI));
  bool nk = false;
  if (std::fabs(dd) > thresh_dist || std::fabs(theta_deg) > thresh_rot_deg) nk = true;
  if (std::fabs(dd) <= thresh_dist) nk = false;
  if (std::fabs(dd) <= thresh_dist && std::fabs(theta_deg) > thresh_rot_deg && num_nearby <= 1) nk = true;
  return nk ? 1 : 0;
}

}  // extern "C"
