// gicp_math.cuh — small fixed-size fp64 linear algebra used by the kernels and by the host-side LM
// driver (all __host__ __device__).  Everything is written for registers: symmetric 3x3 matrices
// are 6 values {xx,xy,xz,yy,yz,zz}; loops over matrix indices are compile-time unrolled.
//
// Reference arithmetic replaced here (reference paths):
//   Eigen::JacobiSVD<Matrix3d> + U diag(v) V^T        include/nano_gicp/impl/nano_gicp_impl.hpp:332-352
//   (C_B + T C_A T^T)^-1 with the 4th row/col pinned   include/nano_gicp/impl/nano_gicp_impl.hpp:205-209
//   Eigen::LDLT<6x6>::solve                            include/nano_gicp/impl/lsq_registration_impl.hpp:147,172
//   so3_exp + Quaternion::toRotationMatrix             include/nano_gicp/gicp/so3.hpp:99-118
//   is_converged                                       include/nano_gicp/impl/lsq_registration_impl.hpp:118-127
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>

#define NG_HD __host__ __device__ __forceinline__

namespace ngicp {

// symmetric 3x3: s[0]=xx s[1]=xy s[2]=xz s[3]=yy s[4]=yz s[5]=zz
NG_HD void sym3_inverse(const double s[6], double o[6]) {
  const double c00 = s[3] * s[5] - s[4] * s[4];
  const double c01 = s[2] * s[4] - s[1] * s[5];
  const double c02 = s[1] * s[4] - s[2] * s[3];
  const double det = s[0] * c00 + s[1] * c01 + s[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = c01 * id;
  o[2] = c02 * id;
  o[3] = (s[0] * s[5] - s[2] * s[2]) * id;
  o[4] = (s[1] * s[2] - s[0] * s[4]) * id;
  o[5] = (s[0] * s[3] - s[1] * s[1]) * id;
}

// o = B + R A R^T   (R row-major 3x3: R[r*3+c])
NG_HD void sym3_rcr(const double B[6], const double R[9], const double A[6], double o[6]) {
  // RA = R * A (3x3)
  double RA[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double r0 = R[r * 3 + 0], r1 = R[r * 3 + 1], r2 = R[r * 3 + 2];
    RA[r * 3 + 0] = r0 * A[0] + r1 * A[1] + r2 * A[2];
    RA[r * 3 + 1] = r0 * A[1] + r1 * A[3] + r2 * A[4];
    RA[r * 3 + 2] = r0 * A[2] + r1 * A[4] + r2 * A[5];
  }
  o[0] = B[0] + RA[0] * R[0] + RA[1] * R[1] + RA[2] * R[2];
  o[1] = B[1] + RA[0] * R[3] + RA[1] * R[4] + RA[2] * R[5];
  o[2] = B[2] + RA[0] * R[6] + RA[1] * R[7] + RA[2] * R[8];
  o[3] = B[3] + RA[3] * R[3] + RA[4] * R[4] + RA[5] * R[5];
  o[4] = B[4] + RA[3] * R[6] + RA[4] * R[7] + RA[5] * R[8];
  o[5] = B[5] + RA[6] * R[6] + RA[7] * R[7] + RA[8] * R[8];
}

// ---------------------------------------------------------------------------------------------
// two-sided Jacobi SVD of a real 3x3 (the algorithm Eigen::JacobiSVD runs for square real input).
// W, U, V are row-major 3x3 held in registers; the (p,q) pair is a template parameter.
// ---------------------------------------------------------------------------------------------
struct JRot { double c, s; };

NG_HD JRot jacobi_sym2(double x, double y, double z) {
  JRot r;
  const double deno = 2.0 * fabs(y);
  if (deno < DBL_MIN) { r.c = 1.0; r.s = 0.0; return r; }
  const double tau = (x - z) / deno;
  const double w = sqrt(tau * tau + 1.0);
  const double t = tau > 0.0 ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0.0 ? 1.0 : -1.0;
  const double n = 1.0 / sqrt(t * t + 1.0);
  r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
  r.c = n;
  return r;
}

template <int P, int Q>
NG_HD bool svd3_pair(double W[9], double U[9], double V[9], double& maxDiag) {
  const double thr = fmax(DBL_MIN, 2.0 * DBL_EPSILON * maxDiag);
  if (!(fabs(W[P * 3 + Q]) > thr || fabs(W[Q * 3 + P]) > thr)) return false;
  // 2x2 block -> make it symmetric with rot1, then diagonalise with a Jacobi rotation
  const double m00 = W[P * 3 + P], m01 = W[P * 3 + Q], m10 = W[Q * 3 + P], m11 = W[Q * 3 + Q];
  JRot r1;
  const double t = m00 + m11, d = m10 - m01;
  if (fabs(d) < DBL_MIN) { r1.s = 0.0; r1.c = 1.0; }
  else { const double u = t / d; const double tmp = sqrt(1.0 + u * u); r1.s = 1.0 / tmp; r1.c = u / tmp; }
  const double a00 = r1.c * m00 + r1.s * m10, a01 = r1.c * m01 + r1.s * m11, a11 = -r1.s * m01 + r1.c * m11;
  const JRot jr = jacobi_sym2(a00, a01, a11);
  // j_left = rot1 * j_right^T
  JRot jl;
  jl.c = r1.c * jr.c + r1.s * jr.s;
  jl.s = -r1.c * jr.s + r1.s * jr.c;
  // W <- J_left applied to rows P,Q
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double x = W[P * 3 + i], y = W[Q * 3 + i];
    W[P * 3 + i] = jl.c * x + jl.s * y;
    W[Q * 3 + i] = -jl.s * x + jl.c * y;
  }
  // U <- U * J_left^T on columns P,Q  (applyOnTheRight(p,q,jl.transpose()) == rotate columns by jl)
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double x = U[i * 3 + P], y = U[i * 3 + Q];
    U[i * 3 + P] = jl.c * x + jl.s * y;
    U[i * 3 + Q] = -jl.s * x + jl.c * y;
  }
  // W <- W * J_right, V <- V * J_right on columns P,Q (rotate columns by jr^T)
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double x = W[i * 3 + P], y = W[i * 3 + Q];
    W[i * 3 + P] = jr.c * x - jr.s * y;
    W[i * 3 + Q] = jr.s * x + jr.c * y;
    x = V[i * 3 + P]; y = V[i * 3 + Q];
    V[i * 3 + P] = jr.c * x - jr.s * y;
    V[i * 3 + Q] = jr.s * x + jr.c * y;
  }
  maxDiag = fmax(maxDiag, fmax(fabs(W[P * 3 + P]), fabs(W[Q * 3 + Q])));
  return true;
}

template <int I, int J>
NG_HD void svd3_swapcols(double sv[3], double U[9], double V[9]) {
  double t = sv[I]; sv[I] = sv[J]; sv[J] = t;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    t = U[r * 3 + I]; U[r * 3 + I] = U[r * 3 + J]; U[r * 3 + J] = t;
    t = V[r * 3 + I]; V[r * 3 + I] = V[r * 3 + J]; V[r * 3 + J] = t;
  }
}

// A row-major 3x3 in; U, sv (descending), V out
NG_HD void svd3(const double A[9], double U[9], double sv[3], double V[9]) {
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) scale = fmax(scale, fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  double W[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { W[i] = A[i] / scale; U[i] = (i % 4 == 0) ? 1.0 : 0.0; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  double maxDiag = fmax(fabs(W[0]), fmax(fabs(W[4]), fabs(W[8])));
  for (int sweep = 0; sweep < 64; sweep++) {
    bool any = false;
    any |= svd3_pair<1, 0>(W, U, V, maxDiag);
    any |= svd3_pair<2, 0>(W, U, V, maxDiag);
    any |= svd3_pair<2, 1>(W, U, V, maxDiag);
    if (!any) break;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double a = W[i * 3 + i];
    sv[i] = fabs(a) * scale;
    if (a < 0.0) { U[0 * 3 + i] = -U[0 * 3 + i]; U[1 * 3 + i] = -U[1 * 3 + i]; U[2 * 3 + i] = -U[2 * 3 + i]; }
  }
  // selection sort, descending, first maximum wins (maxCoeff semantics)
  if (sv[1] > sv[0] && sv[1] >= sv[2]) svd3_swapcols<0, 1>(sv, U, V);
  else if (sv[2] > sv[0] && sv[2] > sv[1]) svd3_swapcols<0, 2>(sv, U, V);
  if (sv[2] > sv[1]) svd3_swapcols<1, 2>(sv, U, V);
}

// covariance regularisation (nano_gicp_impl.hpp:323-353).  cov: full symmetric 3x3 as sym6; out sym6.
NG_HD void regularize_cov(const double cov[6], int method, double out[6]) {
  if (method == 0) {  // NONE
#pragma unroll
    for (int i = 0; i < 6; i++) out[i] = cov[i];
    return;
  }
  if (method == 4) {  // FROBENIUS
    double C[6] = {cov[0] + 1e-3, cov[1], cov[2], cov[3] + 1e-3, cov[4], cov[5] + 1e-3};
    double Ci[6];
    sym3_inverse(C, Ci);
    const double nrm = sqrt(Ci[0] * Ci[0] + Ci[3] * Ci[3] + Ci[5] * Ci[5] + 2.0 * (Ci[1] * Ci[1] + Ci[2] * Ci[2] + Ci[4] * Ci[4]));
#pragma unroll
    for (int i = 0; i < 6; i++) Ci[i] /= nrm;
    sym3_inverse(Ci, out);
    return;
  }
  const double A[9] = {cov[0], cov[1], cov[2], cov[1], cov[3], cov[4], cov[2], cov[4], cov[5]};
  double U[9], V[9], sv[3], val[3];
  svd3(A, U, sv, V);
  if (method == 3) { val[0] = 1.0; val[1] = 1.0; val[2] = 1e-3; }               // PLANE
  else if (method == 1) { val[0] = fmax(sv[0], 1e-3); val[1] = fmax(sv[1], 1e-3); val[2] = fmax(sv[2], 1e-3); }  // MIN_EIG
  else { const double mx = fmax(sv[0], fmax(sv[1], sv[2])); val[0] = fmax(sv[0] / mx, 1e-3); val[1] = fmax(sv[1] / mx, 1e-3); val[2] = fmax(sv[2] / mx, 1e-3); }
  // out(r,c) = sum_j U(r,j) val_j V(c,j), upper triangle
  out[0] = U[0] * val[0] * V[0] + U[1] * val[1] * V[1] + U[2] * val[2] * V[2];
  out[1] = U[0] * val[0] * V[3] + U[1] * val[1] * V[4] + U[2] * val[2] * V[5];
  out[2] = U[0] * val[0] * V[6] + U[1] * val[1] * V[7] + U[2] * val[2] * V[8];
  out[3] = U[3] * val[0] * V[3] + U[4] * val[1] * V[4] + U[5] * val[2] * V[5];
  out[4] = U[3] * val[0] * V[6] + U[4] * val[1] * V[7] + U[5] * val[2] * V[8];
  out[5] = U[6] * val[0] * V[6] + U[7] * val[1] * V[7] + U[8] * val[2] * V[8];
}

// ---------------------------------------------------------------------------------------------
// SE(3) state and the scalar side of Gauss-Newton / Levenberg-Marquardt
// ---------------------------------------------------------------------------------------------
struct Iso3 { double R[9]; double t[3]; };  // R row-major

NG_HD void iso_identity(Iso3& x) {
#pragma unroll
  for (int i = 0; i < 9; i++) x.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  x.t[0] = x.t[1] = x.t[2] = 0.0;
}
NG_HD void iso_from_colmajor16(const double* T, Iso3& x) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) x.R[r * 3 + c] = T[c * 4 + r];
    x.t[r] = T[12 + r];
  }
}
NG_HD void iso_to_colmajor16(const Iso3& x, double* T) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) T[c * 4 + r] = x.R[r * 3 + c];
    T[12 + r] = x.t[r];
    T[r * 4 + 3] = 0.0;
  }
  T[15] = 1.0;
}
NG_HD void iso_mul(const Iso3& a, const Iso3& b, Iso3& o) {
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) o.R[r * 3 + c] = a.R[r * 3 + 0] * b.R[0 * 3 + c] + a.R[r * 3 + 1] * b.R[1 * 3 + c] + a.R[r * 3 + 2] * b.R[2 * 3 + c];
    o.t[r] = a.R[r * 3 + 0] * b.t[0] + a.R[r * 3 + 1] * b.t[1] + a.R[r * 3 + 2] * b.t[2] + a.t[r];
  }
}

// delta = [so3_exp(d[0:3]) | d[3:6]]
NG_HD void delta_from_step(const double d[6], Iso3& delta) {
  const double ox = d[0], oy = d[1], oz = d[2];
  const double theta_sq = ox * ox + oy * oy + oz * oz;
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq);
    const double half_theta = 0.5 * theta;
    imag_factor = sin(half_theta) / theta;
    real_factor = cos(half_theta);
  }
  const double w = real_factor, x = imag_factor * ox, y = imag_factor * oy, z = imag_factor * oz;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  delta.R[0] = 1.0 - (tyy + tzz); delta.R[1] = txy - twz;         delta.R[2] = txz + twy;
  delta.R[3] = txy + twz;         delta.R[4] = 1.0 - (txx + tzz); delta.R[5] = tyz - twx;
  delta.R[6] = txz - twy;         delta.R[7] = tyz + twx;         delta.R[8] = 1.0 - (txx + tyy);
  delta.t[0] = d[3]; delta.t[1] = d[4]; delta.t[2] = d[5];
}

NG_HD bool lm_is_converged(const Iso3& delta, double rot_eps, double trans_eps) {
  double rmax = 0.0, tmax = 0.0;
#pragma unroll
  for (int i = 0; i < 9; i++) rmax = fmax(rmax, 1.0 / rot_eps * fabs(delta.R[i] - ((i % 4 == 0) ? 1.0 : 0.0)));
#pragma unroll
  for (int i = 0; i < 3; i++) tmax = fmax(tmax, 1.0 / trans_eps * fabs(delta.t[i]));
  return fmax(rmax, tmax) < 1.0;
}

// Pivoted LDL^T solve of the symmetric 6x6 system A x = rhs (A full col-major or row-major: symmetric).
// Pivots on the largest remaining diagonal and zeroes solution rows whose pivot is ~0, like Eigen::LDLT.
__host__ __device__ inline void ldlt6_solve(const double* Ain, const double* rhs, double* xout) {
  double A[36];
  for (int i = 0; i < 36; i++) A[i] = Ain[i];
#define NG_A(r, c) A[(c) * 6 + (r)]
  int perm[6];
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double big = fabs(NG_A(k, k));
    for (int i = k + 1; i < 6; i++) if (fabs(NG_A(i, i)) > big) { big = fabs(NG_A(i, i)); piv = i; }
    perm[k] = piv;
    if (piv != k) {
      for (int j = 0; j < k; j++) { double t = NG_A(k, j); NG_A(k, j) = NG_A(piv, j); NG_A(piv, j) = t; }
      for (int i = piv + 1; i < 6; i++) { double t = NG_A(i, k); NG_A(i, k) = NG_A(i, piv); NG_A(i, piv) = t; }
      { double t = NG_A(k, k); NG_A(k, k) = NG_A(piv, piv); NG_A(piv, piv) = t; }
      for (int i = k + 1; i < piv; i++) { double t = NG_A(i, k); NG_A(i, k) = NG_A(piv, i); NG_A(piv, i) = t; }
    }
    if (k > 0) {
      double tmp[6];
      for (int j = 0; j < k; j++) tmp[j] = NG_A(j, j) * NG_A(k, j);
      double s = 0.0;
      for (int j = 0; j < k; j++) s += NG_A(k, j) * tmp[j];
      NG_A(k, k) -= s;
      for (int i = k + 1; i < 6; i++) {
        double t = 0.0;
        for (int j = 0; j < k; j++) t += NG_A(i, j) * tmp[j];
        NG_A(i, k) -= t;
      }
    }
    const double dk = NG_A(k, k);
    if (fabs(dk) > 0.0) for (int i = k + 1; i < 6; i++) NG_A(i, k) /= dk;
  }
  double x[6];
  for (int i = 0; i < 6; i++) x[i] = rhs[i];
  for (int k = 0; k < 6; k++) { double t = x[k]; x[k] = x[perm[k]]; x[perm[k]] = t; }
  for (int i = 0; i < 6; i++) for (int j = 0; j < i; j++) x[i] -= NG_A(i, j) * x[j];
  const double tol = 1.0 / DBL_MAX;
  for (int i = 0; i < 6; i++) x[i] = fabs(NG_A(i, i)) > tol ? x[i] / NG_A(i, i) : 0.0;
  for (int i = 5; i >= 0; i--) for (int j = i + 1; j < 6; j++) x[i] -= NG_A(j, i) * x[j];
  for (int k = 5; k >= 0; k--) { double t = x[k]; x[k] = x[perm[k]]; x[perm[k]] = t; }
  for (int i = 0; i < 6; i++) xout[i] = x[i];
#undef NG_A
}

// Fast path for the LM system: unpivoted LDL^T of a symmetric positive-definite 6x6 held in registers
// (fully unrolled).  Returns false when a pivot is not safely positive — the caller then uses the pivoted
// routine above, which also reproduces Eigen's behaviour on singular systems.
NG_HD bool ldlt6_solve_spd(const double* A, const double* rhs, double* x) {
  double L[6][6], D[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double dj = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; k++) dj -= L[j][k] * L[j][k] * D[k];
    D[j] = dj;
    ok = ok && (dj > 1e-280) && (dj < 1e280);
    const double inv = 1.0 / dj;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double v = A[j * 6 + i];   // symmetric: (i,j) == (j,i)
#pragma unroll
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * D[k];
      L[i][j] = v * inv;
    }
  }
  if (!ok) return false;
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double v = rhs[i];
#pragma unroll
    for (int k = 0; k < i; k++) v -= L[i][k] * y[k];
    y[i] = v;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] /= D[i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double v = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) v -= L[k][i] * x[k];
    x[i] = v;
  }
  return true;
}

NG_HD void lm_solve(const double* A, const double* rhs, double* x) {
  if (!ldlt6_solve_spd(A, rhs, x)) ldlt6_solve(A, rhs, x);
}

// The reduced quantities of one linearisation: 21 upper-triangle H entries, 6 b entries, error.
constexpr int NRED = 28;
// index of H(r,c), r<=c, in the packed upper triangle
NG_HD int hidx(int r, int c) { return r * 6 - (r * (r - 1)) / 2 + (c - r); }

NG_HD void unpack_H(const double* red, double* H36) {
  for (int r = 0; r < 6; r++)
    for (int c = r; c < 6; c++) { const double v = red[hidx(r, c)]; H36[c * 6 + r] = v; H36[r * 6 + c] = v; }
}

// per-point contribution to H (packed upper triangle), b and error.
//   e = q_B - T p_A ; J = [skew(T p_A) | -I] ; H += J^T M J ; b += J^T M e ; err += e^T M e
// (nano_gicp_impl.hpp:238-257).  M symmetric sym6.  With S = skew(a):  J^T M J =
//   [ S^T M S   -S^T M ]
//   [ -M S       M     ]
NG_HD void gicp_accumulate(const double a[3] /* T p_A */, const double e[3], const double M[6], double* acc /* NRED */) {
  // full M
  const double m00 = M[0], m01 = M[1], m02 = M[2], m11 = M[3], m12 = M[4], m22 = M[5];
  // MS = M * S, S = [[0,-az,ay],[az,0,-ax],[-ay,ax,0]]
  const double ax = a[0], ay = a[1], az = a[2];
  const double ms00 = m01 * az - m02 * ay, ms01 = -m00 * az + m02 * ax, ms02 = m00 * ay - m01 * ax;
  const double ms10 = m11 * az - m12 * ay, ms11 = -m01 * az + m12 * ax, ms12 = m01 * ay - m11 * ax;
  const double ms20 = m12 * az - m22 * ay, ms21 = -m02 * az + m22 * ax, ms22 = m02 * ay - m12 * ax;
  // S^T (MS): S^T = [[0,az,-ay],[-az,0,ax],[ay,-ax,0]]
  acc[hidx(0, 0)] += az * ms10 - ay * ms20;
  acc[hidx(0, 1)] += az * ms11 - ay * ms21;
  acc[hidx(0, 2)] += az * ms12 - ay * ms22;
  acc[hidx(1, 1)] += -az * ms01 + ax * ms21;
  acc[hidx(1, 2)] += -az * ms02 + ax * ms22;
  acc[hidx(2, 2)] += ay * ms02 - ax * ms12;
  // top-right block: -S^T M = -(M S)^T  => H(r, 3+c) = -MS(c, r)
  acc[hidx(0, 3)] -= ms00; acc[hidx(0, 4)] -= ms10; acc[hidx(0, 5)] -= ms20;
  acc[hidx(1, 3)] -= ms01; acc[hidx(1, 4)] -= ms11; acc[hidx(1, 5)] -= ms21;
  acc[hidx(2, 3)] -= ms02; acc[hidx(2, 4)] -= ms12; acc[hidx(2, 5)] -= ms22;
  // bottom-right: M
  acc[hidx(3, 3)] += m00; acc[hidx(3, 4)] += m01; acc[hidx(3, 5)] += m02;
  acc[hidx(4, 4)] += m11; acc[hidx(4, 5)] += m12; acc[hidx(5, 5)] += m22;
  // Me
  const double me0 = m00 * e[0] + m01 * e[1] + m02 * e[2];
  const double me1 = m01 * e[0] + m11 * e[1] + m12 * e[2];
  const double me2 = m02 * e[0] + m12 * e[1] + m22 * e[2];
  // b = J^T M e = [S^T Me ; -Me]
  acc[21] += az * me1 - ay * me2;
  acc[22] += -az * me0 + ax * me2;
  acc[23] += ay * me0 - ax * me1;
  acc[24] -= me0; acc[25] -= me1; acc[26] -= me2;
  acc[27] += e[0] * me0 + e[1] * me1 + e[2] * me2;
}

}  // namespace ngicp
