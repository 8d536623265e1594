// host_math.h -- tiny fixed-size linear algebra shared by the host GN driver, the VisualOdometry
// state machine and (as __device__ code) the on-device GN loop.  Column-major like Eigen, so
// Matrix44::data() of the reference maps 1:1.
//
// What it stands in for (reference file:line):
//   * Eigen::LDLT<Matrix<float,6,6>>::compute/solve and (H*dp).isApprox(G)   pose_estimator_base.h:90-111
//   * the double / damped retry solve2Augmented                               pose_estimator_base.h:136-148
//   * math::TwistToMatrix                                                      math_utils.h:140-168
//   * RigidBodyWarp::paramsToPose / scalePose                                  rigid_body_warp.h:130-138
#pragma once

#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define BP_HD __host__ __device__ __forceinline__
#else
#define BP_HD inline
#endif

namespace bp {

struct M44 {
  float m[16];
  BP_HD float& operator()(int r, int c) { return m[c * 4 + r]; }
  BP_HD float operator()(int r, int c) const { return m[c * 4 + r]; }
};

BP_HD M44 identity44() {
  M44 r;
  for (int i = 0; i < 16; ++i) r.m[i] = 0.0f;
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}

BP_HD M44 mul44(const M44& a, const M44& b) {
  M44 r;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float s = a(i, 0) * b(0, j);
      for (int k = 1; k < 4; ++k) s += a(i, k) * b(k, j);
      r(i, j) = s;
    }
  return r;
}

// inverse of a rigid-body / similarity-like 4x4 through the general adjugate (Matrix4f::inverse())
BP_HD M44 inverse44(const M44& a) {
  const float* m = a.m;
  float inv[16];
  inv[0] = m[5]*m[10]*m[15] - m[5]*m[11]*m[14] - m[9]*m[6]*m[15] + m[9]*m[7]*m[14] + m[13]*m[6]*m[11] - m[13]*m[7]*m[10];
  inv[4] = -m[4]*m[10]*m[15] + m[4]*m[11]*m[14] + m[8]*m[6]*m[15] - m[8]*m[7]*m[14] - m[12]*m[6]*m[11] + m[12]*m[7]*m[10];
  inv[8] = m[4]*m[9]*m[15] - m[4]*m[11]*m[13] - m[8]*m[5]*m[15] + m[8]*m[7]*m[13] + m[12]*m[5]*m[11] - m[12]*m[7]*m[9];
  inv[12] = -m[4]*m[9]*m[14] + m[4]*m[10]*m[13] + m[8]*m[5]*m[14] - m[8]*m[6]*m[13] - m[12]*m[5]*m[10] + m[12]*m[6]*m[9];
  inv[1] = -m[1]*m[10]*m[15] + m[1]*m[11]*m[14] + m[9]*m[2]*m[15] - m[9]*m[3]*m[14] - m[13]*m[2]*m[11] + m[13]*m[3]*m[10];
  inv[5] = m[0]*m[10]*m[15] - m[0]*m[11]*m[14] - m[8]*m[2]*m[15] + m[8]*m[3]*m[14] + m[12]*m[2]*m[11] - m[12]*m[3]*m[10];
  inv[9] = -m[0]*m[9]*m[15] + m[0]*m[11]*m[13] + m[8]*m[1]*m[15] - m[8]*m[3]*m[13] - m[12]*m[1]*m[11] + m[12]*m[3]*m[9];
  inv[13] = m[0]*m[9]*m[14] - m[0]*m[10]*m[13] - m[8]*m[1]*m[14] + m[8]*m[2]*m[13] + m[12]*m[1]*m[10] - m[12]*m[2]*m[9];
  inv[2] = m[1]*m[6]*m[15] - m[1]*m[7]*m[14] - m[5]*m[2]*m[15] + m[5]*m[3]*m[14] + m[13]*m[2]*m[7] - m[13]*m[3]*m[6];
  inv[6] = -m[0]*m[6]*m[15] + m[0]*m[7]*m[14] + m[4]*m[2]*m[15] - m[4]*m[3]*m[14] - m[12]*m[2]*m[7] + m[12]*m[3]*m[6];
  inv[10] = m[0]*m[5]*m[15] - m[0]*m[7]*m[13] - m[4]*m[1]*m[15] + m[4]*m[3]*m[13] + m[12]*m[1]*m[7] - m[12]*m[3]*m[5];
  inv[14] = -m[0]*m[5]*m[14] + m[0]*m[6]*m[13] + m[4]*m[1]*m[14] - m[4]*m[2]*m[13] - m[12]*m[1]*m[6] + m[12]*m[2]*m[5];
  inv[3] = -m[1]*m[6]*m[11] + m[1]*m[7]*m[10] + m[5]*m[2]*m[11] - m[5]*m[3]*m[10] - m[9]*m[2]*m[7] + m[9]*m[3]*m[6];
  inv[7] = m[0]*m[6]*m[11] - m[0]*m[7]*m[10] - m[4]*m[2]*m[11] + m[4]*m[3]*m[10] + m[8]*m[2]*m[7] - m[8]*m[3]*m[6];
  inv[11] = -m[0]*m[5]*m[11] + m[0]*m[7]*m[9] + m[4]*m[1]*m[11] - m[4]*m[3]*m[9] - m[8]*m[1]*m[7] + m[8]*m[3]*m[5];
  inv[15] = m[0]*m[5]*m[10] - m[0]*m[6]*m[9] - m[4]*m[1]*m[10] + m[4]*m[2]*m[9] + m[8]*m[1]*m[6] - m[8]*m[2]*m[5];
  float det = m[0]*inv[0] + m[1]*inv[4] + m[2]*inv[8] + m[3]*inv[12];
  float idet = 1.0f / det;
  M44 r;
  for (int i = 0; i < 16; ++i) r.m[i] = inv[i] * idet;
  return r;
}

// se(3) exponential, rotation part first: p[0:3] = omega, p[3:6] = v  (math_utils.h:140-168)
BP_HD M44 twist_to_matrix(const float p[6]) {
  M44 ret = identity44();
  const float theta = sqrtf(p[0]*p[0] + p[1]*p[1] + p[2]*p[2]);
  if (theta > 1e-8) {
    // `T a = ::sin(theta)` with T = float resolves to the float overloads in a libstdc++ >= 6 build of the reference
    // (checked against the reference's own code in oracle/_ref); `1.0 - ::cos(theta)` and `1.0 / theta` go through double
    const float a = sinf(theta);
    const float b = (float) (1.0 - (double) cosf(theta));
    const float t_i = (float) (1.0 / (double) theta);
    float S[3][3], S2[3][3];
    S[0][0] = 0.0f;           S[0][1] = t_i * (-p[2]);  S[0][2] = t_i * p[1];
    S[1][0] = t_i * p[2];     S[1][1] = 0.0f;           S[1][2] = t_i * (-p[0]);
    S[2][0] = t_i * (-p[1]);  S[2][1] = t_i * p[0];     S[2][2] = 0.0f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        float s = S[i][0] * S[0][j]; s += S[i][1] * S[1][j]; s += S[i][2] * S[2][j];
        S2[i][j] = s;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ret(i, j) = ((i == j) ? 1.0f : 0.0f) + a * S[i][j] + b * S2[i][j];
    const float c1 = b * t_i, c2 = (theta - a) * t_i;
    for (int i = 0; i < 3; ++i) {
      float s = 0.0f;
      for (int j = 0; j < 3; ++j) {
        const float v = ((i == j) ? 1.0f : 0.0f) + c1 * S[i][j] + c2 * S2[i][j];
        s = (j == 0) ? v * p[3] : s + v * p[3 + j];
      }
      ret(i, 3) = s;
    }
  } else {
    ret(0, 3) = p[3]; ret(1, 3) = p[4]; ret(2, 3) = p[5];
  }
  return ret;
}

// Tn^-1 * exp(p) * Tn (rigid_body_warp.h:130-138)
BP_HD M44 params_to_pose(const M44& Tn, const M44& Tn_inv, const float p[6]) {
  return mul44(mul44(Tn_inv, twist_to_matrix(p)), Tn);
}

// pivoted LDL^T of a 6x6 SPD-ish matrix (lower storage), Eigen 3.2 semantics: largest-|diagonal|
// pivoting, epsilon cutoff, tolerance on D in the solve.
template <typename S> struct Ldlt6 {
  S a[6][6];
  int tr[6];
  BP_HD static S absS(S v) { return v < 0 ? -v : v; }
  BP_HD void compute(const S* H /* column-major 6x6 */, S eps) {
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) a[i][j] = H[j * 6 + i];
    S cutoff = 0, temp[6];
    for (int k = 0; k < 6; ++k) tr[k] = k;
    for (int k = 0; k < 6; ++k) {
      int idx = k; S big = absS(a[k][k]);
      for (int i = k + 1; i < 6; ++i) if (absS(a[i][i]) > big) { big = absS(a[i][i]); idx = i; }
      if (k == 0) cutoff = absS(eps * big);
      if (big < cutoff) { for (int i = k; i < 6; ++i) tr[i] = i; break; }
      tr[k] = idx;
      if (k != idx) {
        const int s = 6 - idx - 1;
        for (int j = 0; j < k; ++j) { S t = a[k][j]; a[k][j] = a[idx][j]; a[idx][j] = t; }
        for (int i = 0; i < s; ++i) { S t = a[6 - s + i][k]; a[6 - s + i][k] = a[6 - s + i][idx]; a[6 - s + i][idx] = t; }
        { S t = a[k][k]; a[k][k] = a[idx][idx]; a[idx][idx] = t; }
        for (int i = k + 1; i < idx; ++i) { S t = a[i][k]; a[i][k] = a[idx][i]; a[idx][i] = t; }
      }
      const int rs = 6 - k - 1;
      if (k > 0) {
        for (int j = 0; j < k; ++j) temp[j] = a[j][j] * a[k][j];
        S s = 0; for (int j = 0; j < k; ++j) s += a[k][j] * temp[j];
        a[k][k] -= s;
        for (int i = 0; i < rs; ++i) {
          S t = 0; for (int j = 0; j < k; ++j) t += a[k + 1 + i][j] * temp[j];
          a[k + 1 + i][k] -= t;
        }
      }
      if (rs > 0 && absS(a[k][k]) > cutoff)
        for (int i = 0; i < rs; ++i) a[k + 1 + i][k] /= a[k][k];
    }
  }
  BP_HD void solve(const S* b, S* x, S eps, S tiny) const {
    for (int i = 0; i < 6; ++i) x[i] = b[i];
    for (int k = 0; k < 6; ++k) if (tr[k] != k) { S t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < i; ++j) x[i] -= a[i][j] * x[j];
    S dmax = 0; for (int i = 0; i < 6; ++i) { S v = absS(a[i][i]); if (v > dmax) dmax = v; }
    S tol = dmax * eps; if (tiny > tol) tol = tiny;
    for (int i = 0; i < 6; ++i) { if (absS(a[i][i]) > tol) x[i] /= a[i][i]; else x[i] = 0; }
    for (int i = 5; i >= 0; --i) for (int j = i + 1; j < 6; ++j) x[i] -= a[j][i] * x[j];
    for (int k = 5; k >= 0; --k) if (tr[k] != k) { S t = x[k]; x[k] = x[tr[k]]; x[tr[k]] = t; }
  }
};

template <typename S> BP_HD bool is_approx6(const S* a, const S* b, S prec) {
  S d = 0, na = 0, nb = 0;
  for (int i = 0; i < 6; ++i) { d += (a[i] - b[i]) * (a[i] - b[i]); na += a[i] * a[i]; nb += b[i] * b[i]; }
  const S mn = na < nb ? na : nb;
  return d <= prec * prec * mn;
}

// PoseEstimatorData_::solve (pose_estimator_base.h:90-111) with the solve2Augmented(0.001) retry (:136-148)
BP_HD bool solve6(const float* H, const float* G, float* dp) {
  Ldlt6<float> l;
  l.compute(H, FLT_EPSILON);
  l.solve(G, dp, FLT_EPSILON, 1.0f / FLT_MAX);
  float Hd[6];
  for (int i = 0; i < 6; ++i) {
    float s = H[0 * 6 + i] * dp[0];
    for (int k = 1; k < 6; ++k) s += H[k * 6 + i] * dp[k];
    Hd[i] = s;
  }
  bool ok = is_approx6<float>(Hd, G, 1e-5f);
  if (!ok) {
    float dmax = H[0];
    for (int i = 1; i < 6; ++i) if (H[i * 6 + i] > dmax) dmax = H[i * 6 + i];
    const double u = 0.001 * (double) dmax;
    double Hq[36], Gq[6], dq[6], Hd2[6];
    for (int i = 0; i < 36; ++i) Hq[i] = (double) H[i];
    for (int i = 0; i < 6; ++i) { Gq[i] = (double) G[i]; Hq[i * 6 + i] += u; }
    Ldlt6<double> l2;
    l2.compute(Hq, DBL_EPSILON);
    l2.solve(Gq, dq, DBL_EPSILON, 1.0 / DBL_MAX);
    for (int i = 0; i < 6; ++i) {
      double s = Hq[0 * 6 + i] * dq[0];
      for (int k = 1; k < 6; ++k) s += Hq[k * 6 + i] * dq[k];
      Hd2[i] = s;
    }
    ok = is_approx6<double>(Hd2, Gq, 1e-12);
    for (int i = 0; i < 6; ++i) dp[i] = (float) dq[i];
  }
  return ok;
}

}  // namespace bp
