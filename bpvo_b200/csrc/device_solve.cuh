// device_solve.cuh -- register-resident 6x6 solve and pose update for the on-device GN loop.
//
// Same algorithm as host_math.h (Eigen-3.2-style pivoted LDL^T, isApprox acceptance, damped fp64 retry;
// reference: pose_estimator_base.h:90-148), but written so that EVERY array index is a compile-time constant
// after unrolling: the 21 lower-triangle entries, the transpositions and the right-hand side all live in
// registers instead of local memory.  One thread runs it in ~1 us instead of ~7 us.
#pragma once

#include "host_math.h"

namespace bp {

template <int K, int Q> struct SwapRowsCols {
  // Eigen's in-place symmetric transposition k <-> q on lower storage (LDLT.h, ldlt_inplace<Lower>::unblocked)
  static __device__ __forceinline__ void run(float (&a)[6][6]) {
    constexpr int S = 6 - Q - 1;
#pragma unroll
    for (int j = 0; j < K; ++j) { const float t = a[K][j]; a[K][j] = a[Q][j]; a[Q][j] = t; }
#pragma unroll
    for (int i = 0; i < S; ++i) { const float t = a[6 - S + i][K]; a[6 - S + i][K] = a[6 - S + i][Q]; a[6 - S + i][Q] = t; }
    { const float t = a[K][K]; a[K][K] = a[Q][Q]; a[Q][Q] = t; }
#pragma unroll
    for (int i = K + 1; i < Q; ++i) { const float t = a[i][K]; a[i][K] = a[Q][i]; a[Q][i] = t; }
  }
};

template <int K, bool PIVOT> struct LdltStep {
  static __device__ __forceinline__ void run(float (&a)[6][6], int (&tr)[6], float& cutoff, bool& done) {
    if (!done) {
      int idx = K; float big = fabsf(a[K][K]);
      if (PIVOT) {
#pragma unroll
        for (int i = K + 1; i < 6; ++i) { const float v = fabsf(a[i][i]); if (v > big) { big = v; idx = i; } }
      }
      if (K == 0) cutoff = fabsf(FLT_EPSILON * big);
      if (big < cutoff) {
        done = true;                 // remaining transpositions stay identity
      } else {
        tr[K] = idx;
        if (PIVOT) {
        if (K + 1 < 6 && idx == K + 1) SwapRowsCols<K, (K + 1 < 6 ? K + 1 : 5)>::run(a);
        if (K + 2 < 6 && idx == K + 2) SwapRowsCols<K, (K + 2 < 6 ? K + 2 : 5)>::run(a);
        if (K + 3 < 6 && idx == K + 3) SwapRowsCols<K, (K + 3 < 6 ? K + 3 : 5)>::run(a);
        if (K + 4 < 6 && idx == K + 4) SwapRowsCols<K, (K + 4 < 6 ? K + 4 : 5)>::run(a);
        if (K + 5 < 6 && idx == K + 5) SwapRowsCols<K, (K + 5 < 6 ? K + 5 : 5)>::run(a);
        }
        float temp[6];
        if (K > 0) {
#pragma unroll
          for (int j = 0; j < K; ++j) temp[j] = a[j][j] * a[K][j];
          float s = 0.0f;
#pragma unroll
          for (int j = 0; j < K; ++j) s += a[K][j] * temp[j];
          a[K][K] -= s;
#pragma unroll
          for (int i = K + 1; i < 6; ++i) {
            float t = 0.0f;
#pragma unroll
            for (int j = 0; j < K; ++j) t += a[i][j] * temp[j];
            a[i][K] -= t;
          }
        }
        if (K < 5 && fabsf(a[K][K]) > cutoff) {
          const float id = 1.0f / a[K][K];       // one reciprocal per pivot (the host path divides entry by entry)
#pragma unroll
          for (int i = K + 1; i < 6; ++i) a[i][K] *= id;
        }
      }
    }
  }
};

// returns true when (H*dp).isApprox(G) holds for the fp32 factorisation.  PIVOT = false skips Eigen's diagonal pivoting
// (all swap code compiles away: ~4x fewer instructions on the single-thread critical path); the caller retries with
// PIVOT = true and finally with the generic damped fp64 solve6() when the acceptance test fails.
template <bool PIVOT>
__device__ __forceinline__ bool solve6_fp32_registers(const float* __restrict__ H, const float* __restrict__ G, float* __restrict__ dp) {
  float a[6][6]; int tr[6]; float x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < 6; ++j) a[i][j] = H[j * 6 + i];
    tr[i] = i; x[i] = G[i];
  }
  float cutoff = 0.0f; bool done = false;
  LdltStep<0, PIVOT>::run(a, tr, cutoff, done); LdltStep<1, PIVOT>::run(a, tr, cutoff, done); LdltStep<2, PIVOT>::run(a, tr, cutoff, done);
  LdltStep<3, PIVOT>::run(a, tr, cutoff, done); LdltStep<4, PIVOT>::run(a, tr, cutoff, done); LdltStep<5, PIVOT>::run(a, tr, cutoff, done);
  // solve: P b, L^-1, D^-1 (with Eigen 3.2's tolerance), L^-T, P^T
  if (PIVOT) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
      for (int q = k + 1; q < 6; ++q) if (tr[k] == q) { const float t = x[k]; x[k] = x[q]; x[q] = t; }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) x[i] -= a[i][j] * x[j];
  }
  float dmax = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; ++i) dmax = fmaxf(dmax, fabsf(a[i][i]));
  const float tol = fmaxf(dmax * FLT_EPSILON, 1.0f / FLT_MAX);
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = (fabsf(a[i][i]) > tol) ? x[i] / a[i][i] : 0.0f;
#pragma unroll
  for (int i = 5; i >= 0; --i) {
#pragma unroll
    for (int j = i + 1; j < 6; ++j) x[i] -= a[j][i] * x[j];
  }
  if (PIVOT) {
#pragma unroll
    for (int k = 5; k >= 0; --k) {
#pragma unroll
      for (int q = k + 1; q < 6; ++q) if (tr[k] == q) { const float t = x[k]; x[k] = x[q]; x[q] = t; }
    }
  }
  float d = 0.0f, na = 0.0f, nb = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float s = H[0 * 6 + i] * x[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) s += H[k * 6 + i] * x[k];
    d += (s - G[i]) * (s - G[i]); na += s * s; nb += G[i] * G[i];
    dp[i] = x[i];
  }
  return d <= 1e-5f * 1e-5f * fminf(na, nb);
}

// Fast path of the on-device loop: the same unpivoted LDL^T / substitutions / isApprox test as
// solve6_fp32_registers<false>, written WITHOUT any data-dependent branch (straight-line code: the scheduler overlaps the
// independent chains).  Rejects (returns false; the caller retries with the exact Eigen-order paths) when a pivot is not
// safely positive or the acceptance test fails, so an accepted result is the one the branchy version produces.
__device__ __forceinline__ bool solve6_fast(const float* __restrict__ H, const float* __restrict__ G, float* __restrict__ dp) {
  float a[6][6], x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) a[i][j] = H[j * 6 + i];
    x[i] = G[i];
  }
  const float cutoff = fabsf(FLT_EPSILON * a[0][0]);
  bool good = true;
#pragma unroll
  for (int K = 0; K < 6; ++K) {
    float temp[6];
#pragma unroll
    for (int j = 0; j < K; ++j) temp[j] = a[j][j] * a[K][j];
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < K; ++j) s += a[K][j] * temp[j];
    a[K][K] -= s;
#pragma unroll
    for (int i = K + 1; i < 6; ++i) {
      float t = 0.0f;
#pragma unroll
      for (int j = 0; j < K; ++j) t += a[i][j] * temp[j];
      a[i][K] -= t;
    }
    good = good && (a[K][K] > cutoff);        // also false for NaN
    const float id = 1.0f / a[K][K];
#pragma unroll
    for (int i = K + 1; i < 6; ++i) a[i][K] *= id;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) x[i] -= a[i][j] * x[j];
  }
  float dmax = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; ++i) dmax = fmaxf(dmax, a[i][i]);
  const float tol = fmaxf(dmax * FLT_EPSILON, 1.0f / FLT_MAX);      // Eigen 3.2 zeroes the components whose pivot is below this
#pragma unroll
  for (int i = 0; i < 6; ++i) { good = good && (a[i][i] > tol); x[i] = x[i] / a[i][i]; }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
#pragma unroll
    for (int j = i + 1; j < 6; ++j) x[i] -= a[j][i] * x[j];
  }
  float d = 0.0f, na = 0.0f, nb = 0.0f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float s = H[0 * 6 + i] * x[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) s += H[k * 6 + i] * x[k];
    d += (s - G[i]) * (s - G[i]); na += s * s; nb += G[i] * G[i];
    dp[i] = x[i];
  }
  return good && d <= 1e-5f * 1e-5f * fminf(na, nb);
}

// T <- T * (Tn^-1 exp(-dp) Tn) with Tn = [sI, -s c; 0 1] in closed form:
//   Tn^-1 [R t; 0 1] Tn = [R, c - R c + t / s; 0 1]      (rigid_body_warp.h:130-138, math_utils.h:140-168)
// For |w| < 0.5 rad (every sane GN step) the Rodrigues coefficients are evaluated as series in theta^2:
//   R = I + A W + B W^2,  t = v + B (w x v) + C (w x (w x v)),  A = sin(t)/t, B = (1 - cos t)/t^2, C = (t - sin t)/t^3
// -- no sqrt, no division, no sincos call, no branch on theta: the serial critical path of the GN loop is ~4x shorter.
__device__ __forceinline__ void apply_update(M44& T, const float dp[6], float s, float c1, float c2, float c3) {
  const float w0 = -dp[0], w1 = -dp[1], w2 = -dp[2], v0 = -dp[3], v1 = -dp[4], v2 = -dp[5];
  float R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {v0, v1, v2};
  const float th2 = w0 * w0 + w1 * w1 + w2 * w2;
  if (th2 < 0.25f) {
    const float A = 1.0f + th2 * (-1.0f / 6 + th2 * (1.0f / 120 + th2 * (-1.0f / 5040 + th2 * (1.0f / 362880 + th2 * (-1.0f / 39916800)))));
    const float B = 0.5f + th2 * (-1.0f / 24 + th2 * (1.0f / 720 + th2 * (-1.0f / 40320 + th2 * (1.0f / 3628800 + th2 * (-1.0f / 479001600)))));
    const float Cc = 1.0f / 6 + th2 * (-1.0f / 120 + th2 * (1.0f / 5040 + th2 * (-1.0f / 362880 + th2 * (1.0f / 39916800))));
    // W^2 = w w^T - theta^2 I
    const float w[3] = {w0, w1, w2};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = ((i == j) ? 1.0f - B * th2 : 0.0f) + B * w[i] * w[j];
    R[0][1] -= A * w2; R[0][2] += A * w1; R[1][0] += A * w2; R[1][2] -= A * w0; R[2][0] -= A * w1; R[2][1] += A * w0;
    const float x0 = w1 * v2 - w2 * v1, x1 = w2 * v0 - w0 * v2, x2 = w0 * v1 - w1 * v0;          // w x v
    const float y0 = w1 * x2 - w2 * x1, y1 = w2 * x0 - w0 * x2, y2 = w0 * x1 - w1 * x0;          // w x (w x v)
    t[0] = v0 + B * x0 + Cc * y0; t[1] = v1 + B * x1 + Cc * y1; t[2] = v2 + B * x2 + Cc * y2;
  } else {
    const float theta = sqrtf(th2);
    float sn, cs; sincosf(theta, &sn, &cs);
    const float hs = sinf(0.5f * theta);
    const float a = sn, b = 2.0f * hs * hs;        // 1 - cos(theta) without cancellation
    const float t_i = 1.0f / theta;
    const float S[3][3] = {{0.0f, t_i * (-w2), t_i * w1}, {t_i * w2, 0.0f, t_i * (-w0)}, {t_i * (-w1), t_i * w0, 0.0f}};
    float S2[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) S2[i][j] = S[i][0] * S[0][j] + S[i][1] * S[1][j] + S[i][2] * S[2][j];
    const float k1 = b * t_i, k2 = (theta - a) * t_i;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float id = (i == j) ? 1.0f : 0.0f;
        R[i][j] = id + a * S[i][j] + b * S2[i][j];
        acc += (id + k1 * S[i][j] + k2 * S2[i][j]) * ((j == 0) ? v0 : (j == 1) ? v1 : v2);
      }
      t[i] = acc;
    }
  }
  const float is = 1.0f / s;
  float m[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    m[i][0] = R[i][0]; m[i][1] = R[i][1]; m[i][2] = R[i][2];
    const float ci = (i == 0) ? c1 : (i == 1) ? c2 : c3;
    m[i][3] = ci - (R[i][0] * c1 + R[i][1] * c2 + R[i][2] * c3) + t[i] * is;
  }
  // T * [m; 0 0 0 1]
  M44 out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float acc = T(i, 0) * m[0][j] + T(i, 1) * m[1][j] + T(i, 2) * m[2][j];
      if (j == 3) acc += T(i, 3);
      out(i, j) = acc;
    }
  }
  T = out;
}

}  // namespace bp
