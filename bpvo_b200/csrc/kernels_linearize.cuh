// kernels_linearize.cuh -- the per-iteration hot path: one PoseEstimatorGN::linearize on the device,
// and the whole coarse-to-fine Gauss-Newton loop as ONE persistent cooperative kernel (sm_100a).
//
// Replaces (reference file:line):
//   RigidBodyWarp::setPose + PhotoError::init/run (fp64 projection, bilinear)   rigid_body_warp.h:111-114, photo_error.cc:344-389
//   replicateValidFlags                                                         pose_estimator_base.h:307-320
//   AutoScaleEstimator::estimateScale (exact median of |r| over valid)          mestimator.cc:452-490, utils.h:224-252
//   MEstimator::ComputeWeights (Huber / Tukey / L2)                             mestimator.cc:242-415
//   LinearSystemBuilder::Run (H = sum w J'J, G = sum w r J, sqrt(sum w r^2))    linear_system_builder.cc:140-284, 334-350
//   PoseEstimatorBase::run + solve + testConvergence (device loop)              pose_estimator_base.h:90-148, 258-282, 324-407
//   VisualOdometryPoseEstimator::estimatePose (level loop)                      vo_pose_estimator.cc:63-93
//
// Structure of one linearize (kernel boundaries in the host-driven path, one barrier + one exchange in the persistent one):
//   P1 residuals : thread per point: fp64 project -> floor -> valid -> 4 bilinear taps x C channels
//                  (one 32-B sector per tap for C = 8) -> r = f32(Iw - I0); bookkeeping for the robust scale
//   median       : exact order statistics n/2-1, n/2 of |r|.  Persistent kernel: bracket around the previous median ->
//                  bracket histogram + the few candidates in the two wanted bins (bracket_select); fallback and
//                  host-driven path: 3-level radix select over the float bit patterns (P2 select-2, P3 select-3)
//   P4 reduce    : sigma -> weight -> rank-2 per-point update of the 21+6+1 normal-equation scalars
//                  (J = gx*A + gy*B) -> transposing warp butterfly -> fp64 CTA totals -> fixed-order sum over CTAs
//                  (persistent: two-level flag-in-data exchange; host-driven: last-CTA fold)
//   solve        : (persistent) 6x6 LDL^T, acceptance test, pose update, convergence tests in every CTA
//   multi-GPU    : (peer-memory mode) the bracket histogram, the candidates and the 30 sums also cross the ranks, inside
//                  the kernel, through peer-mapped mailboxes
// Latency / L2 bound at semi-dense sizes, HBM bound for dense megapixel levels: no tensor cores (no dense contraction).
#pragma once

#include "device_types.h"
#include "device_solve.cuh"

namespace bp {

// Point -> thread mapping of every linearize phase: warp g = warp_in_cta * nblocks + cta owns points
// [32 g, 32 g + 32) (+ multiples of the grid size).  Small levels therefore spread one or two warps onto EVERY
// SM instead of filling a few CTAs, and a thread meets the same points in every phase (no cross-CTA hand-over).
__device__ __forceinline__ int first_point(int block, int nblocks) {
  return (((int) (threadIdx.x >> 5)) * nblocks + block) * 32 + (int) (threadIdx.x & 31);
}

// Streaming levels (template fields that do not fit the shared-memory cache, e.g. level 0 of a dense 1080p template: 49 points per
// thread) run one point per thread at a time, a dependent chain of load -> compute; with 8 warps per SM that keeps ~28 KB per SM in
// flight, about 45 % of the HBM bandwidth.  L2 prefetches issued two points ahead (template records) and one point ahead (the four
// bilinear taps, from an fp32 estimate of the projection) turn the demand loads into L2 hits: no registers, no shared memory.
#ifndef BP_PREFETCH
#define BP_PREFETCH 1
#endif
// BP_MSG_SELECT = 1: the median candidates travel as per-CTA flag-in-data messages (message_select) instead of grid barrier + global
// lists.  Built, parity-green -- and 8 % SLOWER in the same-box A/B (profiles/README.md): a bracket usually holds 1-2.5 k candidates,
// so some CTA nearly always has more than the 13 a 128-byte message carries and the iteration pays the message round AND the barrier.
#ifndef BP_BRACKET_THREAD
#define BP_BRACKET_THREAD (kLinThreads - 1)      // 0 = the round-2a arrangement: thread 0, followed by a CTA barrier (A/B)
#endif
#ifndef BP_MSG_SELECT
#define BP_MSG_SELECT 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if BP_PREFETCH
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

template <int C> struct VecC {
  float v[C];
  static_assert(C == 1 || C == 3 || C == 5 || C == 8, "channel counts of the supported descriptors");
  __device__ __forceinline__ void set(const float4& a, const float4& b) {
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = (c == 0) ? a.x : (c == 1) ? a.y : (c == 2) ? a.z : (c == 3) ? a.w : (c == 4) ? b.x : (c == 5) ? b.y : (c == 6) ? b.z : b.w;
  }
  __device__ __forceinline__ void load(const float* __restrict__ p) {
    if (C == 1) { v[0] = __ldg(p); return; }
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = (C > 4) ? __ldg(reinterpret_cast<const float4*>(p) + 1) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    set(a, b);
  }
  __device__ __forceinline__ void load_plain(const float* p) {   // data written earlier in the same kernel
    if (C == 1) { v[0] = *p; return; }
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = (C > 4) ? *(reinterpret_cast<const float4*>(p) + 1) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    set(a, b);
  }
  __device__ __forceinline__ void store(float* p) const {
    if (C == 1) { *p = v[0]; return; }
    float w[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) w[c] = (c < C) ? v[c < C ? c : 0] : 0.0f;
    reinterpret_cast<float4*>(p)[0] = make_float4(w[0], w[1], w[2], w[3]);
    if (C > 4) reinterpret_cast<float4*>(p)[1] = make_float4(w[4], w[5], w[6], w[7]);
  }
};

// Shared-memory copy of the template records (and residuals) of the points THIS CTA's threads own.  The on-device GN
// loop visits the same points ~50 times per level: staging (X,Y,Z), I0, gx, gy once per level and keeping r between the
// phases removes every per-iteration global load except the four bilinear taps.  Slot k of thread t holds point
// first_point + k * gridsize.  Layout (all conflict-free): float4 pts[K][512] | float f[K][4 fields][C][512] | u8 valid[K][512].
// Fields are addressed by 32-bit byte offsets into the kernel's dynamic shared memory (kTcNone = not cached: read from
// global memory): six registers instead of six 64-bit pointers in a kernel that sits at the register limit.
constexpr unsigned kTcNone = 0xffffffffu;
struct TplCache {
  unsigned pts;          // float4 [K][256]
  unsigned f[4];         // TC_I0, TC_GX, TC_GY, TC_R: float [K][C][256] each
  unsigned valid;        // u8 [K][256], cached together with TC_R
  int K;                 // slots per thread the level needs (0 = cache off: host-driven kernels)
};
enum { TC_I0 = 0, TC_GX = 1, TC_GY = 2, TC_R = 3 };
__host__ __device__ __forceinline__ TplCache tpl_cache_off() { TplCache t; t.pts = t.f[0] = t.f[1] = t.f[2] = t.f[3] = t.valid = kTcNone; t.K = 0; return t; }
__device__ __forceinline__ unsigned char* tc_base() { extern __shared__ __align__(16) unsigned char dyn_smem[]; return dyn_smem; }
__device__ __forceinline__ float4& tc_point(const TplCache& tc, int k) { return reinterpret_cast<float4*>(tc_base() + tc.pts)[k * kLinThreads + threadIdx.x]; }
__device__ __forceinline__ uint8_t& tc_valid(const TplCache& tc, int k) { return (tc_base() + tc.valid)[k * kLinThreads + threadIdx.x]; }

template <int C> __device__ __forceinline__ void tc_get(const TplCache& tc, int k, int field, VecC<C>& v) {
  const float* p = reinterpret_cast<const float*>(tc_base() + tc.f[field]) + ((size_t) k * C) * kLinThreads + threadIdx.x;
#pragma unroll
  for (int c = 0; c < C; ++c) v.v[c] = p[c * kLinThreads];
}
template <int C> __device__ __forceinline__ void tc_put(const TplCache& tc, int k, int field, const VecC<C>& v) {
  float* p = reinterpret_cast<float*>(tc_base() + tc.f[field]) + ((size_t) k * C) * kLinThreads + threadIdx.x;
#pragma unroll
  for (int c = 0; c < C; ++c) p[c * kLinThreads] = v.v[c];
}

// Which fields of a level live in shared memory: all-or-nothing per field, in the order of the global traffic they save
// per GN iteration -- residuals + valid flags (written by P1, read by the select passes and P4), points (P1 + P4),
// gx, gy (P4), I0 (P1).  KITTI semi-dense: everything (1 slot); KITTI dense level 0
// (11 slots): residuals + points; a 1080p level sharded over 8 GPUs (6 slots): everything but I0.  `first` = offset of the cache area, `bytes` its size.
template <int C>
__host__ __device__ __forceinline__ TplCache tpl_cache_plan(unsigned first, int bytes, int need) {
  TplCache t = tpl_cache_off();
  if (need <= 0) return t;
  t.K = need;
  const int sz_f = need * kLinThreads * 4 * C, sz_pts = need * kLinThreads * 16, sz_valid = need * kLinThreads;
  const bool has_r = sz_f + sz_valid <= bytes;   if (has_r) bytes -= sz_f + sz_valid;
  const bool has_p = sz_pts <= bytes;            if (has_p) bytes -= sz_pts;
  const bool has_gx = sz_f <= bytes;             if (has_gx) bytes -= sz_f;
  const bool has_gy = sz_f <= bytes;             if (has_gy) bytes -= sz_f;
  const bool has_i0 = sz_f <= bytes;
  unsigned p = first;                            // layout: pts (16-B aligned) | R | I0 | GX | GY | valid
  if (has_p) { t.pts = p; p += sz_pts; }
  if (has_r) { t.f[TC_R] = p; p += sz_f; }
  if (has_i0) { t.f[TC_I0] = p; p += sz_f; }
  if (has_gx) { t.f[TC_GX] = p; p += sz_f; }
  if (has_gy) { t.f[TC_GY] = p; p += sz_f; }
  if (has_r) t.valid = p;
  return t;
}

// stage this CTA's points of level L into the cache (thread-private slots: no barrier needed afterwards)
template <int C> __device__ __forceinline__ void tc_fill(const TplCache& tc, const LevelTemplate& L, int n, int block, int nblocks) {
  int k = 0;
  for (int i = first_point(block, nblocks); i < n && k < tc.K; i += nblocks * kLinThreads, ++k) {
    if (tc.pts != kTcNone) tc_point(tc, k) = __ldg(L.pts + i);
    VecC<C> v;
    if (tc.f[TC_I0] != kTcNone) { v.load(L.i0 + (size_t) i * kStride<C>); tc_put<C>(tc, k, TC_I0, v); }
    if (tc.f[TC_GX] != kTcNone) { v.load(L.gx + (size_t) i * kStride<C>); tc_put<C>(tc, k, TC_GX, v); }
    if (tc.f[TC_GY] != kTcNone) { v.load(L.gy + (size_t) i * kStride<C>); tc_put<C>(tc, k, TC_GY, v); }
  }
}

// fine-grained in-kernel profile (debug aid, compile with -DBP_FINE_PROFILE): CTA 0 / thread 0 accumulates the cycles
// between consecutive marks into prof[16 + ...]
// profile counters of CTA 0 live in shared memory while the kernel runs (a global read-modify-write per mark would
// cost more than the phases being measured) and are added to the ctx's device buffer once, at the end of the launch:
// [0..63] cycle / event slots, [64] last coarse mark, [65] last fine mark, [66] fine marks enabled
__device__ __forceinline__ long long* prof_smem() { __shared__ long long s_prof[68]; return s_prof; }
#ifdef BP_FINE_PROFILE
#define BP_FINE(slot)                                                                                          \
  do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long* _s = prof_smem(); if (_s[66] == 0x5eed) { const long long _t = clock64(); _s[slot] += _t - _s[65]; _s[65] = _t; } } } while (0)
#define BP_FINE_INIT(ptr) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long* _s = prof_smem(); _s[66] = (ptr) ? 0x5eed : 0; _s[65] = clock64(); } } while (0)
#else
#define BP_FINE(slot) do { } while (0)
#define BP_FINE_INIT(ptr) do { } while (0)
#endif

struct Sel {            // radix-select bookkeeping of one linearize (global, written by CTA 0)
  unsigned n;           // number of valid residuals (C * valid points)
  unsigned b1[2], rem1[2];
  unsigned b2[2], rem2[2];
};

struct LinShared {      // static shared memory of the linearize phases
  unsigned hist[2 * kHist2Bins];          // 16 KB, reused by every phase
  double red[kLinThreads / 32][kPartialStride];
  double xch[kLinThreads / 32][kPartialStride];   // staging of exchange_sums() (separate from red: no barrier between P4's CTA sum and the exchange)
  unsigned scan[kLinThreads / 32];
  unsigned found[8];
};

// ---------------------------------------------------------------------------------------------
// CTA-wide: locate the bins holding ranks ra and rb in a histogram of NB bins (NB multiple of blockDim).
// Returns (bin, rank inside bin) for both, and the total count.  All threads get the results.
// ---------------------------------------------------------------------------------------------
// (counts already in registers: thread t holds bins [t * PER, t * PER + PER))
template <int PER>
__device__ __forceinline__ void block_find2_regs(const unsigned (&loc)[PER], unsigned ra, unsigned rb, LinShared& sh,
                                                 unsigned& bin_a, unsigned& rem_a, unsigned& bin_b, unsigned& rem_b, unsigned& total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) sum += loc[k];
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  __syncthreads();                       // protects sh.scan / sh.found reuse across calls
  if (lane == 31) sh.scan[warp] = incl;
  if (tid < 8) sh.found[tid] = 0;
  __syncthreads();
  unsigned woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kLinThreads / 32; ++w) { const unsigned v = sh.scan[w]; if (w < warp) woff += v; tot += v; }
  unsigned excl = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const unsigned lo = excl, hi = excl + loc[k];
    if (ra >= lo && ra < hi) { sh.found[0] = tid * PER + k; sh.found[1] = ra - lo; }
    if (rb >= lo && rb < hi) { sh.found[2] = tid * PER + k; sh.found[3] = rb - lo; }
    excl = hi;
  }
  __syncthreads();
  bin_a = sh.found[0]; rem_a = sh.found[1]; bin_b = sh.found[2]; rem_b = sh.found[3]; total = tot;
}
template <int NB>
__device__ __forceinline__ void block_find2(const unsigned* __restrict__ hist, unsigned ra, unsigned rb, LinShared& sh,
                                            unsigned& bin_a, unsigned& rem_a, unsigned& bin_b, unsigned& rem_b, unsigned& total) {
  constexpr int PER = (NB + kLinThreads - 1) / kLinThreads;
  unsigned loc[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) { const int b = threadIdx.x * PER + k; loc[k] = (b < NB) ? hist[b] : 0u; }
  block_find2_regs<PER>(loc, ra, rb, sh, bin_a, rem_a, bin_b, rem_b, total);
}

// projection matrix P = K * T[0:3,:] in fp32 (rigid_body_warp.h:111-114), column-major 3x4.  Every product and sum is rounded
// on its own, on the host AND on the device (no FMA contraction: the reference is built with -mavx, without -mfma) -- the device
// loop computes P itself and must get the bits the reference's dense 3x3 * 3x4 product gets.
__host__ __device__ __forceinline__ float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
__host__ __device__ __forceinline__ float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
__host__ __device__ __forceinline__ void make_projection(const LevelTemplate& L, const M44& T, float P[12]) {
  for (int j = 0; j < 4; ++j) {
    // K = [fx 0 cx; 0 fy cy; 0 0 1]; the zero terms are kept so that the rounding matches a dense 3x3 * 3x4 product
    const float r0 = add_rn(add_rn(mul_rn(L.fx, T(0, j)), mul_rn(0.0f, T(1, j))), mul_rn(L.cx, T(2, j)));
    const float r1 = add_rn(add_rn(mul_rn(0.0f, T(0, j)), mul_rn(L.fy, T(1, j))), mul_rn(L.cy, T(2, j)));
    const float r2 = add_rn(add_rn(mul_rn(0.0f, T(0, j)), mul_rn(0.0f, T(1, j))), mul_rn(1.0f, T(2, j)));
    P[j * 3 + 0] = r0; P[j * 3 + 1] = r1; P[j * 3 + 2] = r2;
  }
}

// ---------------------------------------------------------------------------------------------
// P1: residuals (+ level-1 histogram when the scale is to be re-estimated)
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// The other InterpolationTypes of PhotoError::Impl::run (photo_error.cc:391-444): cosine, cubic (A = -0.5) and cubic
// Hermite.  Unlike the linear case the reference evaluates them in FLOAT (coefficients from float(xf), float dot
// products), with a few double sub-expressions (cos(x * M_PI) / 2.0, the Hermite tangents' / 2.0); every operation
// is spelled with its own rounding (no FMA contraction), sums in the unvectorised left-to-right order of a fixed-size
// Eigen dot product.  Quirks kept: the cubic / Hermite footprint starts AT xi (not xi - 1) in x and at yi - 1 in y
// (:413-416, :430-433).  Cold path (kLinear is the default and the benchmarked one): kept out of line.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void coeffs_cubic(float x, float c[4]) {          // interpolateCubic<float>, photo_error.cc:267-279
  const float A = -0.5f;
  const float x1 = __fadd_rn(x, 1.0f);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.0f * A), x1), 8.0f * A), x1), 4.0f * A);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f), x), x), 1.0f);
  const float y = __fsub_rn(1.0f, x);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, y), A + 3.0f), y), y), 1.0f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}
__device__ __forceinline__ void coeffs_cosine(float x, float c[2]) {         // interpolateCosine<float>, :281-290 (double inside)
  const double m = __ddiv_rn(__dsub_rn(1.0, cos(__dmul_rn((double) x, 3.14159265358979323846))), 2.0);
  c[0] = (float) __dsub_rn(1.0, m); c[1] = (float) m;
}
__device__ __forceinline__ float hermite1(float y0, float y1, float y2, float y3, float mu) {   // interpolateCubicHermite(y, mu), :311-334
  const float mu2 = __fmul_rn(mu, mu), mu3 = __fmul_rn(mu, mu2);
  const float m0 = (float) __dadd_rn(__ddiv_rn((double) __fsub_rn(y1, y0), 2.0), __ddiv_rn((double) __fsub_rn(y2, y1), 2.0));
  const float m1 = (float) __dadd_rn(__ddiv_rn((double) __fsub_rn(y2, y1), 2.0), __ddiv_rn((double) __fsub_rn(y3, y2), 2.0));
  const float a0 = __fadd_rn(__fsub_rn(__fmul_rn(2.0f, mu3), __fmul_rn(3.0f, mu2)), 1.0f);
  const float a1 = __fadd_rn(__fsub_rn(mu3, __fmul_rn(2.0f, mu2)), mu);
  const float a2 = __fsub_rn(mu3, mu2);
  const float a3 = __fadd_rn(__fmul_rn(-2.0f, mu3), __fmul_rn(3.0f, mu2));
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a0, y1), __fmul_rn(a1, m0)), __fmul_rn(a2, m1)), __fmul_rn(a3, y2));
}
template <int C>
__device__ __noinline__ void sample_nonlinear(int interp, const float* __restrict__ desc, int cols, int xi, int yi, float xf, float yf,
                                              const float* __restrict__ i0, float* __restrict__ r) {
  if (interp == 1) {                                     // kCosine: 2 x 2 footprint at (xi, yi)
    float Cx[2], Cy[2];
    coeffs_cosine(xf, Cx); coeffs_cosine(yf, Cy);
    const float* p1 = desc + ((size_t) yi * cols + xi) * kStride<C>;
    const float* p2 = p1 + (size_t) cols * kStride<C>;
    for (int c = 0; c < C; ++c) {
      const float d1 = __fadd_rn(__fmul_rn(__ldg(p1 + c), Cx[0]), __fmul_rn(__ldg(p1 + kStride<C> + c), Cx[1]));
      const float d2 = __fadd_rn(__fmul_rn(__ldg(p2 + c), Cx[0]), __fmul_rn(__ldg(p2 + kStride<C> + c), Cx[1]));
      r[c] = __fsub_rn(__fadd_rn(__fmul_rn(Cy[0], d1), __fmul_rn(Cy[1], d2)), i0[c]);
    }
  } else if (interp == 2) {                              // kCubic: rows yi-1 .. yi+2, columns xi .. xi+3
    float Cx[4], Cy[4];
    coeffs_cubic(xf, Cx); coeffs_cubic(yf, Cy);
    for (int c = 0; c < C; ++c) {
      float Iw = 0.0f;
      for (int k = 0; k < 4; ++k) {
        const float* p = desc + ((size_t) (yi - 1 + k) * cols + xi) * kStride<C> + c;
        const float d = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__ldg(p), Cx[0]), __fmul_rn(__ldg(p + kStride<C>), Cx[1])), __fmul_rn(__ldg(p + 2 * kStride<C>), Cx[2])), __fmul_rn(__ldg(p + 3 * kStride<C>), Cx[3]));
        Iw = (k == 0) ? __fmul_rn(Cy[0], d) : __fadd_rn(Iw, __fmul_rn(Cy[k], d));
      }
      r[c] = __fsub_rn(Iw, i0[c]);
    }
  } else {                                               // kCubicHermite: same footprint
    for (int c = 0; c < C; ++c) {
      float V[4];
      for (int k = 0; k < 4; ++k) {
        const float* p = desc + ((size_t) (yi - 1 + k) * cols + xi) * kStride<C> + c;
        V[k] = hermite1(__ldg(p), __ldg(p + kStride<C>), __ldg(p + 2 * kStride<C>), __ldg(p + 3 * kStride<C>), xf);
      }
      r[c] = __fsub_rn(hermite1(V[0], V[1], V[2], V[3], yf), i0[c]);
    }
  }
}

struct Bracket {       // median bracket carried from the previous GN iteration (on-device loop only)
  bool on;
  float lo, hi;        // candidates are the valid |r| with lo <= |r| <= hi
  float inv_w;         // kSelBins / (hi - lo): candidates are also counted by linear bin over the bracket
};
__device__ __forceinline__ int sel_bin(float v, float lo, float inv_w) { return min(kSelBins - 1, (int) ((v - lo) * inv_w)); }

// BLEND: arithmetic of the bilinear blend for C = 8 (bit-planes, values in [0, 1]).
//   0  fp64 with separate roundings = the reference's double expression, bit for bit (photo_error.cc:381-388): 32 F2F.F64.F32 +
//      8 F2F.F32.F64 conversions and ~90 FP64 operations per point -- the conversion pipe runs at 14-15 lanes/clk/SM and bounds
//      the phase (profiles/r1_micro_pipes.txt);
//   1  projection, Floor and validity stay fp64 (same taps, same valid flags as the reference), the fractions are rounded to fp32
//      once and the 4-tap blend runs as fp32 FMAs: |r - r_reference| <= 2e-7 on [0, 1] data against the north star's 1e-5.
// C = 1 (intensity, values up to 255) always takes the fp64 expression: 5 conversions per point, nothing to gain.
// NB: points a thread has in flight on fully streaming levels (nothing of the level fits the shared-memory cache).  2 = the
// loop handles two points per pass -- both projections, then all eight tap loads and both I0 loads, then both blends: twice the
// bytes in flight per thread on levels where the phase is bound by memory latency at 8 warps per SM (1080p dense level 0).
// The arithmetic per point is the same expression sequence: results are bit-identical to NB = 1.
template <int C, int BLEND, int NB = 1>
__device__ __forceinline__ void phase_residuals(const LevelTemplate& L, const LevelImage& I, const float* P, const Work& W,
                                                unsigned* __restrict__ hist1, bool do_hist, Bracket br, const TplCache& tc,
                                                const TemplateMeta& m, unsigned* scratch, LinShared& sh, int block, int nblocks, int interp = 0,
                                                unsigned msg_seq = 0) {
  const int tid = threadIdx.x;
  const bool do_hist1 = do_hist && !br.on;   // with a bracket the level-1 histogram is only built (phase_hist1) if the bracket misses
  if (do_hist1) { for (int b = tid; b < kHist1Bins; b += kLinThreads) sh.hist[b] = 0; }
  if (tid < 4) sh.found[4 + tid] = 0;        // CTA-level counters: [4] valid points, [5] below bracket, [6] candidates of this CTA
  float* cta_cand = reinterpret_cast<float*>(scratch);    // CTA-local candidate list (bracket on)
  __syncthreads();
  BP_FINE(16);
  unsigned cnt_valid = 0, cnt_below = 0;
  const double P00 = P[0], P10 = P[1], P20 = P[2], P01 = P[3], P11 = P[4], P21 = P[5],
               P02 = P[6], P12 = P[7], P22 = P[8], P03 = P[9], P13 = P[10], P23 = P[11];
  const int cols = I.cols, rows = I.rows;
  const int border_lo = (interp == 0 || interp == 1) ? 0 : 1, border_hi = (interp == 0 || interp == 1) ? 1 : 3;
  const int n_pts = m.n;
  int my_first = 0x7fffffff;
  int k = 0;
  // look-ahead only where NOTHING of the level fits the cache (K of several dozen): on partially cached, L2-resident levels the
  // extra instructions cost more than they hide (same-box A/B: KITTI dense level 0 32.8 -> 36.3 us with them, 1080p level 0 206.6 -> 187.8)
  const bool streaming = BP_PREFETCH && tc.K > 1 && tc.pts == kTcNone && tc.f[TC_R] == kTcNone;
  const int stride = nblocks * kLinThreads;
  float4 Xnext = make_float4(0.0f, 0.0f, 1.0f, 1.0f);
  int i_start = first_point(block, nblocks);
  if (NB == 2 && streaming && interp == 0) {
    // ---- two points per pass (see NB above); the tail (at most one point) and every other case take the loop below ----
    auto project = [&](const float4& X, int& xi, int& yi, double& xf, double& yf) -> bool {
      const double X0 = X.x, X1 = X.y, X2 = X.z, X3 = X.w;
      const double h0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P00, X0), __dmul_rn(P01, X1)), __dmul_rn(P02, X2)), __dmul_rn(P03, X3));
      const double h1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P10, X0), __dmul_rn(P11, X1)), __dmul_rn(P12, X2)), __dmul_rn(P13, X3));
      const double h2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P20, X0), __dmul_rn(P21, X1)), __dmul_rn(P22, X2)), __dmul_rn(P23, X3));
      const double w = __ddiv_rn(1.0, h2);
      const double u = __dmul_rn(w, h0), v = __dmul_rn(w, h1);
      bool ok = isfinite(u) && isfinite(v) && fabs(u) < 1e9 && fabs(v) < 1e9;
      xi = 0; yi = 0; xf = 0.0; yf = 0.0;
      if (ok) {
        xi = (int) u; xi -= (xi > u);
        yi = (int) v; yi -= (yi > v);
        ok = xi >= border_lo && xi < cols - border_hi && yi >= border_lo && yi < rows - 1;
        xf = __dsub_rn(u, (double) xi); yf = __dsub_rn(v, (double) yi);
      }
      return ok;
    };
    auto prefetch_taps = [&](const float4& X) {      // fp32 estimate of a later point's projection, only to pull its taps into the L2
      const float h2 = P[2] * X.x + P[5] * X.y + P[8] * X.z + P[11];
      const float iw = __frcp_rn(h2);
      const float uu = (P[0] * X.x + P[3] * X.y + P[6] * X.z + P[9]) * iw, vv = (P[1] * X.x + P[4] * X.y + P[7] * X.z + P[10]) * iw;
      if (uu >= 0.0f && vv >= 0.0f && uu < (float) (cols - 1) && vv < (float) (rows - 1)) {
        const float* tp = I.desc + ((size_t) (int) vv * cols + (int) uu) * kStride<C>;
        prefetch_l2(tp); prefetch_l2(tp + kStride<C>); prefetch_l2(tp + (size_t) cols * kStride<C>); prefetch_l2(tp + (size_t) cols * kStride<C> + kStride<C>);
      }
    };
    int i = i_start;
    float4 Xa = make_float4(0.0f, 0.0f, 1.0f, 1.0f), Xb = Xa;
    if (i + stride < n_pts) { Xa = __ldg(L.pts + i); Xb = __ldg(L.pts + i + stride); }
    for (; i + stride < n_pts; i += 2 * stride, k += 2) {
      const int ia = i, ib = i + stride, ja = i + 2 * stride, jb = i + 3 * stride;
      const float4 Xa0 = Xa, Xb0 = Xb;
      if (jb < n_pts) {                       // next pass: its points into registers now, its taps and I0 into the L2
        Xa = __ldg(L.pts + ja); Xb = __ldg(L.pts + jb);
        prefetch_l2(L.i0 + (size_t) ja * kStride<C>); prefetch_l2(L.i0 + (size_t) jb * kStride<C>);
        const int fa = i + 4 * stride, fb = i + 5 * stride;
        if (fb < n_pts) { prefetch_l2(L.pts + fa); prefetch_l2(L.pts + fb); }
      }
      int xia, yia, xib, yib; double xfa, yfa, xfb, yfb;
      const bool oka = project(Xa0, xia, yia, xfa, yfa), okb = project(Xb0, xib, yib, xfb, yfb);
      VecC<C> ta[4], tb[4], i0a, i0b;
      if (oka) {
        const float* tap = I.desc + ((size_t) yia * cols + xia) * kStride<C>;
        ta[0].load(tap); ta[1].load(tap + kStride<C>); ta[2].load(tap + (size_t) cols * kStride<C>); ta[3].load(tap + (size_t) cols * kStride<C> + kStride<C>);
        i0a.load(L.i0 + (size_t) ia * kStride<C>);
      }
      if (okb) {
        const float* tap = I.desc + ((size_t) yib * cols + xib) * kStride<C>;
        tb[0].load(tap); tb[1].load(tap + kStride<C>); tb[2].load(tap + (size_t) cols * kStride<C>); tb[3].load(tap + (size_t) cols * kStride<C> + kStride<C>);
        i0b.load(L.i0 + (size_t) ib * kStride<C>);
      }
      if (jb < n_pts) { prefetch_taps(Xa); prefetch_taps(Xb); }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool ok = h ? okb : oka;
        const int ii = h ? ib : ia;
        const double xf = h ? xfb : xfa, yf = h ? yfb : yfa;
        VecC<C> r;
        if (ok) {
          const double wx = __dsub_rn(1.0, xf), wy = __dsub_rn(1.0, yf);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float t00 = h ? tb[0].v[c] : ta[0].v[c], t01 = h ? tb[1].v[c] : ta[1].v[c], t10 = h ? tb[2].v[c] : ta[2].v[c], t11 = h ? tb[3].v[c] : ta[3].v[c];
            const float i0v = h ? i0b.v[c] : i0a.v[c];
            if (BLEND == 1 && C == 8) {
              const float xf32 = (float) xf, yf32 = (float) yf, wx32 = (float) wx, wy32 = (float) wy;
              const float top = fmaf(t01, xf32, t00 * wx32);
              const float bot = fmaf(t11, xf32, t10 * wx32);
              r.v[c] = fmaf(yf32, bot, wy32 * top) - i0v;
            } else {
              const double top = __dadd_rn(__dmul_rn((double) t00, wx), __dmul_rn((double) t01, xf));
              const double bot = __dadd_rn(__dmul_rn((double) t10, wx), __dmul_rn((double) t11, xf));
              const double Iw = __dadd_rn(__dmul_rn(wy, top), __dmul_rn(yf, bot));
              r.v[c] = (float) __dsub_rn(Iw, (double) i0v);
            }
          }
          if (do_hist) {
            if (do_hist1) {
#pragma unroll
              for (int c = 0; c < C; ++c) atomicAdd(&sh.hist[__float_as_uint(fabsf(r.v[c])) >> 20], 1u);
            }
            my_first = min(my_first, ii);
            ++cnt_valid;
            if (br.on) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float av = fabsf(r.v[c]);
                cnt_below += (av < br.lo) ? 1u : 0u;
                if (av >= br.lo && av <= br.hi) {
                  const unsigned slot = atomicAdd(&sh.found[6], 1u);
                  if (slot < (unsigned) kCtaCandCap) cta_cand[slot] = av;
                  atomicAdd(hist1 + kHistBins + 8 + sel_bin(av, br.lo, br.inv_w), 1u);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) r.v[c] = 0.0f;
        }
        r.store(W.res + (size_t) ii * kStride<C>); W.valid[ii] = ok ? 1 : 0;      // streaming level: residuals live in global memory
      }
    }
    i_start = i;
  }
  if (streaming) { const int i0 = i_start; if (i0 < n_pts) Xnext = (tc.pts != kTcNone) ? tc_point(tc, k) : __ldg(L.pts + i0); }
  for (int i = i_start; i < n_pts; i += stride, ++k) {
    float4 X;
    if (streaming) {
      X = Xnext;
      const int i1 = i + stride, i2 = i + 2 * stride;
      if (i2 < n_pts) { if (tc.pts == kTcNone) prefetch_l2(L.pts + i2); if (tc.f[TC_I0] == kTcNone) prefetch_l2(L.i0 + (size_t) i2 * kStride<C>); }
      if (i1 < n_pts) {
        Xnext = (tc.pts != kTcNone) ? tc_point(tc, k + 1) : __ldg(L.pts + i1);
        // fp32 estimate of the next point's projection, only to prefetch its taps (a wrong guess costs nothing but the prefetch)
        const float h2 = P[2] * Xnext.x + P[5] * Xnext.y + P[8] * Xnext.z + P[11];
        const float iw = __frcp_rn(h2);
        const float uu = (P[0] * Xnext.x + P[3] * Xnext.y + P[6] * Xnext.z + P[9]) * iw, vv = (P[1] * Xnext.x + P[4] * Xnext.y + P[7] * Xnext.z + P[10]) * iw;
        if (uu >= 0.0f && vv >= 0.0f && uu < (float) (cols - 1) && vv < (float) (rows - 1)) {
          const float* tp = I.desc + ((size_t) (int) vv * cols + (int) uu) * kStride<C>;
          prefetch_l2(tp); prefetch_l2(tp + kStride<C>); prefetch_l2(tp + (size_t) cols * kStride<C>); prefetch_l2(tp + (size_t) cols * kStride<C> + kStride<C>);
        }
      }
    } else {
      X = (tc.pts != kTcNone) ? tc_point(tc, k) : __ldg(L.pts + i);
    }
    const double X0 = X.x, X1 = X.y, X2 = X.z, X3 = X.w;
    const double h0 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P00, X0), __dmul_rn(P01, X1)), __dmul_rn(P02, X2)), __dmul_rn(P03, X3));
    const double h1 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P10, X0), __dmul_rn(P11, X1)), __dmul_rn(P12, X2)), __dmul_rn(P13, X3));
    const double h2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P20, X0), __dmul_rn(P21, X1)), __dmul_rn(P22, X2)), __dmul_rn(P23, X3));
    const double w = __ddiv_rn(1.0, h2);
    const double u = __dmul_rn(w, h0), v = __dmul_rn(w, h1);
    bool ok = isfinite(u) && isfinite(v) && fabs(u) < 1e9 && fabs(v) < 1e9;
    int xi = 0, yi = 0;
    if (ok) {
      xi = (int) u; xi -= (xi > u);      // Floor(double), photo_error.cc:261-265
      yi = (int) v; yi -= (yi > v);
      ok = xi >= border_lo && xi < cols - border_hi && yi >= border_lo && yi < rows - 1;     // photo_error.cc:346-358 (y bound is always rows - 1)
    }
    VecC<C> r;
    if (ok) {
      const double xf = __dsub_rn(u, (double) xi), yf = __dsub_rn(v, (double) yi);
      const double wx = __dsub_rn(1.0, xf), wy = __dsub_rn(1.0, yf);
      VecC<C> i0;
      if (tc.f[TC_I0] != kTcNone) tc_get<C>(tc, k, TC_I0, i0); else i0.load(L.i0 + (size_t) i * kStride<C>);
      if (interp == 0) {
        const float* tap = I.desc + ((size_t) yi * cols + xi) * kStride<C>;
        VecC<C> t00, t01, t10, t11;
        t00.load(tap); t01.load(tap + kStride<C>); t10.load(tap + (size_t) cols * kStride<C>); t11.load(tap + (size_t) cols * kStride<C> + kStride<C>);
        if (BLEND == 1 && C == 8) {
          const float xf32 = (float) xf, yf32 = (float) yf, wx32 = (float) wx, wy32 = (float) wy;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float top = fmaf(t01.v[c], xf32, t00.v[c] * wx32);
            const float bot = fmaf(t11.v[c], xf32, t10.v[c] * wx32);
            r.v[c] = fmaf(yf32, bot, wy32 * top) - i0.v[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const double top = __dadd_rn(__dmul_rn((double) t00.v[c], wx), __dmul_rn((double) t01.v[c], xf));
            const double bot = __dadd_rn(__dmul_rn((double) t10.v[c], wx), __dmul_rn((double) t11.v[c], xf));
            const double Iw = __dadd_rn(__dmul_rn(wy, top), __dmul_rn(yf, bot));
            r.v[c] = (float) __dsub_rn(Iw, (double) i0.v[c]);
          }
        }
      } else {
        float i0l[C], rl[C];                 // only these copies live in local memory (the out-of-line call takes addresses)
#pragma unroll
        for (int c = 0; c < C; ++c) i0l[c] = i0.v[c];
        sample_nonlinear<C>(interp, I.desc, cols, xi, yi, (float) xf, (float) yf, i0l, rl);
#pragma unroll
        for (int c = 0; c < C; ++c) r.v[c] = rl[c];
      }
      if (do_hist) {
        if (do_hist1) {
#pragma unroll
          for (int c = 0; c < C; ++c) atomicAdd(&sh.hist[__float_as_uint(fabsf(r.v[c])) >> 20], 1u);
        }
        my_first = min(my_first, i);
        ++cnt_valid;
        if (br.on) {
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float a = fabsf(r.v[c]);
            cnt_below += (a < br.lo) ? 1u : 0u;
            if (a >= br.lo && a <= br.hi) {
              const unsigned slot = atomicAdd(&sh.found[6], 1u);        // shared-memory append; overflow is only counted
              if (slot < (unsigned) kCtaCandCap) cta_cand[slot] = a;
              atomicAdd(hist1 + kHistBins + 8 + sel_bin(a, br.lo, br.inv_w), 1u);      // global bracket histogram (fire and forget)
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) r.v[c] = 0.0f;
    }
    // residuals / valid flags stay in shared memory while the level is cached (C = 8: written back once, after the last
    // iteration of the finest level); C = 1 keeps the global copy for the n < 3 median rule of finish_scale()
    if (!(tc.f[TC_R] != kTcNone && C != 1)) { r.store(W.res + (size_t) i * kStride<C>); W.valid[i] = ok ? 1 : 0; }
    if (tc.f[TC_R] != kTcNone) { tc_put<C>(tc, k, TC_R, r); tc_valid(tc, k) = ok ? 1 : 0; }
  }
  BP_FINE(17);
  if (do_hist) {
    __syncthreads();
    BP_FINE(18);
    if (do_hist1) { for (int b = tid; b < kHist1Bins; b += kLinThreads) { const unsigned v = sh.hist[b]; if (v) atomicAdd(hist1 + b, v); } }
    my_first = __reduce_min_sync(0xffffffffu, my_first);        // REDUX: one instruction per warp-wide integer reduction
    cnt_valid = __reduce_add_sync(0xffffffffu, cnt_valid);
    cnt_below = __reduce_add_sync(0xffffffffu, cnt_below);
    if ((tid & 31) == 0 && my_first != 0x7fffffff) {
      atomicMax(hist1 + kHistBins, ~(unsigned) my_first);   // word zero-initialised
      atomicAdd(&sh.found[4], cnt_valid);
      if (cnt_below) atomicAdd(&sh.found[5], cnt_below);
    }
    __syncthreads();
    if (tid == 0) {
      if (sh.found[4]) atomicAdd(hist1 + kHistBins + 1, sh.found[4]);
      if (sh.found[5]) atomicAdd(hist1 + kHistBins + 2, sh.found[5]);
      // candidate count; a CTA whose fixed region overflowed poisons it so that the select falls back to the radix path
      // (correctness never depends on the bracket).  No return value is needed: nothing here waits for the L2.
      const unsigned nc = sh.found[6];
      if (nc) atomicAdd(hist1 + kHistBins + 3, nc > (unsigned) kCtaCandCap ? kCandPoison : nc);
    }
    if (br.on) {
      const unsigned nc = sh.found[6];
      if (msg_seq && tid < kMsgWords) {       // the same facts as ONE self-validating 128-byte message (persistent kernel: no barrier needed to read it)
        unsigned lo, hi;
        if (tid == 0) { lo = sh.found[4]; hi = sh.found[5]; }
        else if (tid == 1) { lo = nc; hi = (nc > 0u && nc <= (unsigned) kMsgCand) ? __float_as_uint(cta_cand[0]) : 0xbf800000u; }
        else {
          const unsigned i0 = 2u * tid - 3u, i1 = 2u * tid - 2u;
          lo = (i0 < nc && nc <= (unsigned) kMsgCand) ? __float_as_uint(cta_cand[i0]) : 0xbf800000u;      // -1.0f = empty
          hi = (i1 < nc && nc <= (unsigned) kMsgCand) ? __float_as_uint(cta_cand[i1]) : 0xbf800000u;
        }
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(W.msg + ((size_t) (msg_seq & 1u) * kMaxGrid + block) * kMsgWords + tid),
                     "r"(lo), "r"(msg_seq), "r"(hi), "r"(msg_seq) : "memory");
      }
      if (tid < kCandPerCta)                  // this CTA's region of the candidate buffer: values, then -1 = empty
        W.cand[(size_t) block * kCandPerCta + tid] = ((unsigned) tid < nc) ? cta_cand[tid] : -1.0f;
      if (nc > (unsigned) kCandPerCta && nc <= (unsigned) kCtaCandCap) {       // rare (wide bracket): the rest goes to the shared overflow list
        if (tid == 0) {
          const unsigned extra = nc - kCandPerCta, base = atomicAdd(hist1 + kHistBins + 4, extra);
          if (base + extra > (unsigned) kOvfCap) atomicAdd(hist1 + kHistBins + 3, kCandPoison);
          sh.found[7] = base;
        }
        __syncthreads();
        const unsigned base = sh.found[7];
        float* ovf = W.cand + (size_t) nblocks * kCandPerCta;
        for (unsigned j = tid; j < nc - kCandPerCta; j += kLinThreads) if (base + j < (unsigned) kOvfCap) ovf[base + j] = cta_cand[kCandPerCta + j];
      }
    }
  }
  BP_FINE(19);
}

// One pass over this CTA's residuals for the select phases.  Residuals in the shared-memory cache: one point per step.
// Residuals in global memory (streaming levels, host-driven kernels): FOUR points per step, flags and vectors loaded before any
// is used -- at 8 warps per SM a one-load-per-step loop is bound by memory latency (1080p dense level 0: ~110 us per pass
// against ~15 us of bandwidth time).  f(r) sees the residual vector of every valid point; the order does not matter (counting).
// (MLP = 1 keeps the one-point loop everywhere: the persistent kernel of the cached workloads sits at the register limit.)
template <int C, int MLP, class F>
__device__ __forceinline__ void for_each_valid_residual(const Work& W, const TplCache& tc, int n_pts, int block, int nblocks, F&& f) {
  const int stride = nblocks * kLinThreads;
  if (MLP == 1) {
    int k = 0;
    for (int i = first_point(block, nblocks); i < n_pts; i += stride, ++k) {
      if (!((tc.f[TC_R] != kTcNone) ? tc_valid(tc, k) : W.valid[i])) continue;
      VecC<C> r;
      if (tc.f[TC_R] != kTcNone) tc_get<C>(tc, k, TC_R, r); else r.load_plain(W.res + (size_t) i * kStride<C>);
      f(r);
    }
    return;
  }
  if (tc.f[TC_R] != kTcNone) {
    int k = 0;
    for (int i = first_point(block, nblocks); i < n_pts; i += stride, ++k) {
      if (!tc_valid(tc, k)) continue;
      VecC<C> r; tc_get<C>(tc, k, TC_R, r);
      f(r);
    }
    return;
  }
  for (int i = first_point(block, nblocks); i < n_pts; i += 4 * stride) {
    bool v[4]; VecC<C> r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int j = i + q * stride; v[q] = j < n_pts && W.valid[j] != 0; }
#pragma unroll
    for (int q = 0; q < 4; ++q) if (v[q]) r[q].load_plain(W.res + (size_t) (i + q * stride) * kStride<C>);
#pragma unroll
    for (int q = 0; q < 4; ++q) if (v[q]) f(r[q]);
  }
}

// level-1 histogram of |r| as a separate pass (on-device loop, only after a bracket miss)
template <int C, int MLP = 1>
__device__ __forceinline__ void phase_hist1(const Work& W, unsigned* __restrict__ hist1, const TplCache& tc, const TemplateMeta& m,
                                            LinShared& sh, int block, int nblocks) {
  const int tid = threadIdx.x;
  for (int b = tid; b < kHist1Bins; b += kLinThreads) sh.hist[b] = 0;
  __syncthreads();
  for_each_valid_residual<C, MLP>(W, tc, m.n, block, nblocks, [&](const VecC<C>& r) {
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(&sh.hist[__float_as_uint(fabsf(r.v[c])) >> 20], 1u);
  });
  __syncthreads();
  for (int b = tid; b < kHist1Bins; b += kLinThreads) { const unsigned v = sh.hist[b]; if (v) atomicAdd(hist1 + b, v); }
}

// ---------------------------------------------------------------------------------------------
// P2 / P3: refine the radix select by 11 then 9 more bits.  level == 2: input hist1, output hist2 sets;
// level == 3: input hist2 sets, output hist3 sets.  slot 0 = rank n/2-1 ("lo"), slot 1 = rank n/2 ("hi").
// ---------------------------------------------------------------------------------------------
template <int C, int LEVEL, int MLP = 1>
__device__ __forceinline__ void phase_select(const LevelTemplate& L, const Work& W, unsigned* __restrict__ hset, Sel* __restrict__ sel,
                                             const TplCache& tc, const TemplateMeta& m, LinShared& sh, int block, int nblocks) {
  const int tid = threadIdx.x;
  unsigned* hist1 = hset;
  unsigned* hist2 = hset + kHist1Bins;               // [2][kHist2Bins]
  unsigned* hist3 = hset + kHist1Bins + 2 * kHist2Bins;   // [2][kHist3Bins]
  unsigned pa, pb;                                   // prefixes to match
  constexpr int NBOUT = (LEVEL == 2) ? kHist2Bins : kHist3Bins;
  if (LEVEL == 2) {
    unsigned ba, ra, bb, rb, n;
    // total first (ranks depend on n): scan with dummy ranks, then the real ones
    block_find2<kHist1Bins>(hist1, 0u, 0u, sh, ba, ra, bb, rb, n);
    const unsigned t_hi = n / 2, t_lo = (n >= 1) ? ((n / 2) - ((n % 2 == 0 && n >= 2) ? 1u : 0u)) : 0u;
    block_find2<kHist1Bins>(hist1, t_lo, t_hi, sh, ba, ra, bb, rb, n);
    pa = ba; pb = bb;
    if (block == 0 && tid == 0) { sel->n = n; sel->b1[0] = ba; sel->rem1[0] = ra; sel->b1[1] = bb; sel->rem1[1] = rb; }
    if (n < 3) return;                               // median rule for tiny n needs no refinement
  } else {
    const unsigned n = sel->n;
    if (n < 3) return;
    unsigned ba, ra, bb, rb, tot;
    block_find2<kHist2Bins>(hist2, sel->rem1[0], 0xffffffffu, sh, ba, ra, bb, rb, tot);
    unsigned bb2, rb2, dummy0, dummy1;
    block_find2<kHist2Bins>(hist2 + kHist2Bins, sel->rem1[1], 0xffffffffu, sh, bb2, rb2, dummy0, dummy1, tot);
    if (block == 0 && tid == 0) { sel->b2[0] = ba; sel->rem2[0] = ra; sel->b2[1] = bb2; sel->rem2[1] = rb2; }
    pa = (sel->b1[0] << 11) | ba; pb = (sel->b1[1] << 11) | bb2;
  }
  for (int b = tid; b < 2 * NBOUT; b += kLinThreads) sh.hist[b] = 0;
  __syncthreads();
  constexpr int SHIFT_MATCH = (LEVEL == 2) ? 20 : 9;
  constexpr int SHIFT_BIN = (LEVEL == 2) ? 9 : 0;
  for_each_valid_residual<C, MLP>(W, tc, m.n, block, nblocks, [&](const VecC<C>& r) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const unsigned bits = __float_as_uint(fabsf(r.v[c]));
      const unsigned pre = bits >> SHIFT_MATCH, bin = (bits >> SHIFT_BIN) & (NBOUT - 1);
      if (pre == pa) atomicAdd(&sh.hist[bin], 1u);
      if (pre == pb) atomicAdd(&sh.hist[NBOUT + bin], 1u);
    }
  });
  __syncthreads();
  unsigned* out = (LEVEL == 2) ? hist2 : hist3;
  for (int b = tid; b < 2 * NBOUT; b += kLinThreads) { const unsigned v = sh.hist[b]; if (v) atomicAdd(out + b, v); }
}

// scale = 1.4826 (1 + 5/(n-6)) median with the reference's size_t arithmetic and the "< 1e-6 -> 1" rule
__device__ __forceinline__ float scale_from_median(unsigned n, float med) {
  const float denom = (n >= 6) ? (float) (n - 6) : 1.8446744073709552e19f;     // size_t wrap-around of `size()-6`
  float s = __fmul_rn(__fmul_rn(1.4826f, __fadd_rn(1.0f, __fdiv_rn(5.0f, denom))), med);
  if ((double) s < 1e-6) s = 1.0f;
  return s;
}

// sigma from the finished select (every CTA, identical result): median rule of utils.h:224-252 and
// scale = 1.4826 (1 + 5/(n-6)) median, "< 1e-6 -> 1" (mestimator.cc:452-482)
template <int C>
__device__ __forceinline__ float finish_scale(const Work& W, const unsigned* __restrict__ hset, const Sel* __restrict__ sel, LinShared& sh,
                                              float* lo_out = nullptr, float* hi_out = nullptr) {
  const unsigned n = sel->n;
  float med, lo = 0.0f, hi = 0.0f;
  if (n == 0) {
    med = 0.0f;
  } else if (n < 3) {
    med = fabsf(W.res[(size_t) (~hset[kHistBins]) * kStride<C>]);         // data[0]: channel 0 of the first valid point
    lo = hi = med;
  } else {
    const unsigned* hist3 = hset + kHist1Bins + 2 * kHist2Bins;
    unsigned ba, ra, bb, rb, t0, d0, d1;
    block_find2<kHist3Bins>(hist3, sel->rem2[0], 0xffffffffu, sh, ba, ra, d0, d1, t0);
    block_find2<kHist3Bins>(hist3 + kHist3Bins, sel->rem2[1], 0xffffffffu, sh, bb, rb, d0, d1, t0);
    lo = __uint_as_float((sel->b1[0] << 20) | (sel->b2[0] << 9) | ba);
    hi = __uint_as_float((sel->b1[1] << 20) | (sel->b2[1] << 9) | bb);
    med = (n % 2 != 0) ? hi : (float) ((double) __fadd_rn(lo, hi) / 2.0);
    if (n % 2 != 0) lo = hi;
  }
  if (lo_out) { *lo_out = lo; *hi_out = hi; }
  return scale_from_median(n, med);
}

struct AbortCtl;
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned nblocks, AbortCtl* abort_flag);
__device__ __forceinline__ bool wait_expired(unsigned& spins, AbortCtl* ac);

// Bracketed exact median (on-device loop): P1 counted the valid residuals below the bracket and left the ones inside it
// in the per-CTA regions of W.cand.  If both middle ranks fall inside and no region overflowed, the two order statistics
// are found among the candidates (a few hundred) by linear binning + rank counting -- no further pass over the
// residuals, no further grid barrier.  The candidates stay in REGISTERS (one L2 round trip fetches all regions and the
// counters); only the histogram and the final short list touch shared memory.
// Returns false when the bracket missed (caller falls back to the 3-level radix select).
constexpr int kSelPre = (148 * kCandPerCta + kLinThreads - 1) / kLinThreads + ((149 * kCandPerCta > ((148 * kCandPerCta + kLinThreads - 1) / kLinThreads) * kLinThreads) ? 1 : 0);    // candidate slots per thread held in registers (covers 149 CTAs x kCandPerCta); more are re-read
static_assert(kSelBins == 4 * kLinThreads, "bracket_select loads the bracket histogram as one uint4 per thread");
template <int C>
__device__ __forceinline__ bool bracket_select(const Work& W, unsigned* __restrict__ hset, LinShared& sh, unsigned* scratch, int blk, int nblocks,
                                               const Bracket& br, unsigned* bar_counter, unsigned& bar_epoch, AbortCtl* abort_flag,
                                               unsigned& n_out, unsigned& ncand_out, float& lo_out, float& hi_out) {
  const int tid = threadIdx.x;
  const unsigned total = (unsigned) nblocks * kCandPerCta;
  // ONE L2 round trip: bracket histogram, counters and every CTA's candidate region are requested together
  const uint4 gb = __ldcg(reinterpret_cast<const uint4*>(hset + kHistBins + 8) + tid);
  float pre[kSelPre];
#pragma unroll
  for (int q = 0; q < kSelPre; ++q) { const unsigned j = tid + q * kLinThreads; pre[q] = (j < total) ? __ldcg(W.cand + j) : -1.0f; }
  const unsigned nv = __ldcg(hset + kHistBins + 1), below = __ldcg(hset + kHistBins + 2), ncand = __ldcg(hset + kHistBins + 3);
  const unsigned novf = __ldcg(hset + kHistBins + 4);
  const float* ovf = W.cand + (size_t) nblocks * kCandPerCta;
  const unsigned n = nv * (unsigned) C;
  n_out = n; ncand_out = ncand;
  if (n < 3) return false;
  const unsigned t_hi = n / 2, t_lo = (n % 2 == 0) ? t_hi - 1 : t_hi;
  if (ncand >= kCandPoison || novf > (unsigned) kOvfCap || below > t_lo || t_hi >= below + ncand) return false;
  const unsigned ra = t_lo - below, rb = t_hi - below;
  BP_FINE(33);
  float* list = reinterpret_cast<float*>(scratch + kCtaCandCap);   // [kSelList]
  // the linear (monotone) binning over [lo, hi] puts the wanted ranks into bins holding a handful of values
  const unsigned loc[4] = {gb.x, gb.y, gb.z, gb.w};
  unsigned bin_a, rem_a, bin_b, rem_b, tot;
  block_find2_regs<4>(loc, ra, rb, sh, bin_a, rem_a, bin_b, rem_b, tot);      // resets sh.found[0..7]
  BP_FINE(36);
  BP_FINE(44);
  unsigned match = 0;                        // branch-free membership test of the register-resident slots, then a (rare) append
#pragma unroll
  for (int q = 0; q < kSelPre; ++q) {
    const unsigned b = (unsigned) sel_bin(pre[q], br.lo, br.inv_w);
    match |= ((pre[q] >= 0.0f && (b == bin_a || b == bin_b)) ? 1u : 0u) << q;
  }
  BP_FINE(45);
  while (match) {                            // usually no bit at all, rarely more than one: one short append per set bit
    const int q = __ffs((int) match) - 1;
    match &= match - 1;
    float v = pre[0];
#pragma unroll
    for (int t = 1; t < kSelPre; ++t) v = (q == t) ? pre[t] : v;
    const unsigned slot = atomicAdd(&sh.found[4], 1u);
    if (slot < (unsigned) kSelList) list[slot] = v;
  }
  BP_FINE(46);
  for (unsigned j = tid + kSelPre * kLinThreads; j < total; j += kLinThreads) {
    const float v = __ldcg(W.cand + j);
    if (v >= 0.0f) {
      const unsigned b = (unsigned) sel_bin(v, br.lo, br.inv_w);
      if (b == bin_a || b == bin_b) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; }
    }
  }
  // Overflow list (wide brackets): short ones are scanned by every CTA.  Long ones (megapixel levels, early iterations) slice by
  // slice: CTA b tests its 1 / nblocks of the list and appends the members of the two wanted bins to a global short list; one more
  // grid barrier, then every CTA reads that short list.  All branches depend on grid-wide counters only: uniform.
  const unsigned lo_j = (novf <= (unsigned) kOvfLocalScan) ? 0u : (unsigned) (((unsigned long long) novf * blk) / nblocks);
  const unsigned hi_j = (novf <= (unsigned) kOvfLocalScan) ? novf : (unsigned) (((unsigned long long) novf * (blk + 1)) / nblocks);
  float* shortl = W.cand + (size_t) nblocks * kCandPerCta + kOvfCap;
  for (unsigned j0 = lo_j + tid; j0 < hi_j; j0 += 8 * kLinThreads) {      // 8 independent loads in flight per thread
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { const unsigned j = j0 + q * kLinThreads; v[q] = (j < hi_j) ? __ldcg(ovf + j) : -1.0f; }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const unsigned b = (unsigned) sel_bin(v[q], br.lo, br.inv_w);
      if (v[q] >= 0.0f && (b == bin_a || b == bin_b)) {
        if (novf <= (unsigned) kOvfLocalScan) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v[q]; }
        else { const unsigned slot = atomicAdd(hset + kHistBins + 5, 1u); if (slot < (unsigned) kSelList) shortl[slot] = v[q]; }
      }
    }
  }
  if (novf > (unsigned) kOvfLocalScan) {
    grid_barrier(bar_counter, bar_epoch, nblocks, abort_flag);
    const unsigned ns = __ldcg(hset + kHistBins + 5);
    if (ns > (unsigned) kSelList) return false;
    if (tid < (int) ns) { const float v = __ldcg(shortl + tid); const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; }
  }
  __syncthreads();
  BP_FINE(37);
  const unsigned nl = sh.found[4];
  if (nl > (unsigned) kSelList) return false;       // pathological pile-up of equal values: let the radix select handle it
  if (tid < (int) nl) {
    const float v = list[tid];
    const unsigned bj = (unsigned) sel_bin(v, br.lo, br.inv_w);
    unsigned rank = 0;
    for (unsigned j = 0; j < nl; ++j) {
      const float u = list[j];
      const unsigned bu = (unsigned) sel_bin(u, br.lo, br.inv_w);
      rank += (bu == bj && (u < v || (u == v && j < (unsigned) tid))) ? 1u : 0u;
    }
    if (bj == bin_a && rank == rem_a) sh.found[6] = __float_as_uint(v);
    if (bj == bin_b && rank == rem_b) sh.found[7] = __float_as_uint(v);
  }
  __syncthreads();
  BP_FINE(38);
  lo_out = __uint_as_float(sh.found[6]); hi_out = __uint_as_float(sh.found[7]);
  return true;
}

// The usual way to the exact median in the persistent kernel: NO grid barrier.  Every CTA posted, at the end of P1, one 128-byte
// flag-in-data message {valid points, residuals below the bracket, candidates inside it, up to 13 candidate values}; every CTA polls
// all messages (one L2 round trip once the last CTA has posted: data and "ready" arrive together), sums the counters, bins the
// candidates into a shared-memory histogram over the bracket and ranks the few values of the two wanted bins -- the same pair of
// order statistics as bracket_select / the radix select.  Verdicts (identical in every CTA: same messages):
//   1 = hit (lo / hi valid);  2 = the bracket missed the middle ranks -> radix select;
//   0 = some CTA had more candidates than a message holds (wide bracket) -> grid barrier + bracket_select on the global lists.
constexpr int kMsgPre = (148 * kMsgWords + kLinThreads - 1) / kLinThreads;       // message words a thread holds in registers (grids <= 148 CTAs)
template <int C>
__device__ __forceinline__ int message_select(const Work& W, unsigned seq, LinShared& sh, unsigned* scratch, int nblocks, const Bracket& br, AbortCtl* abort_flag,
                                              unsigned& n_out, unsigned& ncand_out, float& lo_out, float& hi_out) {
  const int tid = threadIdx.x;
  const uint4* base = W.msg + (size_t) (seq & 1u) * kMaxGrid * kMsgWords;
  const int total = nblocks * kMsgWords;
  unsigned plo[kMsgPre], phi[kMsgPre];
  bool ok[kMsgPre]; bool all = true;
#pragma unroll
  for (int q = 0; q < kMsgPre; ++q) { ok[q] = tid + q * kLinThreads >= total; plo[q] = phi[q] = 0xbf800000u; all = all && ok[q]; }
  for (int b = tid; b < kSelBins; b += kLinThreads) sh.hist[b] = 0;              // histogram of the candidates over the bracket (shared memory)
  if (tid < 4) sh.found[tid] = 0;                                                 // [0] valid points [1] below [2] candidates [3] overflowing CTAs  (P1's tail still reads [4..7])
  unsigned spins = 0;
  while (!all) {
    all = true;
#pragma unroll
    for (int q = 0; q < kMsgPre; ++q) {
      if (!ok[q]) {
        unsigned x, f0, y, f1;
        asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(base + tid + q * kLinThreads) : "memory");
        if (f0 == seq && f1 == seq) { ok[q] = true; plo[q] = x; phi[q] = y; }
      }
      all = all && ok[q];
    }
    if (!all && wait_expired(spins, abort_flag)) break;
  }
  __syncthreads();                                                                // histogram zeroed, counters zeroed
#pragma unroll
  for (int q = 0; q < kMsgPre; ++q) {
    const int j = tid + q * kLinThreads;
    if (j >= total) continue;
    const int k = j & (kMsgWords - 1);
    if (k == 0) { atomicAdd(&sh.found[0], plo[q]); atomicAdd(&sh.found[1], phi[q]); }
    else {
      if (k == 1) { atomicAdd(&sh.found[2], plo[q]); if (plo[q] > (unsigned) kMsgCand) atomicAdd(&sh.found[3], 1u); }
      else { const float v = __uint_as_float(plo[q]); if (v >= 0.0f) atomicAdd(&sh.hist[sel_bin(v, br.lo, br.inv_w)], 1u); }
      const float v = __uint_as_float(phi[q]); if (v >= 0.0f) atomicAdd(&sh.hist[sel_bin(v, br.lo, br.inv_w)], 1u);
    }
  }
  __syncthreads();
  const unsigned n = sh.found[0] * (unsigned) C, below = sh.found[1], ncand = sh.found[2], nover = sh.found[3];
  n_out = n; ncand_out = ncand;
  if (nover != 0u) return 0;
  if (n < 3) return 2;
  const unsigned t_hi = n / 2, t_lo = (n % 2 == 0) ? t_hi - 1 : t_hi;
  if (below > t_lo || t_hi >= below + ncand) return 2;
  const unsigned ra = t_lo - below, rb = t_hi - below;
  unsigned bin_a, rem_a, bin_b, rem_b, tot;
  {
    unsigned loc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) loc[k] = sh.hist[tid * 4 + k];
    block_find2_regs<4>(loc, ra, rb, sh, bin_a, rem_a, bin_b, rem_b, tot);        // resets sh.found[0..7]
  }
  float* list = reinterpret_cast<float*>(scratch + kCtaCandCap);   // [kSelList]
#pragma unroll
  for (int q = 0; q < kMsgPre; ++q) {
    const int j = tid + q * kLinThreads;
    if (j >= total || (j & (kMsgWords - 1)) == 0) continue;
    if ((j & (kMsgWords - 1)) != 1) {
      const float v = __uint_as_float(plo[q]);
      if (v >= 0.0f) { const unsigned b = (unsigned) sel_bin(v, br.lo, br.inv_w); if (b == bin_a || b == bin_b) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; } }
    }
    const float v = __uint_as_float(phi[q]);
    if (v >= 0.0f) { const unsigned b = (unsigned) sel_bin(v, br.lo, br.inv_w); if (b == bin_a || b == bin_b) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; } }
  }
  __syncthreads();
  const unsigned nl = sh.found[4];
  if (nl > (unsigned) kSelList) return 2;           // pathological pile-up of equal values: let the radix select handle it
  if (tid < (int) nl) {
    const float v = list[tid];
    const unsigned bj = (unsigned) sel_bin(v, br.lo, br.inv_w);
    unsigned rank = 0;
    for (unsigned j = 0; j < nl; ++j) {
      const float u = list[j];
      const unsigned bu = (unsigned) sel_bin(u, br.lo, br.inv_w);
      rank += (bu == bj && (u < v || (u == v && j < (unsigned) tid))) ? 1u : 0u;
    }
    if (bj == bin_a && rank == rem_a) sh.found[6] = __float_as_uint(v);
    if (bj == bin_b && rank == rem_b) sh.found[7] = __float_as_uint(v);
  }
  __syncthreads();
  lo_out = __uint_as_float(sh.found[6]); hi_out = __uint_as_float(sh.found[7]);
  return 1;
}

__device__ __forceinline__ float robust_weight(int loss, float r, float sigma_inv) {
  if (loss == 0x12) return 1.0f;
  const float x = __fmul_rn(r, sigma_inv);
  if (loss == 0x10) {                                    // huber_simd: k / max(|x|, k)
    const float k = 1.345f;
    return __fdiv_rn(k, fmaxf(fabsf(x), k));
  }
  const float t = 4.685f, t_i = (float) (1.0 / 4.685f);  // tukey_simd: (|x| < t) ? (1 - (x/t)^2)^2 : 0
  float q = __fmul_rn(x, t_i);
  q = __fsub_rn(1.0f, __fmul_rn(q, q));
  q = __fmul_rn(q, q);
  return (fabsf(x) < t) ? q : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// P4: weights + normal equations.  Per point: channel sums -> rank-2 update with A, B.
// Partial layout (kPartialStride doubles): [0..20] upper triangle of H row-major, [21..26] G, [27] sum w r^2,
// [28] count(w > good_threshold), [29] valid points.
// ---------------------------------------------------------------------------------------------
// Returns, in thread k < 30, the CTA total of scalar k.  write_partials: also store it to W.partials (+ fence) for the
// last-CTA fold of the host-driven kernels; the persistent kernel exchanges the totals itself (exchange_sums).
// NB = 2: on fully streaming levels two points per pass (all eight loads first, then the two accumulations in the usual order:
// the sums are bit-identical to NB = 1); see phase_residuals.
template <int C, int NB = 1>
__device__ __forceinline__ double phase_reduce(const LevelTemplate& L, const Work& W, float sigma, int loss, float good_thr,
                                               const TplCache& tc, const TemplateMeta& m, LinShared& sh, int block, int nblocks, bool write_partials = true) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float acc[30];
#pragma unroll
  for (int k = 0; k < 30; ++k) acc[k] = 0.0f;
  const float sigma_inv = __fdiv_rn(1.0f, sigma);
  const float is = 1.0f / m.s;
  const float w_invalid_good = (1.0f > good_thr) ? (float) C : 0.0f;   // invalid entries carry weight 1 in getWeights() (Q6)
  int ks = 0;
  int i_start = first_point(block, nblocks);
  if (NB == 2 && BP_PREFETCH && tc.K > 1 && tc.pts == kTcNone && tc.f[TC_R] == kTcNone) {
    const int stride = nblocks * kLinThreads;
    auto accumulate = [&](const float4& X, const VecC<C>& r, const VecC<C>& gx, const VecC<C>& gy) {
      float sxx = 0, sxy = 0, syy = 0, bx = 0, by = 0, e = 0, good = 0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float w = robust_weight(loss, r.v[c], sigma_inv);
        const float wr = w * r.v[c];
        const float wgx = w * gx.v[c], wgy = w * gy.v[c];
        sxx += wgx * gx.v[c]; sxy += wgx * gy.v[c]; syy += wgy * gy.v[c];
        bx += wr * gx.v[c]; by += wr * gy.v[c]; e += wr * r.v[c];
        good += (w > good_thr) ? 1.0f : 0.0f;
      }
      const float iz = 1.0f / X.z, iz2 = iz * iz;
      const float xc = X.x - m.c1, yc = X.y - m.c2, zc = (X.z - m.c3) * iz;
      float A[6], B[6];
      A[0] = -X.x * yc * iz2;        B[0] = -X.y * yc * iz2 - zc;
      A[1] = X.x * xc * iz2 + zc;    B[1] = X.y * xc * iz2;
      A[2] = -yc * iz;               B[2] = xc * iz;
      A[3] = iz * is;                B[3] = 0.0f;
      A[4] = 0.0f;                   B[4] = iz * is;
      A[5] = -X.x * iz2 * is;        B[5] = -X.y * iz2 * is;
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const float pa = sxx * A[a] + sxy * B[a], qa = sxy * A[a] + syy * B[a];
#pragma unroll
        for (int b = a; b < 6; ++b) acc[k++] += pa * A[b] + qa * B[b];
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] += bx * A[a] + by * B[a];
      acc[27] += e; acc[28] += good; acc[29] += 1.0f;
    };
    int i = i_start;
    for (; i + stride < m.n; i += 2 * stride, ks += 2) {
      const int ia = i, ib = i + stride, ja = i + 4 * stride, jb = i + 5 * stride;
      if (jb < m.n) {                         // two passes ahead into the L2
        prefetch_l2(L.gx + (size_t) ja * kStride<C>); prefetch_l2(L.gy + (size_t) ja * kStride<C>); prefetch_l2(W.res + (size_t) ja * kStride<C>); prefetch_l2(L.pts + ja);
        prefetch_l2(L.gx + (size_t) jb * kStride<C>); prefetch_l2(L.gy + (size_t) jb * kStride<C>); prefetch_l2(W.res + (size_t) jb * kStride<C>); prefetch_l2(L.pts + jb);
        if ((threadIdx.x & 31) == 0) { prefetch_l2(W.valid + ja); prefetch_l2(W.valid + jb); }
      }
      const bool va = W.valid[ia] != 0, vb = W.valid[ib] != 0;
      float4 Xa, Xb; VecC<C> ra, gxa, gya, rb, gxb, gyb;
      if (va) { Xa = __ldg(L.pts + ia); gxa.load(L.gx + (size_t) ia * kStride<C>); gya.load(L.gy + (size_t) ia * kStride<C>); ra.load_plain(W.res + (size_t) ia * kStride<C>); }
      if (vb) { Xb = __ldg(L.pts + ib); gxb.load(L.gx + (size_t) ib * kStride<C>); gyb.load(L.gy + (size_t) ib * kStride<C>); rb.load_plain(W.res + (size_t) ib * kStride<C>); }
      if (va) accumulate(Xa, ra, gxa, gya); else acc[28] += w_invalid_good;
      if (vb) accumulate(Xb, rb, gxb, gyb); else acc[28] += w_invalid_good;
    }
    i_start = i;
  }
  for (int i = i_start; i < m.n; i += nblocks * kLinThreads, ++ks) {
    if (BP_PREFETCH && tc.K > 1 && tc.pts == kTcNone && tc.f[TC_R] == kTcNone) {      // fully streaming level: two points ahead into the L2
      const int i2 = i + 2 * nblocks * kLinThreads;
      if (i2 < m.n) {
        if (tc.f[TC_GX] == kTcNone) prefetch_l2(L.gx + (size_t) i2 * kStride<C>);
        if (tc.f[TC_GY] == kTcNone) prefetch_l2(L.gy + (size_t) i2 * kStride<C>);
        if (tc.f[TC_R] == kTcNone) { prefetch_l2(W.res + (size_t) i2 * kStride<C>); if ((threadIdx.x & 31) == 0) prefetch_l2(W.valid + i2); }
        if (tc.pts == kTcNone) prefetch_l2(L.pts + i2);
      }
    }
    if (!((tc.f[TC_R] != kTcNone) ? tc_valid(tc, ks) : W.valid[i])) { acc[28] += w_invalid_good; continue; }
    const float4 X = (tc.pts != kTcNone) ? tc_point(tc, ks) : __ldg(L.pts + i);
    VecC<C> r, gx, gy;
    // (this order -- gx, gy, then r -- and per-field tests measured fastest in same-box A/B runs; a separate straight-line
    //  path for the all-cached case, or r first, cost 1-3 %: the kernel sits at the register limit and ptxas is touchy)
    if (tc.f[TC_GX] != kTcNone) tc_get<C>(tc, ks, TC_GX, gx); else gx.load(L.gx + (size_t) i * kStride<C>);
    if (tc.f[TC_GY] != kTcNone) tc_get<C>(tc, ks, TC_GY, gy); else gy.load(L.gy + (size_t) i * kStride<C>);
    if (tc.f[TC_R] != kTcNone) tc_get<C>(tc, ks, TC_R, r); else r.load_plain(W.res + (size_t) i * kStride<C>);
    float sxx = 0, sxy = 0, syy = 0, bx = 0, by = 0, e = 0, good = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float w = robust_weight(loss, r.v[c], sigma_inv);
      const float wr = w * r.v[c];
      const float wgx = w * gx.v[c], wgy = w * gy.v[c];
      sxx += wgx * gx.v[c]; sxy += wgx * gy.v[c]; syy += wgy * gy.v[c];
      bx += wr * gx.v[c]; by += wr * gy.v[c]; e += wr * r.v[c];
      good += (w > good_thr) ? 1.0f : 0.0f;
    }
    const float iz = 1.0f / X.z, iz2 = iz * iz;
    const float xc = X.x - m.c1, yc = X.y - m.c2, zc = (X.z - m.c3) * iz;
    float A[6], B[6];
    A[0] = -X.x * yc * iz2;        B[0] = -X.y * yc * iz2 - zc;
    A[1] = X.x * xc * iz2 + zc;    B[1] = X.y * xc * iz2;
    A[2] = -yc * iz;               B[2] = xc * iz;
    A[3] = iz * is;                B[3] = 0.0f;
    A[4] = 0.0f;                   B[4] = iz * is;
    A[5] = -X.x * iz2 * is;        B[5] = -X.y * iz2 * is;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const float pa = sxx * A[a] + sxy * B[a], qa = sxy * A[a] + syy * B[a];
#pragma unroll
      for (int b = a; b < 6; ++b) acc[k++] += pa * A[b] + qa * B[b];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += bx * A[a] + by * B[a];
    acc[27] += e; acc[28] += good; acc[29] += 1.0f;
  }
  // warp reduction of the 30 scalars as a transposing butterfly: 16+8+4+2+1 = 31 shuffles instead of 30 x 5;
  // afterwards lane l holds the warp total of scalar l.  Warps that own no point skip it altogether.
  const int n_active_warps = min(kLinThreads / 32, max(0, ((m.n + 31) / 32 - block + nblocks - 1) / nblocks));
  BP_FINE(20);
  __syncthreads();
  BP_FINE(21);
  if (warp < n_active_warps) {
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = (k < 30) ? acc[k] : 0.0f;
#pragma unroll
    for (int step = 0; step < 5; ++step) {
      const int off = 16 >> step;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int k = 0; k < (16 >> step); ++k) {
        const float send = upper ? v[k] : v[k + off];
        const float keep = upper ? v[k + off] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    sh.red[warp][lane] = (double) v[0];
  }
  BP_FINE(22);
  __syncthreads();
  BP_FINE(23);
  double v = 0.0;
  if (tid < 30) {
    for (int w = 0; w < n_active_warps; ++w) v += sh.red[w][tid];
    if (write_partials) { W.partials[(size_t) block * kPartialStride + tid] = v; __threadfence(); }   // last-CTA pattern of the host-driven path
  }
  BP_FINE(24);
  return v;
}

// the 30 fp64 totals in sh.red[0][0..29] -> LinOut (H symmetric column-major, G, sqrt(sum w r^2), counts)
__device__ __forceinline__ void finish_sums(float sigma, LinShared& sh, LinOut& out /* shared or global */) {
  const int tid = threadIdx.x;
  if (tid < 36) {                       // H: thread (a, b) picks its upper-triangle entry
    const int a = tid / 6, b = tid % 6, lo = a < b ? a : b, hi = a < b ? b : a;
    out.H[b * 6 + a] = (float) sh.red[0][lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
  } else if (tid < 42) {
    out.G[tid - 36] = (float) sh.red[0][21 + tid - 36];
  } else if (tid == 42) {
    out.f_norm = sqrtf((float) sh.red[0][27]);
    out.n_good = (int) (sh.red[0][28] + 0.5);
    out.n_valid = (int) (sh.red[0][29] + 0.5);
    out.sigma = sigma;
  }
  __syncthreads();
}

// fixed-order sum of the CTA partials -> LinOut (whole CTA participates; result valid in `out` for thread 0
// after the trailing __syncthreads, and broadcast through shared memory by the caller if needed)
__device__ __forceinline__ void final_sum(const Work& W, int nblocks, float sigma, LinShared& sh, LinOut& out /* shared or global */) {
  const int tid = threadIdx.x, k = tid & 31, g = tid >> 5;
  constexpr int G = kLinThreads / 32;
  double v = 0.0;
  for (int b0 = g; b0 < nblocks; b0 += 10 * G) {     // up to 10 independent L2 loads in flight per thread, fixed summation order
    double p[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) { const int b = b0 + j * G; p[j] = (b < nblocks) ? W.partials[(size_t) b * kPartialStride + k] : 0.0; }
#pragma unroll
    for (int j = 0; j < 10; ++j) v += p[j];
  }
  __syncthreads();
  sh.red[g][k] = v;
  __syncthreads();
  if (tid < 30) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < G; ++w) t += sh.red[w][tid];
    sh.red[0][tid] = t;
  }
  __syncthreads();
  finish_sums(sigma, sh, out);
}

// =============================================================================================
// host-driven path: four kernels per linearize (fine seam, bpvo_b200_linearize)
// =============================================================================================
struct LinArgs {
  LevelTemplate tmpl;
  LevelImage img;
  Work work;
  float P[12];
  int loss;
  int interp;
  float good_thr;
  unsigned* hset;     // histogram set (zeroed by the host before K1)
  Sel* sel;
};

template <int C, int BLEND> __global__ void __launch_bounds__(kLinThreads, 1) k_residuals(LinArgs a) {
  __shared__ LinShared sh;
  const bool do_hist = (a.loss != 0x12) && (a.work.scale->delta > 1e-6f);
  phase_residuals<C, BLEND>(a.tmpl, a.img, a.P, a.work, a.hset, do_hist, Bracket{false, 0.0f, 0.0f, 0.0f}, tpl_cache_off(), *a.tmpl.meta, nullptr, sh, blockIdx.x, gridDim.x, a.interp);
}
template <int C, int LEVEL> __global__ void __launch_bounds__(kLinThreads, 1) k_select(LinArgs a) {
  __shared__ LinShared sh;
  const bool do_hist = (a.loss != 0x12) && (a.work.scale->delta > 1e-6f);
  if (!do_hist) return;
  phase_select<C, LEVEL, 4>(a.tmpl, a.work, a.hset, a.sel, tpl_cache_off(), *a.tmpl.meta, sh, blockIdx.x, gridDim.x);
}
// point-sharded mode: the last CTA leaves this rank's 30 fp64 sums in a.sums (all-reduced by the host over NCCL),
// k_finalize_sums then builds the LinOut every rank sees identically.
template <int C> __global__ void __launch_bounds__(kLinThreads, 1) k_reduce_sharded(LinArgs a, double* __restrict__ sums) {
  __shared__ LinShared sh;
  __shared__ bool s_last;
  const bool do_hist = (a.loss != 0x12) && (a.work.scale->delta > 1e-6f);
  float sigma = a.work.scale->scale;
  if (do_hist) sigma = finish_scale<C>(a.work, a.hset, a.sel, sh);
  phase_reduce<C>(a.tmpl, a.work, sigma, a.loss, a.good_thr, tpl_cache_off(), *a.tmpl.meta, sh, blockIdx.x, gridDim.x);
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.work.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    LinOut dummy;
    final_sum(a.work, gridDim.x, sigma, sh, dummy);       // leaves the fp64 totals in sh.red[0][0..29]
    if (threadIdx.x < 32) sums[threadIdx.x] = (threadIdx.x < 30) ? sh.red[0][threadIdx.x] : 0.0;
    if (threadIdx.x == 0) {
      a.work.out->sigma = sigma;
      if (do_hist) { a.work.scale->delta = fabsf(sigma - a.work.scale->scale); a.work.scale->scale = sigma; }
      *a.work.ticket = 0;
    }
  }
}

__global__ void k_finalize_sums(const double* __restrict__ sums, LinOut* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int q = 0;
  for (int a = 0; a < 6; ++a)
    for (int b = a; b < 6; ++b) { const float h = (float) sums[q++]; out->H[b * 6 + a] = h; out->H[a * 6 + b] = h; }
  for (int a = 0; a < 6; ++a) out->G[a] = (float) sums[21 + a];
  out->f_norm = sqrtf((float) sums[27]);
  out->n_good = (int) (sums[28] + 0.5);
  out->n_valid = (int) (sums[29] + 0.5);
}

template <int C> __global__ void __launch_bounds__(kLinThreads, 1) k_reduce(LinArgs a) {
  __shared__ LinShared sh;
  __shared__ bool s_last;
  const bool do_hist = (a.loss != 0x12) && (a.work.scale->delta > 1e-6f);
  float sigma = a.work.scale->scale;
  if (do_hist) sigma = finish_scale<C>(a.work, a.hset, a.sel, sh);
  phase_reduce<C>(a.tmpl, a.work, sigma, a.loss, a.good_thr, tpl_cache_off(), *a.tmpl.meta, sh, blockIdx.x, gridDim.x);
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(a.work.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    final_sum(a.work, gridDim.x, sigma, sh, *a.work.out);
    if (threadIdx.x == 0) {
      if (do_hist) { a.work.scale->delta = fabsf(sigma - a.work.scale->scale); a.work.scale->scale = sigma; }
      *a.work.ticket = 0;
    }
  }
}

__global__ void k_reset_scale(ScaleState* s) { s->scale = 1.0f; s->delta = 1e10f; }

// weights of the last linearize in the reference's channel-major layout (getWeights(), Q6: invalid -> f(0))
template <int C>
__global__ void __launch_bounds__(256) k_export_weights(const float* __restrict__ res, int n, float sigma, int loss,
                                                        float* __restrict__ w_out, float* __restrict__ r_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const int i = t / C, c = t - i * C;
  const float r = res[(size_t) i * kStride<C> + c];
  if (w_out) w_out[(size_t) c * n + i] = robust_weight(loss, r, __fdiv_rn(1.0f, sigma));
  if (r_out) r_out[(size_t) c * n + i] = r;
}

// getPointCloudFromRefFrame (vo.cc:249-281) on the device: xyzw, grey level of the ref image at K_l X, channel-0 weight
struct PointInfo { float x, y, z, w; unsigned rgba; float weight; unsigned pad[2]; };      // bpvo::PointWithInfo's 32-byte layout
template <int C>
__global__ void __launch_bounds__(256) k_point_cloud(const float4* __restrict__ pts, int n, const uint8_t* __restrict__ image, int rows, int cols, int pitch,
                                                     float fx, float fy, float cx, float cy, const float* __restrict__ res, float sigma, int loss,
                                                     PointInfo* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 X = __ldg(pts + i);
  // warp.getImagePoint: x = K * X.head<3>() (rigid_body_warp.h:123-128), dense 3x3 product with the zero terms kept
  float x0 = __fmul_rn(fx, X.x); x0 = __fadd_rn(x0, __fmul_rn(0.0f, X.y)); x0 = __fadd_rn(x0, __fmul_rn(cx, X.z));
  float x1 = __fmul_rn(0.0f, X.x); x1 = __fadd_rn(x1, __fmul_rn(fy, X.y)); x1 = __fadd_rn(x1, __fmul_rn(cy, X.z));
  float x2 = __fmul_rn(0.0f, X.x); x2 = __fadd_rn(x2, __fmul_rn(0.0f, X.y)); x2 = __fadd_rn(x2, __fmul_rn(1.0f, X.z));
  const float z_i = __fdiv_rn(1.0f, x2), u = __fmul_rn(z_i, x0), v = __fmul_rn(z_i, x1);
  const unsigned c = (v >= 0 && v < rows && u >= 0 && u < cols) ? (unsigned) __ldg(image + (size_t) ((int) v) * pitch + (int) u) : 0u;
  PointInfo o;
  o.x = X.x; o.y = X.y; o.z = X.z; o.w = X.w;
  o.rgba = c | (c << 8) | (c << 16) | (255u << 24);
  o.weight = robust_weight(loss, res[(size_t) i * kStride<C>], __fdiv_rn(1.0f, sigma));
  o.pad[0] = o.pad[1] = 0u;
  out[i] = o;
}

// =============================================================================================
// persistent path: the whole estimatePose (all levels, all GN iterations, 6x6 solves, convergence
// tests, pose updates) in ONE cooperative launch; the host sees only T_est and the statistics.
// Every CTA carries the (tiny) solver state redundantly in shared memory and computes identical
// values.  Per GN iteration the CTAs meet twice:
//   * after P1: a grid barrier (counter + release/acquire), then every CTA finds the exact median itself;
//   * after P4: NO barrier -- the 30 fp64 sums travel in a two-level flag-in-data exchange (each 16-byte word
//     carries its own sequence number, so one L2 round trip delivers data and "ready" together), summed in a fixed
//     order: bit-identical in every CTA and from run to run.
// =============================================================================================
// optional in-kernel phase profile (CTA 0, thread 0): cycles accumulated per phase of the GN loop
#define BP_PROF(slot)                                                                       \
  do { if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) { long long* _s = prof_smem(); const long long _t = clock64(); _s[slot] += _t - _s[64]; _s[64] = _t; } } while (0)
enum { PROF_P1 = 0, PROF_SYNC1, PROF_P2, PROF_SYNC2, PROF_P3, PROF_SYNC3, PROF_SCALE, PROF_P4, PROF_SYNC4, PROF_FINAL, PROF_SOLVE, PROF_OTHER, PROF_COUNT };

#ifndef BP_BRACKET_TARGET
#define BP_BRACKET_TARGET 500     /* candidates the bracket is sized for once the median has settled */
#endif
// Every wait of the persistent kernel (grid barrier, flag-in-data polls, fixed-point exchange, cross-rank mailboxes) is bounded by
// ONE absolute deadline on the GPU's global timer, set by the host per launch (2 s on a single GPU: a solve takes milliseconds; 60 s
// in the multi-rank mode, where a peer PROCESS may be seconds behind): a lost CTA / rank ends the launch with an error status
// instead of hanging the GPU.  A wait that expires raises the CTA's flag (every later wait of the CTA returns at once) and the
// rank-wide word behind the barrier counter, which CTA 0 reports to the host together with the statistics.
struct AbortCtl {
  int flag;                      // this CTA gave up
  int pad;
  unsigned long long deadline;   // %globaltimer value (ns) after which waits give up
  int* global_flag;              // rank-wide abort word
};
__device__ __forceinline__ bool wait_expired(unsigned& spins, AbortCtl* ac) {
  if (*(volatile int*) &ac->flag) return true;
  if ((++spins & 511u) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t > ac->deadline || *(volatile int*) ac->global_flag) { ac->flag = 1; *(volatile int*) ac->global_flag = 1; return true; }
  }
  return false;
}

// Grid-wide barrier of the persistent kernel (all CTAs are co-resident: cooperative launch).  `counter` only grows;
// `epoch` is the value it reaches when every CTA has arrived at this barrier (kept uniformly by all threads).
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned nblocks, AbortCtl* abort_flag) {
  epoch += nblocks;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned v, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int) (v - epoch) >= 0) break;
      if (wait_expired(spins, abort_flag)) break;
    }
  }
  __syncthreads();
}

// flag-in-data words: {lo32(value), seq, hi32(value), seq}; a reader accepts a word only when both halves carry the
// sequence number it waits for (needs only 8-byte single-copy atomicity)
__device__ __forceinline__ void ll_post(uint4* p, double v, unsigned seq) {
  const unsigned long long b = (unsigned long long) __double_as_longlong(v);
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned) b), "r"(seq), "r"((unsigned) (b >> 32)), "r"(seq) : "memory");
}
__device__ __forceinline__ bool ll_peek(const uint4* p, unsigned seq, double& v) {
  unsigned x, f0, y, f1;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(f0), "=r"(y), "=r"(f1) : "l"(p) : "memory");
  v = __longlong_as_double((long long) (((unsigned long long) y << 32) | x));
  return f0 == seq && f1 == seq;
}

// sum of up to NS words p[q * stride], q < cnt, all polled concurrently; fixed summation order
template <int NS>
__device__ __forceinline__ double ll_gather(const uint4* p, size_t stride, int cnt, unsigned seq, AbortCtl* abort_flag) {
  double v[NS]; bool ok[NS]; bool all = true;
#pragma unroll
  for (int q = 0; q < NS; ++q) { v[q] = 0.0; ok[q] = q >= cnt; all = all && ok[q]; }
  unsigned spins = 0;
  while (!all) {
    all = true;
#pragma unroll
    for (int q = 0; q < NS; ++q) { if (!ok[q]) ok[q] = ll_peek(p + q * stride, seq, v[q]); all = all && ok[q]; }
    if (!all && wait_expired(spins, abort_flag)) break;
  }
  double t = 0.0;
#pragma unroll
  for (int q = 0; q < NS; ++q) t += ok[q] && q < cnt ? v[q] : 0.0;
  return t;
}

// All-to-all sum of the CTA totals (thread k < 30 passes `mine` = CTA total of scalar k) without a barrier:
//   stage 1: CTA b posts its totals to slot b; the first CTA of every group of kLLGroup polls its members and posts the
//            group totals;  stage 2: every CTA polls the <= ceil(nblocks / kLLGroup) group totals.
// Mailboxes are double-buffered by the parity of `seq`: a CTA can be at most one exchange ahead of any other.
// Result: sh.red[0][0..29] (valid after the trailing __syncthreads).
__device__ __forceinline__ void exchange_sums(double mine, uint4* ll, unsigned seq, LinShared& sh, int blk, int nb, AbortCtl* abort_flag) {
  const int tid = threadIdx.x, k = tid & 31, g = tid >> 5;
  constexpr int G = kLinThreads / 32;
  uint4* box1 = ll + (size_t) (seq & 1u) * kMaxGrid * 32;                       // CTA totals
  uint4* box2 = ll + (size_t) (2 + (seq & 1u)) * kMaxGrid * 32;                 // group totals
  const int grp = blk / kLLGroup, lead = grp * kLLGroup, ngroups = (nb + kLLGroup - 1) / kLLGroup;
  if (blk != lead) {
    if (tid < 30) ll_post(box1 + (size_t) blk * 32 + tid, mine, seq);
  } else {
    BP_FINE(40);
    const int members = min(kLLGroup, nb - lead) - 1;                          // besides the leader itself
    constexpr int NS = (kLLGroup - 1 + G - 1) / G;
    double v = 0.0;
    if (k < 30) {
      const int left = members - g;                                            // members g, g + G, ... of this thread
      v = ll_gather<NS>(box1 + (size_t) (lead + 1 + g) * 32 + k, (size_t) G * 32, left > 0 ? (left + G - 1) / G : 0, seq, abort_flag);
    }
    sh.xch[g][k] = v;
    __syncthreads();
    if (tid < 30) {
      double t = mine;
#pragma unroll
      for (int w = 0; w < G; ++w) t += sh.xch[w][tid];
      ll_post(box2 + (size_t) grp * 32 + tid, t, seq);
    }
    __syncthreads();
    BP_FINE(41);
  }
  double v = 0.0;
  if (k < 30) {
    for (int q0 = g; q0 < ngroups; q0 += 2 * G) {
      const int left = ngroups - q0;
      v += ll_gather<2>(box2 + (size_t) q0 * 32 + k, (size_t) G * 32, (left + G - 1) / G, seq, abort_flag);
    }
  }
  BP_FINE(42);
  sh.xch[g][k] = v;
  __syncthreads();
  if (tid < 30) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < G; ++w) t += sh.xch[w][tid];
    sh.red[0][tid] = t;          // column tid of red[] is only ever touched by thread tid outside barriers
  }
  __syncthreads();
  BP_FINE(43);
}

// =============================================================================================
// Cross-rank exchanges of the point-sharded mode, INSIDE the persistent kernel (peer-memory mode, comm.cu):
// every rank's mailbox lives in its own memory and is written by the peers with plain stores over NVLink (CUDA IPC
// mappings); 8-byte words {value, sequence number} validate themselves, so a message needs no fence and no separate
// flag: the receiver polls its LOCAL memory.  Slots are double-buffered by sequence parity (a rank can be at most one
// exchange ahead of any other: it needs everybody's previous message to get there).  Only CTA 0 of a rank talks to
// the peers; it hands rank-wide results to the other CTAs through a local grid barrier or local flag-in-data words.
// All ranks take identical decisions from identical data, so the exchange sequence is the same everywhere.
// =============================================================================================
__device__ __forceinline__ void x_put(uint2* p, unsigned v, unsigned seq) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(seq) : "memory");
}
__device__ __forceinline__ bool x_peek(const uint2* p, unsigned seq, unsigned& v) {
  unsigned f;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
  return f == seq;
}
__device__ __forceinline__ uint2* x_slot(const PeerArgs& pa, int owner, unsigned seq, int src) {
  return pa.box[owner] + ((size_t) (seq & 1u) * kXRanks + src) * kXWords;
}
__device__ __forceinline__ unsigned x_wait(const uint2* p, unsigned seq, AbortCtl* abort_flag) {
  unsigned v, spins = 0;
  while (!x_peek(p, seq, v)) { if (wait_expired(spins, abort_flag)) { v = 0; break; } }
  return v;
}
// CTA 0: send buf[0..n) (shared memory) to every peer
__device__ __forceinline__ void x_send(const PeerArgs& pa, unsigned seq, const unsigned* buf, int n) {
  for (int p = 0; p < pa.nranks; ++p) {
    if (p == pa.rank) continue;
    uint2* dst = x_slot(pa, p, seq, pa.rank);
    for (int i = threadIdx.x; i < n; i += kLinThreads) x_put(dst + i, buf[i], seq);
  }
}
// CTA 0: sum (u32) of buf[0..n) over all ranks, in place (shared memory); thread t owns words t, t + 256, ...
// All peers are polled concurrently (up to kXRanks x 4 words in flight per thread): one round trip, whatever the rank count.
__device__ __forceinline__ void x_allreduce_u32(const PeerArgs& pa, unsigned& xseq, unsigned* buf, int n, AbortCtl* abort_flag) {
  const unsigned seq = xseq++;
  x_send(pa, seq, buf, n);
  const uint2* base = x_slot(pa, pa.rank, seq, 0);
  constexpr int Q = 5;                                      // words per thread and chunk: the 1027-word bracket message is ONE chunk
  for (int i0 = threadIdx.x; i0 < n; i0 += Q * kLinThreads) {
    unsigned acc[Q];
    unsigned long long pending = 0;                         // bit r * 8 + q: word q of rank r still missing
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0u;
    for (int r = 0; r < pa.nranks; ++r)
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (r != pa.rank && i0 + q * kLinThreads < n) pending |= 1ull << (r * 8 + q);
    unsigned spins = 0;
    while (pending) {
      for (int r = 0; r < pa.nranks; ++r) {
        if (!((pending >> (r * 8)) & 0xffull)) continue;
#pragma unroll
        for (int q = 0; q < Q; ++q)
          if (pending & (1ull << (r * 8 + q))) {
            unsigned v;
            if (x_peek(base + (size_t) r * kXWords + i0 + q * kLinThreads, seq, v)) { acc[q] += v; pending &= ~(1ull << (r * 8 + q)); }
          }
      }
      if (pending && wait_expired(spins, abort_flag)) break;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) if (i0 + q * kLinThreads < n) buf[i0 + q * kLinThreads] += acc[q];
  }
  __syncthreads();
}
// CTA 0, thread k < 30: sum of `mine` over all ranks in RANK ORDER (bit-identical on every rank)
__device__ __forceinline__ double x_allreduce_f64_ordered(const PeerArgs& pa, unsigned& xseq, double mine, AbortCtl* abort_flag) {
  const unsigned seq = xseq++;
  const int k = threadIdx.x;
  double t = 0.0;
  if (k < 30) {
    const unsigned long long b = (unsigned long long) __double_as_longlong(mine);
    for (int p = 0; p < pa.nranks; ++p) {
      if (p == pa.rank) continue;
      uint2* dst = x_slot(pa, p, seq, pa.rank);
      x_put(dst + 2 * k, (unsigned) b, seq); x_put(dst + 2 * k + 1, (unsigned) (b >> 32), seq);
    }
    unsigned lo[kXRanks], hi[kXRanks]; bool ok[kXRanks]; bool all = true;
#pragma unroll
    for (int r = 0; r < kXRanks; ++r) { ok[r] = (r >= pa.nranks) || (r == pa.rank); lo[r] = hi[r] = 0; all = all && ok[r]; }
    unsigned spins = 0;
    while (!all) {
      all = true;
#pragma unroll
      for (int r = 0; r < kXRanks; ++r) {
        if (!ok[r]) { const uint2* src = x_slot(pa, pa.rank, seq, r); unsigned a, c; const bool g0 = x_peek(src + 2 * k, seq, a), g1 = x_peek(src + 2 * k + 1, seq, c); if (g0 && g1) { lo[r] = a; hi[r] = c; ok[r] = true; } }
        all = all && ok[r];
      }
      if (!all && wait_expired(spins, abort_flag)) break;
    }
#pragma unroll
    for (int r = 0; r < kXRanks; ++r)
      if (r < pa.nranks) t += (r == pa.rank) ? mine : __longlong_as_double((long long) (((unsigned long long) hi[r] << 32) | lo[r]));
  }
  return t;
}

// CTA 0 of every rank: the bracketed exact median over ALL ranks' residuals.  Same algorithm as bracket_select(), with
// two exchanges: the bracket histogram + counters are all-reduced, then the (few) candidates of the two wanted bins are
// all-gathered.  Every rank computes the same verdict and the same pair of order statistics.
constexpr unsigned kXPoison = 1u << 26;      // per-rank stand-in for kCandPoison that cannot overflow a u32 sum over 8 ranks
template <int C>
__device__ __forceinline__ bool bracket_select_xrank(const PeerArgs& pa, unsigned& xseq, const Work& W, const unsigned* __restrict__ hset, LinShared& sh,
                                                     unsigned* scratch, int nblocks, const Bracket& br, AbortCtl* abort_flag,
                                                     unsigned& n_out, unsigned& ncand_out, float& lo_out, float& hi_out) {
  const int tid = threadIdx.x;
  const unsigned total = (unsigned) nblocks * kCandPerCta;
  const uint4 gb = __ldcg(reinterpret_cast<const uint4*>(hset + kHistBins + 8) + tid);
  float pre[kSelPre];
#pragma unroll
  for (int q = 0; q < kSelPre; ++q) { const unsigned j = tid + q * kLinThreads; pre[q] = (j < total) ? __ldcg(W.cand + j) : -1.0f; }
  const unsigned nv_l = __ldcg(hset + kHistBins + 1), below_l = __ldcg(hset + kHistBins + 2), ncand_l = __ldcg(hset + kHistBins + 3);
  const unsigned novf = __ldcg(hset + kHistBins + 4);
  const float* ovf = W.cand + (size_t) nblocks * kCandPerCta;
  unsigned* buf = sh.hist;
  buf[4 * tid + 0] = gb.x; buf[4 * tid + 1] = gb.y; buf[4 * tid + 2] = gb.z; buf[4 * tid + 3] = gb.w;
  if (tid == 0) { buf[kSelBins] = nv_l; buf[kSelBins + 1] = below_l; buf[kSelBins + 2] = (ncand_l >= kCandPoison || novf > 16384u) ? kXPoison : ncand_l; }
  __syncthreads();
  x_allreduce_u32(pa, xseq, buf, kSelBins + 3, abort_flag);
  const unsigned n = buf[kSelBins] * (unsigned) C, below = buf[kSelBins + 1], ncand = buf[kSelBins + 2];
  n_out = n; ncand_out = ncand;
  if (n < 3 || ncand >= kXPoison) return false;
  const unsigned t_hi = n / 2, t_lo = (n % 2 == 0) ? t_hi - 1 : t_hi;
  if (below > t_lo || t_hi >= below + ncand) return false;
  const unsigned ra = t_lo - below, rb = t_hi - below;
  float* list = reinterpret_cast<float*>(scratch + kCtaCandCap);   // [kSelList]
  const unsigned loc[4] = {buf[4 * tid + 0], buf[4 * tid + 1], buf[4 * tid + 2], buf[4 * tid + 3]};
  unsigned bin_a, rem_a, bin_b, rem_b, tot;
  block_find2_regs<4>(loc, ra, rb, sh, bin_a, rem_a, bin_b, rem_b, tot);      // resets sh.found[0..7]
  // this rank's candidates in the two wanted bins -> list[0 .. nl)
#pragma unroll
  for (int q = 0; q < kSelPre; ++q) {
    const unsigned b = (unsigned) sel_bin(pre[q], br.lo, br.inv_w);
    if (pre[q] >= 0.0f && (b == bin_a || b == bin_b)) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = pre[q]; }
  }
  for (unsigned j = tid + kSelPre * kLinThreads; j < total; j += kLinThreads) {
    const float v = __ldcg(W.cand + j);
    const unsigned b = (unsigned) sel_bin(v, br.lo, br.inv_w);
    if (v >= 0.0f && (b == bin_a || b == bin_b)) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; }
  }
  for (unsigned j = tid; j < novf; j += kLinThreads) {
    const float v = __ldcg(ovf + j);
    const unsigned b = (unsigned) sel_bin(v, br.lo, br.inv_w);
    if (b == bin_a || b == bin_b) { const unsigned slot = atomicAdd(&sh.found[4], 1u); if (slot < (unsigned) kSelList) list[slot] = v; }
  }
  __syncthreads();
  // all-gather of the short lists: word 0 = count (0xffffffff: too many here -> everybody falls back), then the values.
  // Thread (r = tid / 32, j = tid % 32) fetches value j of rank r: all peers are polled concurrently.
  const unsigned nl_own = sh.found[4];
  const bool own_bad = nl_own > 31u;
  const unsigned seq = xseq++;
  if (tid == 0) buf[0] = own_bad ? 0xffffffffu : nl_own;
  if (!own_bad && tid < (int) nl_own) buf[1 + tid] = __float_as_uint(list[tid]);
  __syncthreads();
  x_send(pa, seq, buf, own_bad ? 1 : 1 + (int) nl_own);
  static_assert(kLinThreads / 32 == kXRanks, "one warp per source rank in the list gather");
  unsigned* cnt = sh.scan;                                   // [kXRanks]: per-rank counts
  float* mine = reinterpret_cast<float*>(buf + 64);         // own values, saved before the merged list overwrites list[]
  if (!own_bad && tid < (int) nl_own) mine[tid] = list[tid];
  if (tid < kXRanks) cnt[tid] = (tid >= pa.nranks) ? 0u : (tid == pa.rank) ? (own_bad ? 0xffffffffu : nl_own) : x_wait(x_slot(pa, pa.rank, seq, tid), seq, abort_flag);
  __syncthreads();
  bool bad = false; unsigned base = 0, my_off = 0;
  {
    const int r_of_t = tid >> 5;
#pragma unroll
    for (int r = 0; r < kXRanks; ++r) { const unsigned c = cnt[r]; if (c == 0xffffffffu) bad = true; else { if (r < r_of_t) my_off += c; base += c; } }
  }
  __syncthreads();
  if (!bad && base <= (unsigned) kSelList) {
    const int r = tid >> 5, j = tid & 31;
    if (r < pa.nranks && j < (int) cnt[r])
      list[my_off + j] = (r == pa.rank) ? mine[j] : __uint_as_float(x_wait(x_slot(pa, pa.rank, seq, r) + 1 + j, seq, abort_flag));
  }
  __syncthreads();
  if (bad || base > (unsigned) kSelList || *(volatile int*) &abort_flag->flag) return false;
  if (tid < (int) base) {
    const float v = list[tid];
    const unsigned bj = (unsigned) sel_bin(v, br.lo, br.inv_w);
    unsigned rank = 0;
    for (unsigned j = 0; j < base; ++j) {
      const float u = list[j];
      const unsigned bu = (unsigned) sel_bin(u, br.lo, br.inv_w);
      rank += (bu == bj && (u < v || (u == v && j < (unsigned) tid))) ? 1u : 0u;
    }
    if (bj == bin_a && rank == rem_a) sh.found[6] = __float_as_uint(v);
    if (bj == bin_b && rank == rem_b) sh.found[7] = __float_as_uint(v);
  }
  __syncthreads();
  lo_out = __uint_as_float(sh.found[6]); hi_out = __uint_as_float(sh.found[7]);
  return true;
}

// all CTAs: a histogram range of this rank's set -> its sum over all ranks, in place; ends with a grid barrier
__device__ __forceinline__ void xrank_hist_allreduce(const PeerArgs& pa, unsigned& xseq, unsigned* hist, int n, LinShared& sh,
                                                     unsigned* counter, unsigned& epoch, int nblocks, AbortCtl* abort_flag) {
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < n; i += kLinThreads) sh.hist[i] = __ldcg(hist + i);
    __syncthreads();
    x_allreduce_u32(pa, xseq, sh.hist, n, abort_flag);
    for (int i = threadIdx.x; i < n; i += kLinThreads) hist[i] = sh.hist[i];
  }                                  // (only CTA 0 ever uses xseq)
  grid_barrier(counter, epoch, nblocks, abort_flag);
}

// what follows a rejected fast LDL^T (rare): the same factorisation with Eigen's early-outs, Eigen's pivoted LDLT, the damped
// fp64 retry of solve6() (pose_estimator_base.h:90-148) -- out of line so that it neither bloats the hot loop nor forces the
// caller's register copy of H into local memory.  Every lane of the calling warp stores the same dp.
__device__ __noinline__ bool solve6_cold(const float* H, const float* G, float* dp_out) {
  float dp[6];
  bool ok = solve6_fp32_registers<false>(H, G, dp);
  if (!ok) ok = solve6_fp32_registers<true>(H, G, dp);
  if (!ok) ok = solve6(H, G, dp);
  for (int k = 0; k < 6; ++k) dp_out[k] = dp[k];
  return ok;
}

struct SolveShared {
  AbortCtl abort;      // a barrier / exchange timed out: every later wait returns immediately, the launch reports an error
  LinOut lin;
  M44 T, Td;
  float P[12];
  float dp[6];
  float scale, delta;
  float dpn, gn;       // |dp| and max|G| of the step just solved (warp0_finish)
  int fx_bad;          // the fixed-point exchange overflowed somewhere: repeat this exchange with the fp64 mailboxes
  // bracketed-median state of the current level (identical in every CTA)
  int br_on;           // a previous median exists
  float br_lo, br_hi;  // previous middle order statistics
  float br_rel;        // relative half-width of the next bracket
  float br_density;    // candidates per unit of relative half-width, from the last bracketed pass
};

// grid-uniform bookkeeping of the persistent kernel (identical in every thread of every CTA)
struct GridSync {
  unsigned* counter;   // grid-barrier counter (zeroed by the host before the launch)
  unsigned epoch;      // value of *counter once every CTA has passed the last barrier
  unsigned seq;        // sequence number of the next exchange
  unsigned xseq;       // sequence number of the next cross-rank exchange (peer-memory mode)
  unsigned fxn;        // fixed-point exchanges done so far (their accumulator sets alternate)
  int hs;              // histogram set of the next linearize (cycles through kHistSets)
};

// The two-points-in-flight reduce phase of the streaming levels as an out-of-line function: its register allocation (30
// accumulators + two sets of gradients / residuals in flight) stays out of the persistent kernel's, which sits at the 255-register
// limit, and the call costs one save / restore per phase, not per point.  Every structure goes in BY VALUE: a reference would force
// the caller's copy into local memory for the whole kernel.  Even so a kernel containing the call runs its CACHED levels a third
// slower (the long-lived state is spilled around the call site): the instantiation that carries it (FIX bit 8) is only ever
// launched for the streaming level itself -- the host splits the solve by level (SolveArgs::chain).
template <int C>
__device__ __noinline__ double phase_reduce_stream2(const LevelTemplate L, const Work W, float sigma, int loss, float good_thr, const TplCache tc,
                                                    const TemplateMeta m, LinShared& sh, int block, int nblocks) {
  return phase_reduce<C, 2>(L, W, sigma, loss, good_thr, tc, m, sh, block, nblocks, false);
}

// Front part of one linearize inside the persistent kernel: residuals -> exact median / scale -> weights and normal equations
// of this CTA's points.  Returns, in thread k < 30, the CTA total of scalar k (layout of phase_reduce).
// FIX: 0 = loss function and interpolant are run-time parameters; 0x10 / 0x11 / 0x12 = that RobustFunction with linear interpolation
// compiled in (the hot configurations of the bit-planes workloads: the other branches fall away, the kernel is 11 % smaller and
// 3 % faster in the same-box A/B)
template <int C, int BLEND, bool PEER, int FIX>
__device__ __forceinline__ double device_linearize(const SolveArgs& a, int lvl, SolveShared& ss, LinShared& sh,
                                                   const TplCache& tc, const TemplateMeta& meta, unsigned* scratch, GridSync& gs, Sel* sel,
                                                   float& sigma_out, bool& do_hist_out) {
  const LevelTemplate& L = a.tmpl[lvl];
  const LevelImage& I = a.img[lvl];
  const int nb = gridDim.x, blk = blockIdx.x, tid = threadIdx.x;
  unsigned* hset = a.work.hist + (size_t) gs.hs * kHistWords;
  // the set of the PREVIOUS linearize: zeroed after this iteration's barrier, next used two iterations from now, i.e.
  // with the next iteration's barrier in between (all sets are zero at the start of a level)
  unsigned* hprev = a.work.hist + (size_t) ((gs.hs + kHistSets - 1) % kHistSets) * kHistWords;
  const int loss = FIX ? (FIX & 0xff) : a.sp.loss, interp = FIX ? 0 : a.sp.interp;      // FIX bit 8: two points in flight on streaming levels (phase_residuals NB)
  const bool do_hist = (loss != 0x12) && (ss.delta > 1e-6f);      // ss.P was set by thread 0 together with the pose
  if (tid == 0) ss.lin.pad[1] = do_hist ? 1 : 0;                       // diagnosis: 0 = scale kept, 1 = radix select, 3 = bracketed select
  BP_PROF(PROF_OTHER);
  Bracket br;
  br.on = do_hist && ss.br_on;
  br.lo = ss.br_lo * (1.0f - ss.br_rel); br.hi = ss.br_hi * (1.0f + ss.br_rel);
  br.inv_w = (br.hi > br.lo) ? (float) kSelBins / (br.hi - br.lo) : 0.0f;
  const bool multi = PEER && a.peer.nranks > 1 && !meta.replicated;      // (PEER = false: the single-GPU kernel carries no cross-rank code)
  const bool use_msg = BP_MSG_SELECT && br.on && !multi && nb <= 148;    // the median facts travel as flag-in-data messages: no grid barrier
  constexpr int SEL_MLP = 1;      // loads in flight per thread in the select passes (4 was measured SLOWER inside this kernel, even in the streaming instantiation: 14.6 -> 21.4 us; the host-driven k_select gains 3 % from it)
  const bool stream2 = (FIX & 0x100) && BP_PREFETCH && tc.K > 1 && tc.pts == kTcNone && tc.f[TC_R] == kTcNone;      // level-uniform
  phase_residuals<C, BLEND, (FIX & 0x100) ? 2 : 1>(L, I, ss.P, a.work, hset, do_hist, br, tc, meta, scratch, sh, blk, nb, interp, use_msg ? gs.seq : 0u);
  BP_PROF(PROF_P1);
  float sigma = ss.scale;
  if (do_hist) {
    unsigned n = 0, ncand = 0; float lo = 0.0f, hi = 0.0f;
    bool hit = false;
    int verdict = 0;
    if (use_msg) verdict = message_select<C>(a.work, gs.seq, sh, scratch, nb, br, &ss.abort, n, ncand, lo, hi);
    if (verdict == 1) hit = true;
    if (verdict == 0) {                 // first evaluation of a level, multi-rank level, or a bracket too wide for the messages: barrier + global lists
      grid_barrier(gs.counter, gs.epoch, nb, &ss.abort);
      BP_PROF(PROF_SYNC1);
      BP_FINE(32);
      if (br.on && !multi) hit = bracket_select<C>(a.work, hset, sh, scratch, blk, nb, br, gs.counter, gs.epoch, &ss.abort, n, ncand, lo, hi);
      if (br.on && multi) {
        // CTA 0 talks to the peers (two exchanges) and hands the verdict to the other CTAs through three local words
        uint4* lb = a.peer.lbox + (size_t) (gs.seq & 1u) * 64;
        if (blk == 0) {
          hit = bracket_select_xrank<C>(a.peer, gs.xseq, a.work, hset, sh, scratch, nb, br, &ss.abort, n, ncand, lo, hi);
          if (tid == 0) {
            ll_post(lb + 0, __longlong_as_double((long long) (((unsigned long long) n << 32) | (hit ? 1u : 0u))), gs.seq);
            ll_post(lb + 1, __longlong_as_double((long long) (((unsigned long long) __float_as_uint(hi) << 32) | __float_as_uint(lo))), gs.seq);
            ll_post(lb + 2, __longlong_as_double((long long) (unsigned long long) ncand), gs.seq);
          }
        } else {
          if (tid < 3) sh.xch[0][tid] = ll_gather<1>(lb + tid, 0, 1, gs.seq, &ss.abort);
          __syncthreads();
          const unsigned long long w0 = (unsigned long long) __double_as_longlong(sh.xch[0][0]), w1 = (unsigned long long) __double_as_longlong(sh.xch[0][1]),
                                   w2 = (unsigned long long) __double_as_longlong(sh.xch[0][2]);
          hit = (w0 & 1u) != 0; n = (unsigned) (w0 >> 32); lo = __uint_as_float((unsigned) w1); hi = __uint_as_float((unsigned) (w1 >> 32)); ncand = (unsigned) w2;
          __syncthreads();
        }
      }
    }
    // the set of the linearize before this one: every CTA has long finished with it (two exchanges ago); it is next used three
    // GN iterations from now
    for (int b = blk * kLinThreads + tid; b < kHistWords; b += nb * kLinThreads) hprev[b] = 0;
    if (hit) {
      const float med = (n % 2 != 0) ? hi : (float) ((double) __fadd_rn(lo, hi) / 2.0);
      sigma = scale_from_median(n, med);
      BP_PROF(PROF_SCALE);
    } else {
      if (br.on) { phase_hist1<C, SEL_MLP>(a.work, hset, tc, meta, sh, blk, nb); grid_barrier(gs.counter, gs.epoch, nb, &ss.abort); }   // bracket missed: build the histogram now
      if (multi) xrank_hist_allreduce(a.peer, gs.xseq, hset, kHist1Bins, sh, gs.counter, gs.epoch, nb, &ss.abort);
      phase_select<C, 2, SEL_MLP>(L, a.work, hset, sel, tc, meta, sh, blk, nb);
      BP_PROF(PROF_P2);
      grid_barrier(gs.counter, gs.epoch, nb, &ss.abort);
      if (multi) xrank_hist_allreduce(a.peer, gs.xseq, hset + kHist1Bins, 2 * kHist2Bins, sh, gs.counter, gs.epoch, nb, &ss.abort);
      BP_PROF(PROF_SYNC2);
      phase_select<C, 3, SEL_MLP>(L, a.work, hset, sel, tc, meta, sh, blk, nb);
      BP_PROF(PROF_P3);
      grid_barrier(gs.counter, gs.epoch, nb, &ss.abort);
      if (multi) xrank_hist_allreduce(a.peer, gs.xseq, hset + kHist1Bins + 2 * kHist2Bins, 2 * kHist3Bins, sh, gs.counter, gs.epoch, nb, &ss.abort);
      BP_PROF(PROF_SYNC3);
      sigma = finish_scale<C>(a.work, hset, sel, sh, &lo, &hi);
      n = sel->n;
      BP_PROF(PROF_SCALE);
    }
    // Next bracket, centred on this median (every CTA computes the same).  Half-width: at least twice the distance
    // the median just moved; otherwise sized from the measured candidate density so that ~500 values fall inside.
    // Off the critical path: done by the LAST thread (its warp owns the fewest points in P4) and not followed by a barrier
    // of its own -- the state is next read at the top of the next linearize, with P4's and the exchange's barriers in between.
    if (tid == BP_BRACKET_THREAD) {
      const float mid_new = 0.5f * (lo + hi), mid_old = 0.5f * (ss.br_lo + ss.br_hi);
      float rel = 0.06f;           // first bracket of a level: wide (the median still moves by percents), the overflow list absorbs it
      if (ss.br_on && mid_new > 0.0f) {
        const float moved = fabsf(mid_new - mid_old) / mid_new;
        if (br.on && ncand > 0) ss.br_density = (float) ncand / fmaxf(ss.br_rel, 1e-6f);       // candidates per unit of rel
        const float rel_density = (ss.br_density > 0.0f) ? (float) BP_BRACKET_TARGET / ss.br_density : 0.002f;
        rel = fmaxf(2.0f * moved, fminf(rel_density, 0.02f));
        rel = fminf(fmaxf(rel, 1e-5f), 0.06f);
      }
      ss.br_rel = rel; ss.br_lo = lo; ss.br_hi = hi; ss.br_on = (n >= 3) ? 1 : 0;
      if (hit) ss.lin.pad[1] = 3;
      if (a.prof && blk == 0) { long long* sp = prof_smem(); sp[12] += hit ? 1 : 0; sp[13] += 1; sp[14] += (br.on && !hit && ncand >= kCandPoison) ? 1 : 0; sp[15] += (br.on && !hit && ncand < kCandPoison) ? 1 : 0; }
    }
    if (BP_BRACKET_THREAD == 0) __syncthreads();
    BP_FINE(39);
  }   // else: P4 reads only what the SAME thread wrote in P1, no grid-wide dependency
  const double mine = stream2 ? phase_reduce_stream2<C>(L, a.work, sigma, loss, a.sp.good_threshold, tc, meta, sh, blk, nb)
                              : phase_reduce<C>(L, a.work, sigma, loss, a.sp.good_threshold, tc, meta, sh, blk, nb, false);
  BP_PROF(PROF_P4);
  sigma_out = sigma; do_hist_out = do_hist;
  return mine;
}

// =============================================================================================
// Tail of one linearize inside the persistent kernel: CTA totals -> grid totals -> LinOut, 6x6 solve, pose update.
//
// Exchange of the 30 sums ("hop B").  Usual path: ONE L2 round trip.  Every CTA converts its fp64 totals to 47-bit fixed point
// (per-scalar power-of-two scales, see below) and adds them with 64-bit integer atomics to 32 accumulator words; the same
// atomic carries the arrival count in the word's top byte, so a word validates itself: a reader polls until the byte says
// "all nb CTAs are in" and has the total -- no fence, no flag word, no leader hop.  Integer addition is associative: the
// totals are bit-identical in every CTA and from run to run whatever the arrival order.  Words are never reset: a reader
// subtracts (mod 2^64) the complete value it saw at the word's previous use; two sets alternate, because a fast CTA may add
// to exchange n + 1 while a slow one still polls exchange n (it cannot reach n + 2 before everybody has added to n + 1).
// Scales: |CTA total of H_ab| <= sqrt(H_aa H_bb), |G_a| <= sqrt(H_aa e) (Cauchy-Schwarz; all terms of H_aa and e = sum w r^2
// are non-negative, so a CTA's share is bounded by the grid total), taken from the PREVIOUS evaluation with a factor 16 in
// hand; resolution 2^-42 of that bound per CTA -- far below the fp32 noise of the per-thread sums feeding it.  A value out
// of range (or NaN) anywhere is reported through word 31 and every CTA repeats that exchange with the fp64 flag-in-data
// mailboxes (exchange_sums), which also serve the first evaluation of a level, grids above 255 CTAs and the multi-rank mode.
// =============================================================================================
#ifndef BP_TAIL_FIXED
#define BP_TAIL_FIXED 1
#endif

constexpr int kFixBits = 46;
struct FixAcc {                  // lives in the registers of warp 0, lane k = scalar k
  unsigned long long prev[2];    // complete value of this lane's word of either set after its last use
  int exp;                       // fixed-point exponent of this lane's scalar for the NEXT exchange
};
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long) (e + 1023) << 52); }   // e in [-1022, 1023]

// warp 0: returns false when some CTA reported an out-of-range value (then `total` is meaningless)
__device__ __forceinline__ bool fixed_exchange(double mine, FixAcc& fa, unsigned long long* acc, unsigned fxn, int nb, AbortCtl* abort_flag, double& total) {
  const int lane = threadIdx.x & 31, set = (int) (fxn & 1u);
  unsigned long long* w = acc + set * 32 + lane;
  const double scaled = mine * pow2d(fa.exp);
  const bool is_data = lane < 30;
  const bool in_range = fabs(scaled) < 70368744177664.0;      // 2^46; false for NaN
  const unsigned any_bad = __ballot_sync(0xffffffffu, is_data && !in_range);
  unsigned long long add = 1ull << 56;
  if (is_data) add += (unsigned long long) ((in_range ? __double2ll_rn(scaled) : 0ll) + (1ll << kFixBits));
  else if (lane == 31) add += any_bad ? 1ull : 0ull;
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(w), "l"(add) : "memory");
  unsigned long long v, d;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
    d = v - fa.prev[set];
    if ((unsigned) (d >> 56) == (unsigned) nb) break;
    if (wait_expired(spins, abort_flag)) break;
  }
  fa.prev[set] = v;
  const long long sum = (long long) (d & ((1ull << 56) - 1)) - (is_data ? (long long) nb << kFixBits : 0ll);
  total = (double) sum * pow2d(-fa.exp);
  return __shfl_sync(0xffffffffu, (int) sum, 31) == 0;
}

// warp 0, every lane redundantly after the gather (lane k < 30 passes the grid total of scalar k): LinOut, the scale
// estimator's state, the exponents of the next fixed-point exchange, and -- unless `solve` is off (parity hook) or the first
// evaluation already meets the gradient tolerance -- the 6x6 solve, the pose update and the next projection matrix
__device__ __forceinline__ void warp0_finish(double total, float sigma, bool do_hist, bool first, bool solve, const SolveArgs& a,
                                             const LevelTemplate& L, const TemplateMeta& meta, SolveShared& ss, FixAcc& fa) {
  const int lane = threadIdx.x & 31;
  const float tf = (float) total;
  float H[36], G[6];
  {
    int k = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) { const float v = __shfl_sync(0xffffffffu, tf, k++); H[c * 6 + r] = v; H[r * 6 + c] = v; }
#pragma unroll
    for (int q = 0; q < 6; ++q) G[q] = __shfl_sync(0xffffffffu, tf, 21 + q);
  }
  // position of this lane's scalar in the upper triangle (lanes < 21), exponents for the next fixed-point exchange
  int rr = 0, cc = 0;
  {
    int kk = lane;
#pragma unroll
    for (int r = 0; r < 6; ++r) { const int len = 6 - r; if (kk >= 0 && kk < len) { rr = r; cc = r + kk; } kk -= len; }
  }
  {
    const int bits = (int) ((unsigned long long) __double_as_longlong(fabs(total)) >> 52);
    const int my_e = max(bits - 1023, -200);                                   // floor(log2 |total|); 0 / denormal -> -200
    const int ga = lane - 21;
    const int src1 = (lane < 21) ? rr * 6 - rr * (rr - 1) / 2 : (lane < 27) ? ga * 6 - ga * (ga - 1) / 2 : 27;
    const int src2 = (lane < 21) ? cc * 6 - cc * (cc - 1) / 2 : 27;
    const int e1 = __shfl_sync(0xffffffffu, my_e, src1), e2 = __shfl_sync(0xffffffffu, my_e, src2);
    fa.exp = (lane < 28) ? min(max(kFixBits - 4 - ((e1 + e2 + 3) >> 1), -900), 900) : 0;      // counts (28, 29) are exact integers
  }
  if (lane < 21) { ss.lin.H[cc * 6 + rr] = tf; ss.lin.H[rr * 6 + cc] = tf; }
  else if (lane < 27) ss.lin.G[lane - 21] = tf;
  else if (lane == 27) { ss.lin.f_norm = sqrtf(tf); ss.lin.sigma = sigma; }
  else if (lane == 28) ss.lin.n_good = (int) (total + 0.5);
  else if (lane == 29) ss.lin.n_valid = (int) (total + 0.5);
  float g_norm = 0.0f;
#pragma unroll
  for (int k = 0; k < 6; ++k) g_norm = fmaxf(g_norm, fabsf(G[k]));
  if (lane == 0) {
    if (do_hist) { ss.delta = fabsf(sigma - ss.scale); ss.scale = sigma; }
    ss.gn = g_norm; ss.fx_bad = 0;
  }
  if (!solve) return;
  if (first && g_norm < a.sp.gradient_tolerance * fmaxf(g_norm, sqrtf(FLT_EPSILON))) return;      // pose_estimator_base.h:346-357: no step
  // unpivoted LDL^T first (H is SPD and Hartley-normalised); Eigen-order paths and the damped fp64 retry when it is rejected
  float dp[6], Pn[12];
  bool ok = solve6_fast(H, G, dp);
  M44 Tn = ss.Td;
  apply_update(Tn, dp, meta.s, meta.c1, meta.c2, meta.c3); make_projection(L, Tn, Pn);          // :371 / :390
  if (!ok) {                                                    // rare: out of line, on the shared-memory copy of H, G
    __syncwarp();
    ok = solve6_cold(ss.lin.H, ss.lin.G, ss.dp);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; ++k) dp[k] = ss.dp[k];
    Tn = ss.Td;
    if (ok) { apply_update(Tn, dp, meta.s, meta.c1, meta.c2, meta.c3); make_projection(L, Tn, Pn); }
  }
  float dpn = 0.0f;
#pragma unroll
  for (int k = 0; k < 6; ++k) dpn += dp[k] * dp[k];
  __syncwarp();                                                 // every lane has read ss.Td
  if (lane == 0) {
    ss.lin.pad[0] = ok ? 1 : 0;
    ss.dpn = sqrtf(dpn);
#pragma unroll
    for (int k = 0; k < 6; ++k) ss.dp[k] = dp[k];
    if (ok) {
      ss.Td = Tn;
#pragma unroll
      for (int k = 0; k < 12; ++k) ss.P[k] = Pn[k];
    }
  }
}

template <int C, int BLEND, bool PEER, int FIX>
__global__ void __launch_bounds__(kLinThreads, 1) k_estimate_pose(const __grid_constant__ SolveArgs a, Sel* sel, int cache_bytes) {
  __shared__ LinShared sh;
  __shared__ SolveShared ss;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  // dynamic shared memory: [candidate scratch kScratchBytes][template cache: cache_bytes, planned per level]
  unsigned* scratch = reinterpret_cast<unsigned*>(dyn_smem);
  const int tid = threadIdx.x;
  GridSync gs;
  gs.counter = a.work.hist + (size_t) kHistSets * kHistWords; gs.epoch = 0; gs.seq = a.seq_base; gs.xseq = a.peer.xseq_base; gs.hs = 0; gs.fxn = 0;
  int total_evals = 0;
  // accumulator words of the fixed-point exchange (behind the barrier counter's 8 words; 2 sets x 32 x u64, never reset): every
  // CTA notes their current values BEFORE its first grid barrier, i.e. before anybody can add to them
  unsigned long long* fix_acc = reinterpret_cast<unsigned long long*>(gs.counter + 8);
  FixAcc fa; fa.exp = 0; fa.prev[0] = fa.prev[1] = 0;
  if (tid < 32) { fa.prev[0] = __ldcg(fix_acc + tid); fa.prev[1] = __ldcg(fix_acc + 32 + tid); }
  if (a.prof && blockIdx.x == 0) {
    long long* sp = prof_smem();
    if (tid < 68) sp[tid] = 0;
    __syncthreads();
    if (tid == 0) sp[64] = clock64();
    BP_FINE_INIT(a.prof);
  }
  if (tid == 0) {
    ss.T = a.chain ? *a.T_out : a.T_init; ss.fx_bad = 0; ss.dpn = 0.0f; ss.gn = 0.0f;
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    ss.abort.flag = 0; ss.abort.deadline = t0 + a.timeout_ns; ss.abort.global_flag = reinterpret_cast<int*>(gs.counter + 1);
  }
  __syncthreads();
  const float sqrt_eps = sqrtf(FLT_EPSILON);
  for (int lvl = a.lvl_first; lvl >= a.lvl_last; --lvl) {
    if (a.dbg.n > 0 && lvl != a.dbg.level) continue;                           // parity hook: one level only
    const LevelTemplate& L = a.tmpl[lvl];
    // PoseEstimatorBase::run (pose_estimator_base.h:324-407); all control flow below is CTA-uniform AND grid-uniform
    if (tid == 0) {                                                            // reset() :287-293
      ss.scale = 1.0f; ss.delta = 1e10f; ss.br_on = 0; ss.br_lo = ss.br_hi = 0.0f; ss.br_rel = 0.06f; ss.br_density = 0.0f;
      ss.Td = ss.T; make_projection(L, ss.Td, ss.P);
    }
    __syncthreads();
    int n_evals = 0, it = 0, status = 0x33;
    float f_prev = 0.0f, g_tol = 0.0f, dp_prev = 0.0f, f_norm, g_norm;
    bool solver_error = false, early = false;
    const TemplateMeta meta = *L.meta;
    const bool multi_rank = PEER && a.peer.nranks > 1 && !meta.replicated;
    // template cache for this level: as many whole fields as fit (tpl_cache_plan)
    const TplCache tc = tpl_cache_plan<C>((unsigned) kScratchBytes, cache_bytes, (meta.n + (int) gridDim.x * kLinThreads - 1) / ((int) gridDim.x * kLinThreads));
    tc_fill<C>(tc, L, meta.n, blockIdx.x, gridDim.x);
    // every histogram set starts the level zeroed; the barrier orders the zeroing before the first atomics
    for (int b = blockIdx.x * kLinThreads + tid; b < kHistSets * kHistWords; b += gridDim.x * kLinThreads) a.work.hist[b] = 0;
    grid_barrier(gs.counter, gs.epoch, gridDim.x, &ss.abort);
    unsigned long long t_level = 0;
    if (blockIdx.x == 0 && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_level));
    long long prof_snap = 0;
    if (a.prof_lvl && blockIdx.x == 0 && tid < 16) prof_snap = prof_smem()[tid];
    if (meta.n_total == 0) {                              // "you should call setData before calling computeResiduals" (template_data.cc:177)
      if (blockIdx.x == 0 && tid == 0) { LevelStats st; st.num_iterations = 0; st.final_error = -1.0f; st.first_order_optimality = -1.0f; st.status = -3; st.num_evals = 0; st.us = 0.0f; a.stats[lvl] = st; }
      continue;
    }

    // One linearize site for the whole of run(): pass "first" is the evaluation before the do-while of :373-393, every
    // later pass is runIteration (pose_estimator_gn.h:83-100).  The pose update, and the projection matrix of the NEXT
    // linearize, are computed by thread 0 right after the solve.
    bool first = true, conv = false;
    for (;;) {
      if (a.dbg.n > 0) {                                                       // parity hook: the caller's pose, no solve
        if (tid == 0) { ss.Td = a.dbg.poses[n_evals]; make_projection(L, ss.Td, ss.P); }
        __syncthreads();
      }
      float sigma; bool do_hist;
      const double mine = device_linearize<C, BLEND, PEER, FIX>(a, lvl, ss, sh, tc, meta, scratch, gs, sel, sigma, do_hist); ++n_evals;
      // grid totals, LinOut, solve, pose update: warp 0 (see warp0_finish); the other warps wait at the barrier
      const bool dbg = a.dbg.n > 0;
      bool done = false;
      if (BP_TAIL_FIXED && n_evals > 1 && gridDim.x <= 255 && !multi_rank) {
        if (tid < 32) {
          double total;
          const bool ok = fixed_exchange(mine, fa, fix_acc, gs.fxn, (int) gridDim.x, &ss.abort, total);
          BP_PROF(PROF_SYNC4);
          if (ok) warp0_finish(total, sigma, do_hist, first, !dbg, a, L, meta, ss, fa);
          else if (tid == 0) ss.fx_bad = 1;
        }
        gs.fxn += 1;
        __syncthreads();
        done = !ss.fx_bad;
      }
      if (!done) {
        exchange_sums(mine, a.work.ll, gs.seq, sh, blockIdx.x, gridDim.x, &ss.abort);
        if (multi_rank) {        // rank totals -> totals over all ranks (summed in rank order by CTA 0, then handed to the other CTAs)
          uint4* lb = a.peer.lbox + (size_t) (gs.seq & 1u) * 64 + 32;
          if (blockIdx.x == 0) {
            const double t = x_allreduce_f64_ordered(a.peer, gs.xseq, tid < 30 ? sh.red[0][tid] : 0.0, &ss.abort);
            if (tid < 30) { ll_post(lb + tid, t, gs.seq); sh.red[0][tid] = t; }
          } else {
            if (tid < 30) sh.red[0][tid] = ll_gather<1>(lb + tid, 0, 1, gs.seq, &ss.abort);
          }
          __syncthreads();
        }
        BP_PROF(PROF_SYNC4);
        if (tid < 32) warp0_finish(tid < 30 ? sh.red[0][tid] : 0.0, sigma, do_hist, first, !dbg, a, L, meta, ss, fa);
        __syncthreads();
      }
      gs.seq += 1;
      gs.hs = (gs.hs + 1) % kHistSets;
      BP_PROF(PROF_FINAL);
      f_norm = ss.lin.f_norm;
      if (dbg) {
        if (blockIdx.x == 0 && tid == 0) a.dbg.out[n_evals - 1] = ss.lin;
        g_norm = 0.0f; status = 0x33; it = n_evals;
        __syncthreads();
        if (n_evals < a.dbg.n) continue;
        early = true; break;
      }
      if (first) {
        g_norm = ss.gn;
        g_tol = a.sp.gradient_tolerance * fmaxf(g_norm, sqrt_eps);
        if (g_norm < g_tol) { status = 0x32; it = 1; early = true; break; }    // :346-357
      }
      BP_FINE(28);
      if (!ss.lin.pad[0]) {                                                    // :359-365 / gn.h:90-97
        status = 0x34; solver_error = true;
        if (first) { early = true; it = 0; g_norm = 0.0f; }
        break;
      }
      if (first) { first = false; f_prev = 0.0f; dp_prev = 0.0f; }
      else if (!(it++ < a.sp.max_iterations && n_evals < a.sp.max_fun_evals)) break;   // the do-while condition (conv is false here)
      // top of the next do-while pass: testConvergence (:258-282) on the step just taken
      const float dpn = ss.dpn;
      g_norm = ss.gn;
      if (dpn < a.sp.parameter_tolerance || dpn < a.sp.parameter_tolerance * (sqrt_eps + dp_prev)) { status = 0x30; conv = true; }
      else if (f_norm < a.sp.function_tolerance || f_norm < a.sp.function_tolerance * (sqrt_eps + f_prev) ||
               fabsf(f_norm - f_prev) < a.sp.function_tolerance) { status = 0x31; conv = true; }
      else if (g_norm < g_tol) { status = 0x32; conv = true; }
      if (a.dbg.trace && blockIdx.x == 0 && tid == 0) {
        const int row = atomicAdd(a.dbg.trace_rows, 1);
        if (row < a.dbg.trace_cap) {
          float* t = a.dbg.trace + (size_t) row * kTraceCols;
          t[0] = (float) lvl; t[1] = (float) n_evals; t[2] = f_norm; t[3] = dpn; t[4] = g_norm; t[5] = ss.lin.sigma; t[6] = (float) ss.lin.pad[1]; t[7] = (float) status;
        }
      }
      dp_prev = dpn; f_prev = f_norm;
      BP_FINE(29);
      BP_PROF(PROF_SOLVE);
      if (conv) {                                                              // the converged pass still applies dp once more (Q1, :390)
        if (tid == 0) apply_update(ss.Td, ss.dp, meta.s, meta.c1, meta.c2, meta.c3);
        __syncthreads();
        it++;                                                                  // `while (it++ < ...)` is evaluated on the way out
        break;
      }
    }
    if (!early) {
      if (!solver_error && tid == 0) ss.T = ss.Td;                             // :395-396
      __syncthreads();
      it -= 1;                                                                 // :398
    }
    total_evals += n_evals;
    // residuals / valid flags of the last linearize of the finest level go back to global memory for getWeights() & co.
    if (tc.f[TC_R] != kTcNone && C != 1 && (lvl == a.sp.max_test_level || a.dbg.n > 0)) {
      int k = 0;
      for (int i = first_point(blockIdx.x, gridDim.x); i < meta.n; i += gridDim.x * kLinThreads, ++k) {
        VecC<C> r; tc_get<C>(tc, k, TC_R, r);
        r.store(a.work.res + (size_t) i * kStride<C>);
        a.work.valid[i] = tc_valid(tc, k);
      }
    }
    if (ss.abort.flag || *(volatile int*) ss.abort.global_flag) status = -4;
    if (blockIdx.x == 0 && tid == 0) {
      LevelStats st; st.num_iterations = it; st.final_error = f_norm; st.first_order_optimality = g_norm; st.status = status; st.num_evals = n_evals;
      unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      st.us = 1e-3f * (float) (t1 - t_level);
      a.stats[lvl] = st;
    }
    if (a.prof_lvl && blockIdx.x == 0) {              // per-level share of the phase counters (thread 0 wrote them; barriers in between)
      __syncthreads();
      if (tid < 16) a.prof_lvl[lvl * 16 + tid] += prof_smem()[tid] - prof_snap;
    }
  }
  if (a.prof && blockIdx.x == 0) {
    __syncthreads();
    if (tid < 64) a.prof[tid] += prof_smem()[tid];
  }
  if (blockIdx.x == 0 && tid == 0) {
    *a.T_out = ss.T; *a.num_fun_evals = total_evals + (a.chain ? *a.num_fun_evals : 0);
    *a.aborted_out = ss.abort.flag | *(volatile int*) ss.abort.global_flag | (a.chain ? *a.aborted_out : 0);
    *a.work.out = ss.lin;
    a.work.scale->scale = ss.scale; a.work.scale->delta = ss.delta;
  }
}

}  // namespace bp
