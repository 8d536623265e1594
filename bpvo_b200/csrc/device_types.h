// device_types.h -- structures shared between the host engine and the sm_100a kernels.
#pragma once

#include <stdint.h>
#include <cuda_runtime.h>
#include "host_math.h"

namespace bp {

constexpr int kMaxLevels = 16;
constexpr int kHist1Bins = 2048;   // |r| float bits [30:20]
constexpr int kHist2Bins = 2048;   // bits [19:9]
constexpr int kHist3Bins = 512;    // bits [8:0]
constexpr int kHistBins = kHist1Bins + 2 * kHist2Bins + 2 * kHist3Bins;
constexpr int kSelBins = 1024, kSelList = 256;   // bracketed median: linear bins over the bracket, short list of the two wanted bins
constexpr int kHistWords = kHistBins + 8 + kSelBins;   // one histogram set + bookkeeping words + the bracket histogram:
//   [kHistBins+0] max(~i) over valid points i   [+1] valid points   [+2] residuals below the bracket   [+3] residuals inside it
//   [+4] length of the candidate overflow list   [+5] length of the short list of a sliced overflow scan
//   [+8 ...] kSelBins counts of the candidates by linear bin over the bracket
#ifndef BP_CAND_PER_CTA
#define BP_CAND_PER_CTA 12
#endif
constexpr int kCandPerCta = BP_CAND_PER_CTA;             // bracketed median: every CTA owns a fixed region of the candidate buffer (no slot reservation) ...
constexpr int kCtaCandCap = 2048;            // ... stages up to this many candidates in shared memory ...
constexpr int kOvfCap = 262144;              // ... and appends what exceeds its region to a shared overflow list (wide brackets, early iterations,
                                             //     megapixel levels: a 0.2 % bracket around the median of 1.5e7 residuals holds 3e4 of them)
constexpr int kOvfLocalScan = 2048;          // overflow lists up to this length are scanned by every CTA; longer ones slice by slice (bracket_select)
constexpr unsigned kCandPoison = 1u << 30;   // added to the candidate count when even that overflowed -> radix fallback
constexpr int kScratchBytes = (kCtaCandCap + kSelList) * 4;   // dynamic-smem scratch of the bracketed select: CTA candidates | list
constexpr int kPartialStride = 32; // doubles per block partial (21 H + 6 G + f + n_good + pad)
constexpr int kHistSets = 4;       // histogram sets of the on-device loop: a set is zeroed two GN iterations (>= two grid-wide exchanges) before its next use
constexpr int kMsgWords = 8;       // 16-byte flag-in-data words of a CTA's median message: {valid points, below} {candidates, c0} {c1, c2} ... {c11, c12}
constexpr int kMsgCand = 2 * kMsgWords - 3;   // candidates a message carries (13)
#ifndef BP_LL_GROUP
#define BP_LL_GROUP 24        /* same-box A/B over 8 / 12 / 16 / 24 / 37 / 74 / 148: 24 is the fastest (by 0.4 % over 12) */
#endif
constexpr int kLLGroup = BP_LL_GROUP;      // CTAs per group of the two-level flag-in-data exchange of the normal-equation sums
constexpr int kMaxGrid = 1024;     // upper bound of CTAs of the persistent kernel (sizes the exchange mailboxes)
constexpr int kLinThreads = 256;   // threads per CTA of the linearize phases (1 CTA per SM; measured: 128 -> 14.7, 256 -> 12.4, 512 -> 13.5 us / GN iteration)

// Channel-interleaved arrays (descriptor [rows][cols][S], template fields and residuals [N][S]) store S = kStride<C> floats per
// pixel / point: 1, 4 (C = 3: intensity + gradient), 8 (C = 5: descriptor fields, C = 8: bit-planes), so that a record is one or
// two aligned 16-byte loads whatever C; the padding channels are zero and never enter a sum.
template <int C> constexpr int kStride = (C == 1) ? 1 : ((C <= 4) ? 4 : 8);

__host__ __device__ inline int channel_stride(int C) { return (C == 1) ? 1 : ((C <= 4) ? 4 : 8); }

// Per-level template ("TemplateData", bpvo/template_data.h) in the device layout:
//   pts   [N]      float4 (X, Y, Z, 1)                                  (reference: _points)
//   gx    [N][C]   fx * Ix of every channel, point-major                (reference: folded into _jacobians)
//   gy    [N][C]   fy * Iy
//   i0    [N][C]   template pixel values                                (reference: _pixels, channel-major)
// The 1x6 Jacobian of a residual is gx*A(X) + gy*B(X) with A, B the rows of the 2x6 warp Jacobian at
// identity (rigid_body_warp.h:94-106); A, B are recomputed in registers from (X,Y,Z).
struct TemplateMeta {     // device-resident header of a level's template (written by the build kernels)
  int n_raw;              // pixels that passed saliency + NMS + disparity gate
  int n;                  // points held by THIS rank (multiple of 16)
  int n_total;            // points kept globally (== n when unsharded)
  int first;              // global scan-order index of this rank's first point
  float s, c1, c2, c3;    // Hartley normalisation: Tn = [sI, -s c; 0 1]  (warps.cc:27-48)
  int replicated;         // multi-GPU: this level is too small to shard -- every rank keeps ALL its points (n == n_total) and
  int pad[3];             // runs it like a single GPU, bit-identically, with no exchange
};

struct LevelTemplate {
  const float4* pts;
  const float* gx;
  const float* gy;
  const float* i0;
  const TemplateMeta* meta;   // n and the normalisation live on the device: set_template never syncs
  float fx, fy, cx, cy;
};

// Per-level dense descriptor of the moving frame: channel-interleaved f32 [rows][cols][C]
struct LevelImage {
  const float* desc;
  int rows, cols;
};

struct ScaleState {      // AutoScaleEstimator state (mestimator.h:62-83)
  float scale;
  float delta;
};

// result of one linearize, written by the last CTA of the reduce phase
struct LinOut {
  float H[36];           // column-major, symmetric
  float G[6];
  float f_norm;          // sqrt(sum w r^2)
  float sigma;
  int   n_valid;         // valid POINTS
  int   n_good;          // weights > goodPointThreshold over all C*N entries (invalid entries count as weight 1, Q6)
  int   pad[2];
};

// scratch owned by a ctx
struct Work {
  float* res;            // [N][C] residuals of the last linearize (point-major; 0 for invalid points)
  uint8_t* valid;        // [N]
  unsigned* hist;        // kHistSets x kHistWords, followed by 8 words ([0] grid-barrier counter of the persistent kernel) and
                         // 2 x 32 u64 accumulator words of its fixed-point exchange (never reset, see fixed_exchange)
  uint4* ll;             // flag-in-data mailboxes of the persistent kernel: [2][kMaxGrid][32] CTA sums, then [2][kMaxGrid][32] group sums
  double* partials;      // [grid][kPartialStride]
  ScaleState* scale;
  LinOut* out;
  unsigned* ticket;      // last-CTA-done counter
  uint4* msg;            // [2][kMaxGrid][kMsgWords] median messages of the CTAs (flag-in-data, double-buffered by sequence parity)
  float* cand;           // [grid][kCandPerCta] |r| values inside the median bracket, -1 = empty slot (on-device GN loop),
                         // followed by the overflow list [kOvfCap] and the short list [kSelList] of a sliced overflow scan
};

// point-sharded multi-GPU mode with the GN loop on the device: every rank owns a mailbox in ITS memory that the peers
// write over NVLink (CUDA IPC mappings); 8-byte flag-in-data words {value, sequence number}
constexpr int kXRanks = 8;
constexpr int kXWords = 4096 + 64;   // 32-bit values per (parity, source rank) slot: the largest message is the level-2 histogram pair
struct PeerArgs {
  int rank, nranks;            // nranks <= 1: single GPU (everything below unused)
  unsigned xseq_base;          // first sequence number of this launch's cross-rank exchanges (identical on every rank)
  uint2* box[kXRanks];         // box[r] = rank r's mailbox [2][kXRanks][kXWords] (box[rank] is local memory)
  uint4* lbox;                 // local broadcast words [2][64]: what rank-wide results CTA 0 hands to the other CTAs
};

struct SolverParams {    // PoseEstimatorParameters (pose_estimator_params.h) + loss
  int   max_iterations;
  int   max_fun_evals;   // 1200
  float parameter_tolerance, function_tolerance, gradient_tolerance;
  int   loss;            // 0x10 huber, 0x11 tukey, 0x12 L2
  int   interp;          // InterpolationType (types.h): 0 linear, 1 cosine, 2 cubic, 3 cubic Hermite
  float good_threshold;
  int   max_test_level;
  int   num_levels;
};

struct LevelStats {      // OptimizerStatistics (types.h:444-482)
  int   num_iterations;
  float final_error;
  float first_order_optimality;
  int   status;
  int   num_evals;       // linearize() calls at this level
  float us;              // device time spent in this level (globaltimer of CTA 0), microseconds
};

// Parity / diagnosis hooks of the persistent kernel (all off in production launches: n == 0, trace == nullptr).
//   poses / out / n / level : instead of the GN loop, evaluate linearize() `n` times at `level` at the caller's poses
//                             (no solve, no pose update; scale state reset before the first) and store every LinOut --
//                             the SAME device_linearize() the GN loop runs (template cache, bracketed median from the
//                             second evaluation on, flag-in-data exchange), comparable entry by entry with the fine seam
//   trace                   : one row of kTraceCols floats per linearize of the GN loop
constexpr int kTraceCols = 8;      // level, eval, f_norm, |dp|, max|G|, sigma, scale re-estimated (0/1) + 2 * bracket hit, status so far
struct DebugArgs {
  const M44* poses;
  LinOut* out;
  int n, level;
  float* trace;
  int trace_cap;         // rows
  int* trace_rows;       // rows written (device word)
};

// everything the on-device GN loop needs for one estimatePose
struct SolveArgs {
  LevelTemplate tmpl[kMaxLevels];
  LevelImage img[kMaxLevels];
  SolverParams sp;
  Work work;
  M44 T_init;
  // outputs (device memory)
  M44* T_out;
  LevelStats* stats;     // [num_levels]
  int* num_fun_evals;
  int* aborted_out;      // receives the rank-wide abort word (an in-kernel wait expired)
  long long* prof;       // optional: per-phase cycle counters of CTA 0 (nullptr = off)
  long long* prof_lvl;   // optional: the same, split by pyramid level [kMaxLevels][16]
  int lvl_first, lvl_last;   // pyramid levels this launch walks, coarse to fine (normally num_levels - 1 .. max_test_level)
  int chain;             // 1 = second launch of a solve split by level: the start pose is *T_out, the evaluation count adds to *num_fun_evals
  unsigned seq_base;     // first sequence number of this launch's exchanges (monotonic across launches of a ctx)
  unsigned long long timeout_ns;   // every in-kernel wait gives up this long after the launch started (AbortCtl)
  PeerArgs peer;
  DebugArgs dbg;
};

}  // namespace bp
