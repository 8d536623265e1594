// engine_internal.h -- the opaque objects behind include/bpvo_b200.h
#pragma once

#include <stdarg.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/bpvo_b200.h"
#include "device_types.h"

namespace bp { struct Sel; }

struct LevelGeom {
  int rows, cols;
  float fx, fy, cx, cy;   // K_l = K / 2^l with K(2,2) = 1 (vo_frame.cc:24-28)
  float Bf;               // b_l * fx_l (level invariant: baseline doubles as fx halves)
  int capacity;           // upper bound of template points at this level (selection window)
};

struct Mailbox {          // results of a linearize / estimate_pose: one device copy (d_mail), one pinned host copy (h_mail), ONE D2H
  bp::LinOut lin;
  bp::M44 T;
  bp::LevelStats stats[bp::kMaxLevels];
  int evals;
  int aborted;            // an in-kernel wait of the on-device GN loop expired (rank-wide abort word)
};

constexpr int kTraceRows = 8192;   // rows of the GN-loop trace buffer (bpvo_b200_debug_set_trace)

struct bpvo_b200_frame;

struct bpvo_b200_ctx {
  bpvo_b200_params p;
  int rows = 0, cols = 0, L = 0, C = 1, CS = 1;     // C channels, stored with a stride of CS floats (device_types.h: kStride)
  float* plane[5] = {};                              // f32 scratch planes of the gradient-based descriptors
  float K[9]; float baseline = 0;
  LevelGeom geom[bp::kMaxLevels];
  int sm_count = 0; bool coop = false; int smem_optin = 0;
  cudaStream_t stream = nullptr;
  // the disparity map of a frame is not needed before a template is built from it (key-frames only): its upload runs on a second
  // stream, beside the pyramid / descriptor kernels and the solve of the frame.  Every consumer and every host-side wait on
  // `stream` first makes `stream` wait for `disp_done` (bp_join_copies / bp_sync_stream).
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t disp_done = nullptr, order_ev = nullptr;
  bool disp_pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, stage_free = nullptr, tm0 = nullptr, tm1 = nullptr;
  int level_evals[bp::kMaxLevels] = {}; float level_us[bp::kMaxLevels] = {};
  bp::Work work{};
  bp::Sel* sel = nullptr;
  float* export_buf = nullptr;
  uint8_t* flags = nullptr; uint8_t* blur_tmp = nullptr; int* block_counts = nullptr; double* hpartials = nullptr;
  Mailbox* d_mail = nullptr;     // work.out, T, stats, evals, aborted live here
  long long* d_prof = nullptr;
  unsigned long long timeout_ns = 2000000000ull;   // deadline of the in-kernel waits (60 s once peer-memory mode is on)
  float* d_trace = nullptr; int* d_trace_rows = nullptr;   // bpvo_b200_debug_set_trace
  Mailbox* h_mail = nullptr;
  uint8_t* stage_img = nullptr; float* stage_disp = nullptr;
  void* flush_buf = nullptr;
  const bpvo_b200_frame* last_ref = nullptr; int last_level = 0;
  bool profiling = false;
  int solver_ctas = 0;           // CTAs (= SMs) of the persistent solve; 0 = all (bpvo_b200_set_solver_ctas, throughput mode)
  unsigned ll_seq = 0;           // next sequence number of the persistent kernel's exchanges
  bpvo_b200_counters counters{};
  // multi-GPU
  int shard_rank = 0, shard_size = 1;
  void* comm = nullptr;          // ncclComm_t
  double* comm_buf = nullptr;    // device staging for the exchanges (sharded mode)
  double* hsums = nullptr;       // [4] Hartley phase totals
  // peer-memory mode: the on-device GN loop exchanges over NVLink through IPC-mapped mailboxes (comm.cu)
  bool peer_mode = false;
  int shard_min_points = 131072; // peer mode: levels with fewer points are replicated on every rank instead of sharded
  uint2* xbox = nullptr;         // this rank's mailbox
  uint2* xpeer[bp::kXRanks] = {};
  uint4* lbox = nullptr;
  unsigned x_seq = 0, x_seq_init = 0;
};

struct bpvo_b200_frame {
  bpvo_b200_ctx* ctx = nullptr;
  bool has_data = false, has_template = false;
  uint8_t* pyr[bp::kMaxLevels] = {};
  float* desc[bp::kMaxLevels] = {};
  float* saliency[bp::kMaxLevels] = {};
  float* disp = nullptr;
  float4* pts[bp::kMaxLevels] = {};
  float* gx[bp::kMaxLevels] = {};
  float* gy[bp::kMaxLevels] = {};
  float* i0[bp::kMaxLevels] = {};
  int* inds[bp::kMaxLevels] = {};
  // TMA descriptors of pyr[l] / desc[l] for the bit-planes descriptor kernel (tma_ok: encoded successfully)
  CUtensorMap* d_maps = nullptr;        // device copy: [level][in, out]
  bool tma_ok = false;
  bp::TemplateMeta* d_meta = nullptr;
  bp::TemplateMeta* h_meta = nullptr;   // pinned mirror, valid after meta_ready
  cudaEvent_t meta_ready = nullptr;
  // CUDA graphs of this frame's fixed launch sequences: [0] pyramid + descriptors (setData), [1] template build
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
  int graph_launches[2] = {0, 0};
  bool graph_failed[2] = {false, false};
};

int bp_fail(int code, const char* fmt, ...);
cudaError_t bp_join_copies(bpvo_b200_ctx* c);      // `stream` waits for the pending disparity upload (no host wait)
cudaError_t bp_sync_stream(bpvo_b200_ctx* c);      // bp_join_copies + cudaStreamSynchronize(stream)
int bp_hartley(bpvo_b200_ctx* c, bpvo_b200_frame* f, int level);
int bp_comm_destroy(bpvo_b200_ctx* c);
int bp_comm_allreduce_u32(bpvo_b200_ctx* c, unsigned* buf, size_t count);
int bp_comm_allreduce_f64(bpvo_b200_ctx* c, double* buf, size_t count);
