// comm.cu -- multi-GPU exchange for the point-sharded mode (no reference counterpart).
//
// Every rank holds the full moving-frame descriptors and a contiguous scan-order block of the template points.
// Per GN iteration the ranks exchange: the level-1/2/3 radix histograms of |r| (exact global median) and the
// 30 fp64 normal-equation sums.  All of it goes through NCCL all-reduce on the ctx stream (NVLink 5 / NVSwitch).
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch already loaded, or the system one), so the
// library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include "engine_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;                       // ncclSuccess == 0
enum { ncclUint32 = 3, ncclFloat64 = 8 };        // ncclDataType_t values of nccl.h (2.x ABI)
enum { ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
  if (!api.handle) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank)) dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy)) dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce)) dlsym(api.handle, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString)) dlsym(api.handle, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  return api;
}

int nccl_fail(const char* what, ncclResult_t r) {
  return bp_fail(BPVO_B200_ERR_COMM, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
}

}  // namespace

static void peer_release(bpvo_b200_ctx* c) {
  for (int r = 0; r < bp::kXRanks; ++r) {
    if (c->xpeer[r] && c->xpeer[r] != c->xbox) cudaIpcCloseMemHandle(c->xpeer[r]);
    c->xpeer[r] = nullptr;
  }
  if (c->xbox) { cudaFree(c->xbox); c->xbox = nullptr; }
  if (c->lbox) { cudaFree(c->lbox); c->lbox = nullptr; }
  c->peer_mode = false; c->x_seq = 0;
}

int bp_comm_destroy(bpvo_b200_ctx* c) {
  peer_release(c);
  if (c->comm) { nccl().CommDestroy((ncclComm_t) c->comm); c->comm = nullptr; }
  if (c->comm_buf) { cudaFree(c->comm_buf); c->comm_buf = nullptr; }
  c->shard_rank = 0; c->shard_size = 1;
  return BPVO_B200_OK;
}

int bp_comm_allreduce_u32(bpvo_b200_ctx* c, unsigned* buf, size_t count) {
  if (c->shard_size <= 1) return BPVO_B200_OK;
  ncclResult_t r = nccl().AllReduce(buf, buf, count, ncclUint32, ncclSum, (ncclComm_t) c->comm, c->stream);
  if (r != 0) return nccl_fail("ncclAllReduce(u32)", r);
  return BPVO_B200_OK;
}

int bp_comm_allreduce_f64(bpvo_b200_ctx* c, double* buf, size_t count) {
  if (c->shard_size <= 1) return BPVO_B200_OK;
  ncclResult_t r = nccl().AllReduce(buf, buf, count, ncclFloat64, ncclSum, (ncclComm_t) c->comm, c->stream);
  if (r != 0) return nccl_fail("ncclAllReduce(f64)", r);
  return BPVO_B200_OK;
}

extern "C" {

int bpvo_b200_comm_unique_id(uint8_t id[128]) {
  if (!id) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null id");
  if (!nccl().ok) return bp_fail(BPVO_B200_ERR_COMM, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  ncclResult_t r = nccl().GetUniqueId(&u);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id, u.internal, 128);
  return BPVO_B200_OK;
}

int bpvo_b200_comm_init(bpvo_b200_ctx* c, int rank, int nranks, const uint8_t id[128]) {
  if (!c || rank < 0 || nranks < 1 || rank >= nranks) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad rank / nranks");
  bp_comm_destroy(c);
  if (nranks == 1) return BPVO_B200_OK;
  if (!id) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null id");
  if (!nccl().ok) return bp_fail(BPVO_B200_ERR_COMM, "libnccl.so.2 could not be loaded");
  if (cudaSetDevice(c->p.device_id) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaSetDevice failed");
  ncclUniqueId u; memcpy(u.internal, id, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = nccl().CommInitRank(&comm, nranks, u, rank);
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  c->comm = comm; c->shard_rank = rank; c->shard_size = nranks;
  if (cudaMalloc(&c->comm_buf, 64 * sizeof(double)) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaMalloc(comm_buf) failed");
  return BPVO_B200_OK;
}

// ---- peer-memory mode (on-device GN loop across GPUs) -----------------------------------------------------------------
static const size_t kXBoxBytes = (size_t) 2 * bp::kXRanks * bp::kXWords * sizeof(uint2);

int bpvo_b200_peer_export(bpvo_b200_ctx* c, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!c || !handle) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (cudaSetDevice(c->p.device_id) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaSetDevice failed");
  if (!c->xbox) {
    if (cudaMalloc(&c->xbox, kXBoxBytes) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaMalloc(mailbox) failed");
    if (cudaMalloc(&c->lbox, 2 * 64 * sizeof(uint4)) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaMalloc(lbox) failed");
    cudaMemset(c->xbox, 0, kXBoxBytes); cudaMemset(c->lbox, 0, 2 * 64 * sizeof(uint4));
    cudaDeviceSynchronize();
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->xbox);
  if (e != cudaSuccess) return bp_fail(BPVO_B200_ERR_COMM, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(handle, &h, 64);
  return BPVO_B200_OK;
}

int bpvo_b200_peer_init(bpvo_b200_ctx* c, const uint8_t* handles) {
  if (!c || !handles) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null argument");
  if (c->shard_size <= 1) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "peer_init needs comm_init with nranks > 1 first");
  if (c->shard_size > bp::kXRanks) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "peer-memory mode supports up to %d ranks", bp::kXRanks);
  if (!c->xbox) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "call peer_export first");
  if (!c->coop) return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "device without cooperative launch");
  if (cudaSetDevice(c->p.device_id) != cudaSuccess) return bp_fail(BPVO_B200_ERR_CUDA, "cudaSetDevice failed");
  for (int r = 0; r < c->shard_size; ++r) {
    if (r == c->shard_rank) { c->xpeer[r] = c->xbox; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, handles + (size_t) r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return bp_fail(BPVO_B200_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e)); }
    c->xpeer[r] = (uint2*) p;
  }
  c->peer_mode = true; c->x_seq = c->x_seq_init ? c->x_seq_init : 1;
  if (!getenv("BPVO_B200_TIMEOUT_MS")) c->timeout_ns = 60000000000ull;      // a peer PROCESS may be seconds behind
  return BPVO_B200_OK;
}

int bpvo_b200_peer_set_min_points(bpvo_b200_ctx* c, int min_points) {
  if (!c || min_points < 0) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad argument");
  c->shard_min_points = min_points;
  return BPVO_B200_OK;
}

int bpvo_b200_comm_destroy(bpvo_b200_ctx* c) { if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx"); return bp_comm_destroy(c); }

}  // extern "C"
