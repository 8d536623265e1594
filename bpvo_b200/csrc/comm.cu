// comm.cu -- multi-GPU exchange for the point-sharded mode.
// Round-1 state: the rendezvous API is in place; the sharded exchange itself is not wired yet,
// so comm_init with nranks > 1 reports BPVO_B200_ERR_UNSUPPORTED (replica mode -- one independent
// ctx per GPU, no exchange -- is what bench.py --gpus N runs).
#include <string.h>
#include "engine_internal.h"

int bp_comm_destroy(bpvo_b200_ctx* c) { (void) c; return BPVO_B200_OK; }
int bp_comm_allreduce_linout(bpvo_b200_ctx* c) { (void) c; return BPVO_B200_OK; }

extern "C" {
int bpvo_b200_comm_unique_id(uint8_t id[128]) { if (!id) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null id"); memset(id, 0, 128); return BPVO_B200_OK; }
int bpvo_b200_comm_init(bpvo_b200_ctx* c, int rank, int nranks, const uint8_t id[128]) {
  (void) id;
  if (!c || rank < 0 || nranks < 1 || rank >= nranks) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "bad rank / nranks");
  if (nranks == 1) { c->shard_rank = 0; c->shard_size = 1; return BPVO_B200_OK; }
  return bp_fail(BPVO_B200_ERR_UNSUPPORTED, "point-sharded exchange is not wired yet; run one independent ctx per GPU");
}
int bpvo_b200_comm_destroy(bpvo_b200_ctx* c) { if (!c) return bp_fail(BPVO_B200_ERR_INVALID_ARG, "null ctx"); return bp_comm_destroy(c); }
}
