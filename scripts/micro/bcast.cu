// Microbenchmark: cost of an L2 "broadcast" -- every CTA of a full-chip grid reads the SAME words with ld.relaxed.gpu
// (the pattern of the persistent solve's exchanges).  Reports cycles per round for several footprints / layouts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bcast bcast.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_relaxed(const uint4* p) {
  uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// each of the first `readers` threads of every CTA reads `per` words; word w lives at base + w * stride_words (uint4 units)
__global__ void k(const uint4* base, int nwords, int stride, int rounds, long long* cyc, unsigned* sink) {
  const int tid = threadIdx.x;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    for (int w = tid; w < nwords; w += blockDim.x) { const uint4 v = ldg_relaxed(base + (size_t) w * stride); acc += v.x + v.y + v.z + v.w; }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint4* buf; cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 1, 64 << 20);
  long long* cyc; cudaMalloc(&cyc, sms * sizeof(long long));
  unsigned* sink; cudaMalloc(&sink, 4);
  const int rounds = 200;
  struct Cfg { int nwords, stride; const char* what; } cfgs[] = {
    {1, 1, "1 word (16 B)"}, {30, 1, "30 words contiguous (480 B)"}, {30, 16, "30 words, 256-B stride"}, {30, 64, "30 words, 1-KB stride"},
    {13 * 32, 1, "13x32 words contiguous (6.6 KB)"}, {13 * 32, 16, "13x32 words, 256-B stride"},
    {256, 1, "4 KB contiguous"}, {704, 1, "11 KB contiguous"}, {704, 16, "704 words, 256-B stride"}, {2048, 1, "32 KB contiguous"}};
  for (int grid : {1, 13, sms}) {
    for (auto& c : cfgs) {
      k<<<grid, 256>>>(buf, c.nwords, c.stride, 10, cyc, sink);
      k<<<grid, 256>>>(buf, c.nwords, c.stride, rounds, cyc, sink);
      long long h[1024]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("grid %3d  %-34s : %7.0f cycles per round (slowest CTA)\n", grid, c.what, (double) mx / rounds);
    }
  }
  return 0;
}
