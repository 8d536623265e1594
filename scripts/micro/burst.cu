// Microbenchmark: every thread of every CTA issues NL independent strong loads (ld.relaxed.gpu) of the SAME small
// region back to back, then consumes them -- the access pattern of bracket_select() right after the grid barrier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o burst burst.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_strong(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_cg(const unsigned* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned ld_plain(const unsigned* p) { return *(const volatile unsigned*) p; }

template <int NL, int MODE> __global__ void k(const unsigned* base, int rounds, long long* cyc, unsigned* sink) {
  const int tid = threadIdx.x;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    unsigned v[NL];
#pragma unroll
    for (int q = 0; q < NL; ++q) { const unsigned* p = base + tid + q * 256; v[q] = MODE == 0 ? ld_strong(p) : MODE == 1 ? ld_cg(p) : ld_plain(p); }
#pragma unroll
    for (int q = 0; q < NL; ++q) acc += v[q];
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
}
template <int NL, int MODE> void run(int grid, const unsigned* buf, long long* cyc, unsigned* sink) {
  const int rounds = 200;
  k<NL, MODE><<<grid, 256>>>(buf, 10, cyc, sink);
  k<NL, MODE><<<grid, 256>>>(buf, rounds, cyc, sink);
  long long h[1024]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("grid %3d  %2d loads/thread (%5d B region) %s : %7.0f cycles per round\n", grid, NL, NL * 1024, MODE == 0 ? "ld.relaxed.gpu" : MODE == 1 ? "ld.cg        " : "ld.volatile  ", (double) mx / rounds);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* buf; cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 1, 1 << 20);
  long long* cyc; cudaMalloc(&cyc, 1024 * sizeof(long long)); unsigned* sink; cudaMalloc(&sink, 4);
  for (int grid : {1, 16, sms}) {
    run<1, 0>(grid, buf, cyc, sink); run<4, 0>(grid, buf, cyc, sink); run<8, 0>(grid, buf, cyc, sink); run<12, 0>(grid, buf, cyc, sink); run<16, 0>(grid, buf, cyc, sink);
    run<12, 1>(grid, buf, cyc, sink); run<12, 2>(grid, buf, cyc, sink);
  }
  return 0;
}
