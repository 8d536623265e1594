// Microbenchmark of the pipes the residual phase leans on (B200, sm_100a): per-SM throughput in lanes/clk of
// DMUL / DADD / DFMA, f32<->f64 conversions, F2I.F64, and the dependent-issue latency of the same ops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int OP> __device__ __forceinline__ void body(double& a, double& b, float& f, int& i) {
  if (OP == 0) a = __dmul_rn(a, b);
  if (OP == 1) a = __dadd_rn(a, b);
  if (OP == 2) a = __fma_rn(a, b, b);
  if (OP == 3) { a = (double) f; f = __int_as_float(__double2hiint(a) ^ i); }          // F2F.F64.F32 (+ 1 ALU)
  if (OP == 4) { f = (float) a; a = __hiloint2double(__float_as_int(f) | i, i); }       // F2F.F32.F64 (+ ALU)
  if (OP == 5) { i = (int) a; a = __hiloint2double(i, i); }                             // F2I.S32.F64
  if (OP == 6) { a = (double) i; i = __double2loint(a) ^ __double2hiint(a); }           // I2F.F64.S32
  if (OP == 7) a = __ddiv_rn(b, a);
  if (OP == 8) f = __fmaf_rn(f, f, f);
}

template <int OP, int ILP> __global__ void k(double* out, int iters, double seed) {
  double a[ILP], b = seed; float f[ILP]; int i[ILP];
  for (int q = 0; q < ILP; ++q) { a[q] = seed + q + threadIdx.x * 1e-3; f[q] = 1.0f + q * 0.25f; i[q] = q + 1; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) body<OP>(a[q], b, f[q], i[q]);
  }
  double s = 0; for (int q = 0; q < ILP; ++q) s += a[q] + f[q] + i[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int ILP> void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, (size_t) sms * 1024 * sizeof(double));
  const int iters = 4096, threads = warps_per_sm * 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP, ILP><<<sms, threads>>>(out, 64, 1.0000001);
  cudaEventRecord(e0); k<OP, ILP><<<sms, threads>>>(out, iters, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * mhz * 1e3;
  printf("%-22s warps/SM %2d ILP %d : %7.2f cycles per op-per-warp-slot, %6.1f lanes/clk/SM\n", name, warps_per_sm, ILP,
         cycles / iters / ILP, (double) threads * ILP * iters / cycles);
  cudaFree(out);
}

int main() {
  run<0, 1>("DMUL latency", 1);   run<0, 8>("DMUL throughput", 16);
  run<1, 1>("DADD latency", 1);   run<1, 8>("DADD throughput", 16);
  run<2, 1>("DFMA latency", 1);   run<2, 8>("DFMA throughput", 16);
  run<3, 1>("F2F.64.32+ALU lat", 1); run<3, 8>("F2F.64.32 thr", 16);
  run<4, 1>("F2F.32.64+ALU lat", 1); run<4, 8>("F2F.32.64 thr", 16);
  run<5, 1>("F2I.F64 lat", 1);    run<5, 8>("F2I.F64 thr", 16);
  run<6, 1>("I2F.F64 lat", 1);    run<6, 8>("I2F.F64 thr", 16);
  run<7, 1>("DDIV latency", 1);   run<7, 4>("DDIV throughput", 16);
  run<8, 1>("FFMA latency", 1);   run<8, 8>("FFMA throughput", 16);
  return 0;
}
