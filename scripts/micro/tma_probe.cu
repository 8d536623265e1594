// tma_probe.cu -- bisects which tensor-map / box configurations the TMA load and store accept on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu && ./tma_probe <variant>
// Each variant runs in its own process (a faulting kernel poisons the context).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("FAIL %s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() { unsigned p; asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(p)); return p != 0; }

// load a (bw x bh)-byte box at (cx, cy) into shared memory, copy it out
__global__ void k_load2d(const CUtensorMap* map, int cx, int cy, int bytes, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"((unsigned long long) map), "r"(cx), "r"(cy), "r"(b) : "memory");
  }
  unsigned done = 0; int spins = 0;
  while (!done && ++spins < (1 << 22)) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? smem[i] : 0xEE;
}
// same with the descriptor as a __grid_constant__ parameter
__global__ void k_load2d_param(const __grid_constant__ CUtensorMap map, int cx, int cy, int bytes, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned b = smem_u32(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"((unsigned long long) &map), "r"(cx), "r"(cy), "r"(b) : "memory");
  }
  unsigned done = 0; int spins = 0;
  while (!done && ++spins < (1 << 22)) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? smem[i] : 0xEE;
}
// fill shared memory with a pattern and store it as a 3-D float box {8, bw, bh} at (0, cx, cy)
__global__ void k_store3d(const CUtensorMap* map, int cx, int cy, int floats) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < floats; i += blockDim.x) s[i] = (float) i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"((unsigned long long) map), "r"(0), "r"(cx), "r"(cy), "r"(smem_u32(smem)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const char* v = argc > 1 ? argv[1] : "A";
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  if (!fn) { printf("FAIL no entry point\n"); return 1; }
  EncodeFn encode = (EncodeFn) fn;
  const int rows = 96, cols = 128, pitch = 128;
  std::vector<unsigned char> h(rows * pitch);
  for (int i = 0; i < rows * pitch; ++i) h[i] = (unsigned char) ((i * 7 + i / pitch) & 0xff);
  unsigned char* d_img; CK(cudaMalloc(&d_img, rows * pitch)); CK(cudaMemcpy(d_img, h.data(), rows * pitch, cudaMemcpyHostToDevice));
  unsigned char* d_out; CK(cudaMalloc(&d_out, 1 << 16)); CK(cudaMemset(d_out, 0, 1 << 16));
  CUtensorMap* d_map; CK(cudaMalloc(&d_map, sizeof(CUtensorMap)));
  CUtensorMap map; memset(&map, 0, sizeof(map));
  if (v[0] != 'S') {
    int bw = 80, bh = 14, cx = -3, cy = -3; bool param = false; int gcols = cols;
    if (!strcmp(v, "B")) { cx = 0; cy = 0; }
    if (!strcmp(v, "C")) { bw = 128; bh = 16; cx = 0; cy = 0; }
    if (!strcmp(v, "D")) { bw = 64; bh = 8; cx = 0; cy = 0; }
    if (!strcmp(v, "E")) { bw = 64; bh = 8; cx = -3; cy = -3; }
    if (!strcmp(v, "F")) { bw = 80; bh = 14; cx = 0; cy = 0; param = true; }
    if (!strcmp(v, "G")) { bw = 96; bh = 14; cx = -3; cy = -3; }
    if (!strcmp(v, "H")) { bw = 80; bh = 14; cx = -3; cy = -3; gcols = 121; }     // extent not a multiple of 16 (pitch is)
    if (!strcmp(v, "I")) { bw = 16; bh = 4; cx = 0; cy = 0; }
    if (!strcmp(v, "J")) { bw = 96; bh = 14; cx = -16; cy = -3; }                 // inner coordinate a multiple of 16 bytes, negative
    if (!strcmp(v, "K")) { bw = 96; bh = 14; cx = 48; cy = 85; }                  // beyond the right / bottom edge
    const cuuint64_t dims[2] = {(cuuint64_t) gcols, (cuuint64_t) rows}; const cuuint64_t strides[1] = {(cuuint64_t) pitch};
    const cuuint32_t box[2] = {(cuuint32_t) bw, (cuuint32_t) bh}, es[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("variant %s: encode FAILED (%d)\n", v, (int) r); return 0; }
    CK(cudaMemcpy(d_map, &map, sizeof(map), cudaMemcpyHostToDevice));
    if (param) k_load2d_param<<<1, 128, bw * bh>>>(map, cx, cy, bw * bh, d_out);
    else k_load2d<<<1, 128, bw * bh>>>(d_map, cx, cy, bw * bh, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %s: box %dx%d at (%d,%d)%s -> KERNEL FAULT: %s\n", v, bw, bh, cx, cy, param ? " [param]" : "", cudaGetErrorString(e)); return 0; }
    std::vector<unsigned char> o(bw * bh); CK(cudaMemcpy(o.data(), d_out, bw * bh, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
      const int gy = cy + y, gx = cx + x;
      const unsigned char want = (gy >= 0 && gy < rows && gx >= 0 && gx < gcols) ? h[gy * pitch + gx] : 0;
      bad += o[y * bw + x] != want;
    }
    printf("variant %s: box %dx%d at (%d,%d)%s -> ran, %d mismatching bytes (first byte 0x%02x)\n", v, bw, bh, cx, cy, param ? " [param]" : "", bad, o[0]);
  } else {
    // store variants: [rows][cols][8] float tensor
    const int orows = 20, ocols = 70;
    float* d_desc; CK(cudaMalloc(&d_desc, (size_t) orows * ocols * 8 * 4)); CK(cudaMemset(d_desc, 0, (size_t) orows * ocols * 8 * 4));
    int bw = 64, bh = 8, cx = 0, cy = 0;
    if (!strcmp(v, "S2")) { cx = 32; cy = 16; }       // partly outside: clipped
    if (!strcmp(v, "S3")) { bw = 32; }
    const cuuint64_t dims[3] = {8, (cuuint64_t) ocols, (cuuint64_t) orows}; const cuuint64_t strides[2] = {32, (cuuint64_t) ocols * 32};
    const cuuint32_t box[3] = {8, (cuuint32_t) bw, (cuuint32_t) bh}, es[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_desc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("variant %s: encode FAILED (%d)\n", v, (int) r); return 0; }
    CK(cudaMemcpy(d_map, &map, sizeof(map), cudaMemcpyHostToDevice));
    k_store3d<<<1, 128, 8 * bw * bh * 4>>>(d_map, cx, cy, 8 * bw * bh);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %s: store box 8x%dx%d at (%d,%d) -> KERNEL FAULT: %s\n", v, bw, bh, cx, cy, cudaGetErrorString(e)); return 0; }
    std::vector<float> o((size_t) orows * ocols * 8); CK(cudaMemcpy(o.data(), d_desc, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int y = 0; y < orows; ++y) for (int x = 0; x < ocols; ++x) for (int c = 0; c < 8; ++c) {
      const int by = y - cy, bx = x - cx;
      const float want = (by >= 0 && by < bh && bx >= 0 && bx < bw) ? (float) ((by * bw + bx) * 8 + c) : 0.0f;
      bad += o[((size_t) y * ocols + x) * 8 + c] != want;
    }
    printf("variant %s: store box 8x%dx%d at (%d,%d) -> ran, %d mismatching floats\n", v, bw, bh, cx, cy, bad);
  }
  return 0;
}
