// tma_offset_probe.cu -- may the innermost start coordinate of a tiled TMA load be any element index, or
// must coord0 * sizeof(T) be a multiple of 16 bytes?   ./tma_offset_probe <offset in elements> [swizzle: 0|1]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, float* out) {
  __shared__ __align__(1024) float tile[8 * 32];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(8 * 32 * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(&m), "r"(b), "r"(c0), "r"(c1) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(b) : "memory");
  out[threadIdx.x] = tile[threadIdx.x];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int off = argc > 1 ? atoi(argv[1]) : 0, swz = argc > 2 ? atoi(argv[2]) : 0;
  const int W = 1500, H = 64;
  float* d; CK(cudaMalloc(&d, W * H * 4));
  float* h = (float*)malloc(W * H * 4);
  for (int i = 0; i < W * H; ++i) h[i] = (float)i;
  CK(cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out, 256 * 4));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap m;
  const cuuint64_t gdim[2] = {W, H}; const cuuint64_t gstr[1] = {W * 4};
  const cuuint32_t box[2] = {32, 8}; const cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&m, swz == 2 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  k<<<1, 256>>>(m, off, 3, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("offset %d swizzle %d: %s\n", off, swz, cudaGetErrorString(e)); return 0; }
  float o[256]; CK(cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost));
  printf("offset %d swizzle %d: ok, tile[0] = %.0f (expected %d)%s\n", off, swz, o[0], 3 * W + off, swz ? " (swizzled layout: element 0 of row 0 stays in place)" : "");
  return 0;
}
