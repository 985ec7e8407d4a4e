// strided_read.cu -- how fast can a B200 stream the gradient planes of the batched lookup backward?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_read strided_read.cu && ./strided_read
//
// The backward reads T tensors [B][NC][HW] fp32; a CTA owns PX consecutive pixels of one batch
// element and reads, per iteration t, NC row pieces of PX*4 bytes that lie HW*4 bytes apart.  This
// probe times exactly that access pattern with nothing else going on, for PX = 32 / 64 / 128
// (128 / 256 / 512-byte pieces), fed by (a) a TMA producer thread and an S-deep ring, (b) plain
// coalesced loads, and a plain contiguous read of the same bytes as the yardstick.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int T = 22, B = 8, NC = 36, H = 80, W = 184, HW = H * W;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Maps { CUtensorMap m[T]; };

// (a) TMA producer + ring; one consumer warp touches one word per stage and releases it
template <int S>
__global__ void __launch_bounds__(64) tile_tma_kernel(const __grid_constant__ Maps maps, int px, float* sink) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
  float* ring = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_raw) + 127) & ~(uintptr_t)127);
  const int stage_floats = NC * px;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int hw0 = blockIdx.x * px, b = blockIdx.y;
  if (tid == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&s_empty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    if (lane == 0) {
      for (int t = 0; t < T; ++t) {
        const int slot = t % S;
        if (t >= S) mbar_wait(smem_u32(&s_empty[slot]), (uint32_t)((t / S) - 1) & 1u);
        const uint32_t bar = smem_u32(&s_full[slot]);
        mbar_expect_tx(bar, stage_floats * 4);
        tma_load_3d(smem_u32(ring + (size_t)slot * stage_floats), &maps.m[t], bar, hw0, 0, b);
      }
    }
  } else {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const int slot = t % S;
      mbar_wait(smem_u32(&s_full[slot]), (uint32_t)(t / S) & 1u);
      acc += ring[(size_t)slot * stage_floats + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[slot]));
    }
    if (acc == 123.456f) sink[0] = acc;
  }
}

// (b) plain loads: 4 warps, warp w reads channels w, w+4, ...; VEC floats per lane (px = 32*VEC)
template <int VEC, int UNROLL>
__global__ void __launch_bounds__(128) tile_ldg_kernel(const float* __restrict__ base, float* sink) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px = 32 * VEC;
  const int hw0 = blockIdx.x * px, b = blockIdx.y;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* p = base + ((size_t)t * B + b) * NC * HW + hw0 + lane * VEC;
#pragma unroll
    for (int c0 = 0; c0 < NC; c0 += 4 * UNROLL) {
      float v[UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int c = c0 + 4 * u + warp;
        if (c < NC) {
          const float* q = p + (size_t)c * HW;
          if (VEC == 1) v[u][0] = __ldg(q);
          else if (VEC == 2) { float2 f = __ldg(reinterpret_cast<const float2*>(q)); v[u][0] = f.x; v[u][1 % VEC] = f.y; }
          else { float4 f = __ldg(reinterpret_cast<const float4*>(q)); v[u][0] = f.x; v[u][1 % VEC] = f.y; v[u][2 % VEC] = f.z; v[u][3 % VEC] = f.w; }
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[u][e] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc += v[u][e];
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

// yardstick: contiguous grid-stride read of the same bytes
__global__ void __launch_bounds__(256) contig_kernel(const float4* __restrict__ p, size_t n4, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
  }
  for (; i < n4; i += stride) acc += __ldg(p + i).x;
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
static float time_it(F launch, int reps = 10) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 2; ++i) launch(i);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch(i);
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms * 1e3f / reps;
}

int main() {
  const size_t n = (size_t)T * B * NC * HW;          // floats of one gradient set (373 MB)
  const int NSETS = 2;                               // rotate: 746 MB >> L2
  float* buf[NSETS];
  for (int s = 0; s < NSETS; ++s) { CK(cudaMalloc(&buf[s], n * 4)); CK(cudaMemset(buf[s], 0, n * 4)); }
  float* sink; CK(cudaMalloc(&sink, 4));
  const double bytes = (double)n * 4;
  printf("pattern: T=%d tensors [B=%d][NC=%d][HW=%d] fp32 = %.0f MB per pass, %d sets rotated\n", T, B, NC, HW, bytes / 1e6, NSETS);

  { float us = time_it([&](int i) { contig_kernel<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(buf[i % NSETS]), n / 4, sink); });
    printf("contiguous read                         : %7.1f us  %5.2f TB/s\n", us, bytes / us / 1e6); }

  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  for (int px : {32, 64, 128}) {
    Maps maps[NSETS];
    for (int s = 0; s < NSETS; ++s)
      for (int t = 0; t < T; ++t) {
        const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)NC, (cuuint64_t)B};
        const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)NC * HW * 4};
        const cuuint32_t box[3] = {(cuuint32_t)px, (cuuint32_t)NC, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        CUresult r = enc(&maps[s].m[t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf[s] + (size_t)t * B * NC * HW, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      }
    const dim3 grid((HW + px - 1) / px, B);
    for (int stages : {3, 6}) {
      for (int pad_kb : {0, 44, 100}) {                // extra shared memory = fewer CTAs per SM (0: as many as fit)
        const size_t smem = (size_t)stages * NC * px * 4 + 128 + (size_t)pad_kb * 1024;
        if (smem > 227 * 1024) continue;
        float us;
        if (stages == 3) {
          CK(cudaFuncSetAttribute(tile_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          us = time_it([&](int i) { tile_tma_kernel<3><<<grid, 64, smem>>>(maps[i % NSETS], px, sink); });
        } else {
          CK(cudaFuncSetAttribute(tile_tma_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          us = time_it([&](int i) { tile_tma_kernel<6><<<grid, 64, smem>>>(maps[i % NSETS], px, sink); });
        }
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 32) per_sm = 32;
        printf("TMA  box {%3d px x %d ch} S=%d smem %3zu KB (~%2d CTA/SM): %7.1f us  %5.2f TB/s\n", px, NC, stages, smem / 1024, per_sm, us, bytes / us / 1e6);
      }
    }
  }
  {
    const dim3 g32((HW + 31) / 32, B), g64((HW + 63) / 64, B), g128((HW + 127) / 128, B);
    float us;
    us = time_it([&](int i) { tile_ldg_kernel<1, 3><<<g32, 128>>>(buf[i % NSETS], sink); });
    printf("LDG  32 px (128 B pieces) 3 loads/warp in flight : %7.1f us  %5.2f TB/s\n", us, bytes / us / 1e6);
    us = time_it([&](int i) { tile_ldg_kernel<1, 9><<<g32, 128>>>(buf[i % NSETS], sink); });
    printf("LDG  32 px (128 B pieces) 9 loads/warp in flight : %7.1f us  %5.2f TB/s\n", us, bytes / us / 1e6);
    us = time_it([&](int i) { tile_ldg_kernel<2, 9><<<g64, 128>>>(buf[i % NSETS], sink); });
    printf("LDG  64 px (256 B pieces) 9 loads/warp in flight : %7.1f us  %5.2f TB/s\n", us, bytes / us / 1e6);
    us = time_it([&](int i) { tile_ldg_kernel<4, 9><<<g128, 128>>>(buf[i % NSETS], sink); });
    printf("LDG 128 px (512 B pieces) 9 loads/warp in flight : %7.1f us  %5.2f TB/s\n", us, bytes / us / 1e6);
  }
  return 0;
}
