// gather_probe.cu -- what bounds the lookup's scattered window reads when the pyramid exceeds L2?
// Standalone microbenchmark (not part of the library).  Rows of PITCH bytes, one row per pixel;
// pixel (y,x) reads SEG bytes starting at 16B-aligned offset ~ (x - disp(y,x))*4 in its row.
//   mode 0: 8 lanes per pixel, one 16-byte ld.global.L1::no_allocate.L2::64B per lane (as the pair kernel)
//   mode 1: same with plain ld.global.nc (128-byte fills)
//   mode 2: one thread per pixel issues ONE cp.async.bulk (SEG bytes) into shared memory
//   mode 3: 8 lanes per pixel, cp.async 16B .L2::64B into shared memory, PPT pixels per lane group in flight
//   layout 0: row-major rows (pixel-private rows, pitch PITCH)         -- the current layout
//   layout 1: "skewed tiles": the windows of 8 consecutive pixels are contiguous (8*SEG bytes)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct P {
  const uint8_t* vol; float* out; const float* disp;
  int W, HW; long long pitch; int seg; int layout;
};

__device__ __forceinline__ long long win_addr(const P& p, long long pix, int x, float d) {
  int start = (int)floorf((float)x - d) - 4;
  if (start < 0) start = 0;
  int off = (start * 4) & ~15;                              // 16B aligned
  if (off + p.seg > p.pitch) off = (int)p.pitch - p.seg;
  if (p.layout == 0) return pix * p.pitch + off;
  // skewed tiles: group of 8 pixels; their windows start at the SAME skewed offset -> contiguous
  const long long grp = pix >> 3;
  return grp * (8 * p.pitch) + (long long)(off / 16) * (8 * 16) * 1 + 0 * (pix & 7) +      // tile of 8 px x 16 B
         0;                                                                              // (see lane term below)
}

template <int MODE>
__global__ void __launch_bounds__(256) k_lanes(P p) {
  // 8 lanes per pixel, 32 pixels per CTA
  const int tid = threadIdx.x, q = tid & 7, pl = tid >> 3;
  const long long pix = (long long)blockIdx.x * 32 + pl;
  if (pix >= p.HW) return;
  const int x = (int)(pix % p.W);
  const float d = __ldg(p.disp + pix);
  float acc = 0.f;
  const int nch = p.seg / 16;
  if (q < nch) {
    long long a;
    if (p.layout == 0) a = win_addr(p, pix, x, d) + q * 16;
    else {
      int start = (int)floorf((float)x - d) - 4; if (start < 0) start = 0;
      int ch = (start * 4) >> 4; if ((ch + nch) * 16 > p.pitch) ch = (int)(p.pitch / 16) - nch;
      // tile layout: [group of 8 px][chunk][px in group][16 B]
      a = (pix >> 3) * (8 * p.pitch) + ((long long)(ch + q) * 8 + (pix & 7)) * 16;
    }
    uint4 t;
    if (MODE == 0)
      asm volatile("ld.global.L1::no_allocate.L2::64B.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p.vol + a));
    else
      t = __ldg(reinterpret_cast<const uint4*>(p.vol + a));
    acc = __uint_as_float(t.x) + __uint_as_float(t.y) + __uint_as_float(t.z) + __uint_as_float(t.w);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (q == 0) p.out[pix] = acc;
}

// one thread per pixel, one bulk copy each; 128 pixels per CTA
__global__ void __launch_bounds__(128) k_bulk(P p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const long long pix = (long long)blockIdx.x * 128 + tid;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(128));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem) + tid * p.seg;
  if (pix < p.HW) {
    const int x = (int)(pix % p.W);
    const float d = __ldg(p.disp + pix);
    long long a;
    if (p.layout == 0) a = win_addr(p, pix, x, d);
    else {   // skewed rows: all 8 px of a group contiguous per chunk is not bulk-friendly; use [group][px][seg window region]
      int start = (int)floorf((float)x - d) - 4; if (start < 0) start = 0;
      int off = (start * 4) & ~15; if (off + p.seg > p.pitch) off = (int)p.pitch - p.seg;
      a = pix * p.pitch + off;
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(p.seg) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(p.vol + a), "r"(p.seg), "r"(bar_a) : "memory");
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(done) : "r"(bar_a) : "memory");
  if (pix < p.HW) {
    float acc = 0.f;
    const float* s = reinterpret_cast<const float*>(smem + tid * p.seg);
    for (int i = 0; i < p.seg / 4; ++i) acc += s[i];
    p.out[pix] = acc;
  }
}

// 8 lanes per pixel, cp.async 16B with L2::64B, PPT pixels per lane group
template <int PPT>
__global__ void __launch_bounds__(256) k_cpasync(P p) {
  __shared__ __align__(16) uint8_t smem[PPT * 32 * 8 * 16];
  const int tid = threadIdx.x, q = tid & 7, pl = tid >> 3;
  const int nch = p.seg / 16;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const long long pix = ((long long)blockIdx.x * PPT + i) * 32 + pl;
    if (pix < p.HW && q < nch) {
      const int x = (int)(pix % p.W);
      const float d = __ldg(p.disp + pix);
      const long long a = win_addr(p, pix, x, d) + q * 16;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem) + ((i * 32 + pl) * 8 + q) * 16;
      asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(dst), "l"(p.vol + a) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const long long pix = ((long long)blockIdx.x * PPT + i) * 32 + pl;
    float acc = 0.f;
    if (pix < p.HW && q < nch) {
      const float4 t = *reinterpret_cast<const float4*>(smem + ((i * 32 + pl) * 8 + q) * 16);
      acc = t.x + t.y + t.z + t.w;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (q == 0 && pix < p.HW) p.out[pix] = acc;
  }
}

// evict the volume from L2 with CLEAN lines (a memset would leave 126 MB of dirty lines whose
// write-back then competes with the gather)
__global__ void k_flush_read(const uint4* src, long long n, float* sink) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint4 t = __ldg(src + i);
    acc += __uint_as_float(t.x ^ t.y ^ t.z ^ t.w);
  }
  if (acc == 123.456f) *sink = acc;
}

// persistent variant of k_lanes<0>: grid = SMs x 8, each lane group walks pixels with UNR loads in flight
template <int UNR>
__global__ void __launch_bounds__(256) k_lanes_persist(P p) {
  const int tid = threadIdx.x, q = tid & 7, pl = tid >> 3;
  const int nch = p.seg / 16;
  const long long stride = (long long)gridDim.x * 32;
  for (long long pix0 = (long long)blockIdx.x * 32 + pl; pix0 < p.HW; pix0 += stride * UNR) {
    uint4 t[UNR];
    bool ok[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long pix = pix0 + u * stride;
      ok[u] = pix < p.HW && q < nch;
      t[u] = make_uint4(0, 0, 0, 0);
      if (ok[u]) {
        const int x = (int)(pix % p.W);
        const float d = __ldg(p.disp + pix);
        const long long a = win_addr(p, pix, x, d) + q * 16;
        asm volatile("ld.global.L1::no_allocate.L2::64B.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(t[u].x), "=r"(t[u].y), "=r"(t[u].z), "=r"(t[u].w) : "l"(p.vol + a));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      float acc = __uint_as_float(t[u].x) + __uint_as_float(t[u].y) + __uint_as_float(t[u].z) + __uint_as_float(t[u].w);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const long long pix = pix0 + u * stride;
      if (q == 0 && pix < p.HW) p.out[pix] = acc;
    }
  }
}

int main(int argc, char** argv) {
  const int B = 8, H = 96, W = 312;
  const int HW = B * H * W;
  const long long pitch = 2432;                    // cfg3 fused row: (312+156+78+39 -> aligned) * 4
  uint8_t* vol; float* out; float* disp;
  CK(cudaMalloc(&vol, (size_t)HW * pitch + 4096));
  CK(cudaMemset(vol, 0, (size_t)HW * pitch + 4096));
  CK(cudaMalloc(&out, HW * sizeof(float) * 4));
  float* hd = (float*)malloc(HW * sizeof(float));
  srand(1);
  for (int r = 0; r < B * H; ++r) {
    float base = 5.f + 60.f * (rand() / (float)RAND_MAX);
    for (int x = 0; x < W; ++x) hd[r * W + x] = base + 3.f * (rand() / (float)RAND_MAX) + 0.02f * x;
  }
  CK(cudaMalloc(&disp, HW * sizeof(float)));
  CK(cudaMemcpy(disp, hd, HW * sizeof(float), cudaMemcpyHostToDevice));
  uint8_t* flush; CK(cudaMalloc(&flush, 256 << 20));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* only = getenv("GP_ONLY");
  const int hw_div = getenv("GP_DIV") ? atoi(getenv("GP_DIV")) : 1;
  const bool noflush = getenv("GP_NOFLUSH") != nullptr;
  auto run = [&](const char* name, int seg, int layout, auto launch) {
    if (only && !strstr(name, only)) return;
    P p{vol, out, disp, W, HW / hw_div, pitch, seg, layout};
    // kernel time = (REPS x (flush + kernel) - REPS x flush) / REPS: no event / launch-gap overhead
    const int REPS = 10;
    auto flush_once = [&]() {
      if (!noflush) k_flush_read<<<148 * 8, 256>>>(reinterpret_cast<const uint4*>(flush), (256 << 20) / 16, out + 2 * HW);
    };
    float t_flush = 0.f, t_both = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
      flush_once(); launch(p);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int it = 0; it < REPS; ++it) { flush_once(); if (pass) launch(p); }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(pass ? &t_both : &t_flush, e0, e1));
    }
    CK(cudaGetLastError());
    const double us = (t_both - t_flush) / REPS * 1e3;
    printf("%-34s seg=%3d layout=%d: %.1f us  useful %.0f GB/s\n", name, seg, layout, us, (double)p.HW * seg / (us * 1e-6) / 1e9);
  };
  for (int layout = 0; layout < 2; ++layout)
    for (int seg : {48, 96, 128}) {
      run("lanes L2::64B", seg, layout, [&](P p) { k_lanes<0><<<(p.HW + 31) / 32, 256>>>(p); });
      run("lanes ld.nc", seg, layout, [&](P p) { k_lanes<1><<<(p.HW + 31) / 32, 256>>>(p); });
    }
  for (int seg : {48, 96, 128, 256}) {
    run("bulk 1/px", seg, 0, [&](P p) { k_bulk<<<(p.HW + 127) / 128, 128, 128 * seg>>>(p); });
  }
  for (int seg : {48, 96, 128}) {
    run("cp.async x1", seg, 0, [&](P p) { k_cpasync<1><<<(p.HW + 31) / 32, 256>>>(p); });
    run("cp.async x2", seg, 0, [&](P p) { k_cpasync<2><<<(p.HW + 63) / 64, 256>>>(p); });
    run("cp.async x4", seg, 0, [&](P p) { k_cpasync<4><<<(p.HW + 127) / 128, 256>>>(p); });
  }
  for (int seg : {48, 96, 128}) {
    run("persist x1", seg, 0, [&](P p) { k_lanes_persist<1><<<148 * 8, 256>>>(p); });
    run("persist x2", seg, 0, [&](P p) { k_lanes_persist<2><<<148 * 8, 256>>>(p); });
    run("persist x4", seg, 0, [&](P p) { k_lanes_persist<4><<<148 * 8, 256>>>(p); });
    run("persist x4 layout1", seg, 1, [&](P p) { k_lanes_persist<4><<<148 * 8, 256>>>(p); });
  }
  // two segments per pixel as the pair kernel does (level 0 + level 2 regions of the same row)
  return 0;
}
