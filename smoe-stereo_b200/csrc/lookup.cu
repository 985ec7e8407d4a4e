// lookup.cu -- pyramid lookup (forward), lookup backward and pyramid-gradient fold for sm_100a.
//
// These kernels are HBM/L2-bound gathers and scatters, not GEMMs.  Design:
//  * forward, default: lookup_fwd_px_kernel -- ONE thread per pixel (or per level pair of a
//    pixel) gathers two contiguous segments (levels 0 and 2; levels 1 and 3 are recomputed
//    bit-exactly) with twelve 16-byte loads issued up front, applies the data-dependent window
//    start with a register select network and stores every output channel as one coalesced
//    128-byte line per warp straight from registers (px_gather.cuh).
//  * forward, lane-group kernels (lookup_fwd_vec_kernel: 4 lanes per pixel, four windows;
//    lookup_fwd_pair_kernel: 8 lanes per pixel, two segments): the earlier designs, kept as
//    bit-exact cross-checks behind smoe_corr_debug_set_lookup_kernel; 3.3x more instructions.
//  * backward, batched (default for training): lookup_bwd_lvl_kernel -- warp = level, lane =
//    pixel, the gradient rows of a 32-pixel tile live in shared memory, all iterations of a step
//    stream through registers, the fold runs in shared memory, level 0 is written once.
//  * backward, accumulate (per call): the gradient tile is staged through shared memory with
//    coalesced loads and each lane read-modify-writes the one 16-byte chunk it owns; rows are
//    pixel-private, so no atomics are needed (as in sampler_kernel.cu:102).
//  * backward, overwrite: one pass writes every element of the dense gradient (zeros outside the
//    window), replacing the reference's zeros_like + scatter pair (sampler_kernel.cu:148-163).
//  * generic scalar kernels cover unaligned widths / pitches and radii other than 4.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "px_gather.cuh"

namespace smoe {

void* tensor_map_encoder_ptr();   // build.cu: cuTensorMapEncodeTiled from the driver (no link-time libcuda dependency)

__device__ __forceinline__ uint32_t ptx_smem(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

constexpr int kTilePx = 32;       // pixels per CTA tile (one 128-byte line of fp32 outputs)
constexpr int kFwdThreads = 128;  // 4 lanes per pixel

// A pyramid that cannot stay L2-resident (126 MB L2, minus the lookup's own outputs) is read with
// the streaming load variant (Vec16).
static bool volume_exceeds_l2(const PyrDev& d, long long rows, int elem_bytes) {
  return rows * d.pitch[0] * elem_bytes > (96ll << 20);   // (64 MB tried: worse for a 90 MB fp16 volume)
}

// ------------------------------------------------------------------- forward (thread per pixel)
// The lane-group kernels above spend most of their issue slots on per-pixel scalar work that every
// warp repeats for only 4 (pair kernel: 8 lanes per pixel) or 8 pixels: ncu shows the pair kernel
// at cfg3 with 18.8 M warp-instructions per launch (314 per warp, IPC 2.8 of 4) and it takes 24 us
// even when the volume is L2-resident -- it is issue-bound, not DRAM-bound
// (tools/issue_bound_probe.py, profiles/r01/README.md).  Here ONE thread owns a pixel:
//  * per level pair it issues the NCH aligned 16-byte loads of the level-2pr segment up front
//    (NP*NCH = 12 independent loads in flight per thread for fp32, no shared memory, no shuffles),
//  * the data-dependent window start inside the aligned segment is one of EPL positions: it is
//    applied with log2(EPL) rounds of register selects (no dynamic register indexing),
//  * the coarse level of the pair is recomputed as (a+b)*0.5f like lookup_fwd_pair_kernel
//    (bit-identical to the stored level),
//  * lane == pixel, so every output channel is stored as one coalesced 128-byte line per warp
//    straight from registers.
// ~7x fewer warp-instructions per pixel; the uncoalesced 16-byte loads cost L1 wavefronts
// instead (32 per load instruction), which is not the binding resource.
// SPLIT: blockDim.y = number of level pairs and each thread handles ONE pair of its pixel (half
// the dependent instruction chain and registers per thread, twice the threads): for L2-resident
// volumes, where the kernel is bound by the latency of that chain rather than by throughput.
// LD256: fp32 volumes fetched with 256-bit loads (PxPair256): the default for volumes beyond L2.
// HINT (LD256 + STREAM only): L2 eviction priorities.  A volume that streams from DRAM is dominated
// by level 0 (16 of every 21 pyramid bytes the lookup reads per pixel row), whose segments are
// single-use; level 2 is 4x smaller per row and the same lines are hit again by the next GRU
// iteration.  bits 0-1: 1 = level-0 fills evict_first; 2 = also level-2 fills evict_last;
// 3 = only level-2 evict_last.  bit 2: output stores evict_first (only when nothing reads them back).
__device__ __forceinline__ void st_hint(float* p, float v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(__half* p, __half v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.b16 [%0], %1, %2;" ::"l"(p), "h"(__half_as_ushort(v)), "l"(pol) : "memory");
}

// Measured and rejected in round 2 (profiles/r02/lookup_loop_variant_*.log): the streaming lookup as
// PERSISTENT CTAs that walk their tiles in a grid-stride loop with the next tile's coordinate in
// flight (k tiles = 1 + k dependent DRAM round trips instead of 2k).  Slower at every streaming
// config (cfg2 8.4 -> 10.1 us, cfg3 17.9 -> 20.0 us, cfg4 33.4 -> 37.5 us at 64 registers; worse at
// 128): the hardware's CTA scheduler already overlaps the two phases across the 16 small CTAs an
// SM holds.  In steady state (32 lookups back to back) the kernel moves 59.6 MB of DRAM reads (64-byte
// fills around 80-byte segments; 39 MB useful) + 34.5 MB of output per config-3 launch in 17.9 us =
// 5.3 TB/s, 80 % of the measured copy peak: it is DRAM-bound on the bytes the fill granularity forces.
// Also rejected (profiles/r02/lookup_thread_per_level_variant.log): one thread per output LEVEL for
// L2-resident volumes (both threads of a level pair gather the segment, one interpolates the fine level,
// the other recomputes the coarse one): the cfg1 step went from 134-138 to 157 us.
// Occupancy is not the limiter of the streaming kernel either: capped at 48 registers (21 instead of 16
// CTAs per SM, no spills) it measured cfg2 8.08 / cfg3 17.9 / cfg4 33.9 us against 8.31 / 17.7 / 33.4.
template <typename VT, typename OT, int R, int NL, bool STREAM, bool SPLIT, bool LD256 = false, int HINT = 0>
__global__ void __launch_bounds__(256)
lookup_fwd_px_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                     float coord_scale, OT* __restrict__ out, int HW, CoordUpd upd) {
  static_assert(!LD256 || sizeof(VT) == 4, "256-bit segment loads are implemented for fp32 volumes");
  static_assert(HINT == 0 || (LD256 && STREAM), "eviction hints ride on the streaming 256-bit loads");
  using Pair = typename std::conditional<LD256, PxPair256<STREAM>, PxPair<VT, STREAM>>::type;
  constexpr int kHint0 = ((HINT & 3) == 1 || (HINT & 3) == 2) ? 1 : 0;
  constexpr int kHint2 = (HINT & 3) >= 2 ? 2 : 0;
  constexpr bool kOutEf = (HINT & 4) != 0;
  constexpr int RD = 2 * R + 1;
  constexpr int NP = (NL + 1) / 2;
  constexpr int NPT = SPLIT ? 1 : NP;                         // pairs per thread
  static_assert(R == Pair::R && NP <= 2, "PxPair assumes radius 4; level scales below <= 4 levels");

  pdl_wait();
  const int b = blockIdx.y;
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  const int pr0 = SPLIT ? (int)threadIdx.y : 0;
  if (hw >= HW) return;
  const float x0 = fetch_coord(coords, coords_bstride, upd, b, hw, pr0 == 0) * coord_scale;
  const long long row = (long long)b * HW + hw;

  Pair pair[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int pr = pr0 + i;
    // (selects instead of pyr.x[2*pr]: a runtime index would copy the parameter struct to local memory)
    const bool first = pr == 0 || NP == 1;
    const VT* fine_row = static_cast<const VT*>(first ? pyr.ptr[0] : pyr.ptr[2]) +
                         row * (first ? pyr.pitch[0] : pyr.pitch[2]);
    pair[i].load(reinterpret_cast<const typename std::conditional<LD256, float, VT>::type*>(fine_row),
                 first ? pyr.width[0] : pyr.width[2], x0 * (pr == 0 ? 1.0f : 0.25f),
                 pr == 0 ? kHint0 : kHint2);
  }
  pdl_launch_dependents();

  unsigned long long pol = 0;
  if (kOutEf) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  auto put = [&](OT* p, float v) {
    if (kOutEf) st_hint(p, from_float<OT>(v), pol);
    else *p = from_float<OT>(v);
  };
  OT* dst = out + (long long)b * (NL * RD) * HW + hw;
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int pr = pr0 + i;
    const int lf = 2 * pr;
    float e[RD];
    if (2 * pr + 1 < NL) {
      pair[i].template coarse<OT>((pr == 0 || NL < 4) ? pyr.width[1] : pyr.width[3], e);
      OT* d1 = dst + (long long)(lf + 1) * RD * HW;
#pragma unroll
      for (int k = 0; k < RD; ++k) put(d1 + (long long)k * HW, e[k]);
    }
    pair[i].template fine<OT>(e);
    OT* d0 = dst + (long long)lf * RD * HW;
#pragma unroll
    for (int k = 0; k < RD; ++k) put(d0 + (long long)k * HW, e[k]);
  }
}

// ---------------------------------------------------------------------------- forward (generic)
template <typename VT, typename OT>
__global__ void __launch_bounds__(128)
lookup_fwd_generic_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                          float coord_scale, int R, OT* __restrict__ out, int HW, CoordUpd upd) {
  pdl_launch_dependents();
  pdl_wait();
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (hw >= HW) return;
  const int RD = 2 * R + 1;
  const float x0 = fetch_coord(coords, coords_bstride, upd, b, hw, true) * coord_scale;
  const long long row = (long long)b * HW + hw;
  float lscale = 1.0f;
  for (int l = 0; l < pyr.num_levels; ++l) {
    int base;
    float dx;
    split_coord(x0 * lscale, R, base, dx);
    lscale *= 0.5f;
    const int width = pyr.width[l];
    const VT* src = static_cast<const VT*>(pyr.ptr[l]) + row * pyr.pitch[l];
    OT* dst = out + ((long long)b * pyr.num_levels * RD + (long long)l * RD) * HW + hw;
    float prev = (base >= 0 && base < width) ? to_float(__ldcg(src + base)) : 0.f;
    for (int k = 0; k < RD; ++k) {
      const int i1 = base + k + 1;
      const float nxt = (i1 >= 0 && i1 < width) ? to_float(__ldcg(src + i1)) : 0.f;
      dst[(long long)k * HW] = lerp_ref<OT>(prev, nxt, dx);
      prev = nxt;
    }
  }
}

// ------------------------------------------------------------------------ backward contributions
// Gradient reaching window element t (0..2r+1) of a pixel, in the reference's order
// (sampler_kernel.cu:92-102): g = 0; g += grad[t-1]*dx (t>0); g += grad[t]*(1-dx) (t<2r+1).
template <typename VT> struct BwdTap;
template <> struct BwdTap<float> {
  __device__ static float eval(float gm1, float g0, bool has_m1, bool has_0, float dx) {
    // __fmul_rn: never contracted with the caller's accumulation `x += eval(...)` into an FMA, so
    // the per-call and the batched kernels round identically (the batched one knows the tap
    // index at compile time and would otherwise fuse the last tap: 1-ulp differences)
    float g = has_m1 ? __fmul_rn(gm1, dx) : 0.f;
    if (has_0) g = fmaf(g0, 1.0f - dx, g);
    return g;
  }
};
template <> struct BwdTap<__half> {   // c10::Half arithmetic: fp32 op, round to half after each
  __device__ static float eval(float gm1, float g0, bool has_m1, bool has_0, float dx) {
    const float w0 = __half2float(__float2half_rn(dx));
    const float w1 = __half2float(__float2half_rn(1.0f - dx));
    float g = 0.f;
    if (has_m1) g = __half2float(__float2half_rn(g + __half2float(__float2half_rn(gm1 * w0))));
    if (has_0) g = __half2float(__float2half_rn(g + __half2float(__float2half_rn(g0 * w1))));
    return g;
  }
};

// -------------------------------------------------------------------- backward accumulate (vector)
template <typename GT, int R, int NL>
__global__ void __launch_bounds__(kFwdThreads)
lookup_bwd_acc_vec_kernel(PyrDev gp, const float* __restrict__ coords, long long coords_bstride,
                          float coord_scale, const GT* __restrict__ gout, int HW) {
  constexpr int EPL = 4, LOG_EPL = 2;
  constexpr int RD = 2 * R + 1;
  __shared__ float s_g[NL * RD][kTilePx + 1];

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  {
    const bool ok = hw0 + lane < HW;
    const GT* src = gout + (long long)b * (NL * RD) * HW + hw0 + lane;
#pragma unroll 4
    for (int c = warp; c < NL * RD; c += kFwdThreads / 32)
      s_g[c][lane] = ok ? to_float(src[(long long)c * HW]) : 0.f;
  }
  __syncthreads();

  const int p = tid >> 2, q = tid & 3;
  const int hw = hw0 + p;
  if (hw >= HW) return;
  const float x0 = __ldg(coords + (long long)b * coords_bstride + hw) * coord_scale;
  const long long row = (long long)b * HW + hw;
  float lscale = 1.0f;
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    int base;
    float dx;
    split_coord(x0 * lscale, R, base, dx);
    lscale *= 0.5f;
    const int s = base & (EPL - 1);
    const int ci = (base >> LOG_EPL) + q;
    const int width = gp.width[l];
    const int nchunks = (width + EPL - 1) >> LOG_EPL;
    if (!((q * EPL < s + 2 * R + 2) && ci >= 0 && ci < nchunks)) continue;
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(gp.ptr[l]) + row * gp.pitch[l] +
                                            (long long)ci * EPL);
    const uint4 cu = ld_window16(dst);      // 64-byte fills, not whole lines (see Vec16)
    float4 cur = make_float4(__uint_as_float(cu.x), __uint_as_float(cu.y), __uint_as_float(cu.z),
                             __uint_as_float(cu.w));
    float* c4 = reinterpret_cast<float*>(&cur);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const int t = q * EPL + j - s;
      if (t >= 0 && t <= RD && ci * EPL + j < width) {
        const float gm1 = (t > 0) ? s_g[l * RD + t - 1][p] : 0.f;
        const float g0 = (t < RD) ? s_g[l * RD + t][p] : 0.f;
        c4[j] += BwdTap<float>::eval(gm1, g0, t > 0, t < RD, dx);
      }
    }
    __stcg(dst, cur);
  }
}

// ------------------------------------------------------------------- backward accumulate (generic)
template <typename GT>
__global__ void __launch_bounds__(128)
lookup_bwd_acc_generic_kernel(PyrDev gp, const float* __restrict__ coords, long long coords_bstride,
                              float coord_scale, int R, const GT* __restrict__ gout, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const int hw = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (hw >= HW) return;
  const int RD = 2 * R + 1;
  const float x0 = __ldg(coords + (long long)b * coords_bstride + hw) * coord_scale;
  const long long row = (long long)b * HW + hw;
  float lscale = 1.0f;
  for (int l = 0; l < gp.num_levels; ++l) {
    int base;
    float dx;
    split_coord(x0 * lscale, R, base, dx);
    lscale *= 0.5f;
    const int width = gp.width[l];
    float* dst = static_cast<float*>(gp.ptr[l]) + row * gp.pitch[l];
    const GT* g = gout + ((long long)b * gp.num_levels * RD + (long long)l * RD) * HW + hw;
    float gm1 = 0.f;
    for (int t = 0; t <= RD; ++t) {
      const float g0 = (t < RD) ? to_float(g[(long long)t * HW]) : 0.f;
      const int idx = base + t;
      if (idx >= 0 && idx < width) dst[idx] += BwdTap<float>::eval(gm1, g0, t > 0, t < RD, dx);
      gm1 = g0;
    }
  }
}

// ------------------------------------------------------------------------ backward overwrite (dense)
// One CTA stages the gradients of 32 consecutive pixels, then each warp writes whole gradient
// rows (zeros outside the 2r+2 window).  VEC = elements per store (16 bytes when aligned, else 1).
constexpr int kDenseThreads = 256;

template <typename VT, typename GT, int VEC>
__global__ void __launch_bounds__(kDenseThreads)
lookup_bwd_dense_kernel(PyrDev gp, const float* __restrict__ coords, long long coords_bstride,
                        float coord_scale, int R, const GT* __restrict__ gout, int HW) {
  extern __shared__ float s_dyn[];   // [NL*RD][kTilePx+1] gradients, then [kTilePx] coords
  const int RD = 2 * R + 1;
  const int NC = gp.num_levels * RD;
  float* s_x = s_dyn + NC * (kTilePx + 1);
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  {
    const bool ok = hw0 + lane < HW;
    const GT* src = gout + (long long)b * NC * HW + hw0 + lane;
    for (int c = warp; c < NC; c += kDenseThreads / 32)
      s_dyn[c * (kTilePx + 1) + lane] = ok ? to_float(src[(long long)c * HW]) : 0.f;
    if (warp == 0)
      s_x[lane] = ok ? __ldg(coords + (long long)b * coords_bstride + hw0 + lane) * coord_scale : 0.f;
  }
  __syncthreads();
  for (int p = warp; p < kTilePx; p += kDenseThreads / 32) {
    const int hw = hw0 + p;
    if (hw >= HW) break;
    const long long row = (long long)b * HW + hw;
    float lscale = 1.0f;
    for (int l = 0; l < gp.num_levels; ++l) {
      int base;
      float dx;
      split_coord(s_x[p] * lscale, R, base, dx);
      lscale *= 0.5f;
      const int width = gp.width[l];
      VT* dst = static_cast<VT*>(gp.ptr[l]) + row * gp.pitch[l];
      const float* g = s_dyn + (l * RD) * (kTilePx + 1) + p;
      for (int i0 = lane * VEC; i0 < width; i0 += 32 * VEC) {
        VT vals[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const int t = i0 + j - base;
          float gv = 0.f;
          if (t >= 0 && t <= RD) {
            const float gm1 = (t > 0) ? g[(t - 1) * (kTilePx + 1)] : 0.f;
            const float g0 = (t < RD) ? g[t * (kTilePx + 1)] : 0.f;
            gv = BwdTap<VT>::eval(gm1, g0, t > 0, t < RD, dx);
          }
          vals[j] = from_float<VT>(gv);
        }
        if (VEC == 1) {
          dst[i0] = vals[0];
        } else {
          *reinterpret_cast<uint4*>(dst + i0) = *reinterpret_cast<const uint4*>(vals);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ backward, batched over iterations
// Autograd executes every lookup backward of a step before the build backward and nothing reads
// the gradient pyramid in between, so the T lookups of a step can be back-propagated TOGETHER:
// a CTA keeps the full gradient rows (all levels) of 32 pixels in shared memory, streams the T
// (coords, grad_out) tiles through a kBatchedStages-deep cp.async pipeline (a 2-stage version exposed
// one DRAM latency per iteration: 663 us vs 589 us for the per-call path), scatters into shared memory,
// folds the levels there and writes ONLY the folded level-0 gradient, once.  Compared with T
// read-modify-write passes over a gradient pyramid that does not fit L2 (plus a memset and the
// fold kernel) the DRAM traffic drops from ~T*97 MB + 2*166 MB to T*17 MB + 87 MB at config 2.
// The arithmetic per element is the same as accumulate + fold (same order), fp32, radius 4.
constexpr int kMaxBatchedIters = 48;
constexpr int kBatchedStages = 4;
struct BwdBatch {
  const float* coords[kMaxBatchedIters];
  const void* gout[kMaxBatchedIters];
  long long coords_bstride[kMaxBatchedIters];
  int T;
};
struct BatchedGeom {
  int width[SMOE_MAX_LEVELS];
  int off[SMOE_MAX_LEVELS];     // word offset of each level inside a shared-memory row
  int rowlen;                   // words per pixel row in shared memory (odd multiple of 4 + pad)
  int num_levels;
};

#ifdef SMOE_CROSSCHECK
#include "xcheck_lookup.cuh"
#endif

// Second version of the batched backward: warp = pyramid level, lane = pixel of the 32-pixel tile.
// The gradient rows of a pixel are private to it and the levels occupy disjoint parts of the row,
// so the T iterations need NO barrier and no shared-memory staging of the gradient tiles: each warp
// streams its 9 gradient channels (+ the coordinate) of iteration t straight from global memory
// into registers -- lane = pixel makes every load one coalesced line -- two iterations ahead of
// the one it is scattering.  One coordinate split per (pixel, level, iteration) instead of two,
// no cp.async address arithmetic, no per-iteration __syncthreads: ~2.2x fewer warp-instructions
// than lookup_bwd_batched_kernel, and no 16-byte alignment requirement on the gradient planes.
// Same accumulation order per element (iterations ascending) => bit-identical results.
template <typename GT, int NL>
__global__ void __launch_bounds__(128)
lookup_bwd_lvl_kernel(const __grid_constant__ BwdBatch batch, const BatchedGeom geo,
                      float* __restrict__ g0_out, long long g0_pitch, int HW, int accumulate) {
#ifndef SMOE_BWD_LVL_DEPTH
#define SMOE_BWD_LVL_DEPTH 5
#endif
  constexpr int R = 4, RD = 9, NC = NL * RD, kDepth = SMOE_BWD_LVL_DEPTH;   // register ring: kDepth-1 iterations in flight
  extern __shared__ __align__(16) uint8_t s_raw[];
  float* s_rows = reinterpret_cast<float*>(s_raw);                    // [kTilePx][rowlen]

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int l = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  const int npx = min(kTilePx, HW - hw0);

  for (int i = tid; i < kTilePx * geo.rowlen / 4; i += 128)
    reinterpret_cast<float4*>(s_rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate) {
    __syncthreads();
    for (int pp = 0; pp < npx; ++pp)
      for (int j = tid; j < geo.width[0]; j += 128)
        s_rows[pp * geo.rowlen + j] = g0_out[((long long)b * HW + hw0 + pp) * g0_pitch + j];
  }
  __syncthreads();

  if (l < NL && lane < npx) {
    float* rowl = s_rows + lane * geo.rowlen + geo.off[l];
    const unsigned width = (unsigned)geo.width[l];
    const float lscale = 1.0f / (float)(1 << l);
    const long long gofs = ((long long)b * NC + l * RD) * HW + hw0 + lane;
    float g[kDepth][RD], x[kDepth];
    auto load = [&](int t, float (&gt)[RD], float& xt) {
      const GT* gp = static_cast<const GT*>(batch.gout[t]) + gofs;
#pragma unroll
      for (int k = 0; k < RD; ++k) gt[k] = to_float(__ldg(gp + (long long)k * HW));
      xt = __ldg(batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0 + lane);
    };
    auto scatter = [&](const float (&gt)[RD], float xt) {
      int base;
      float dx;
      split_coord(xt * lscale, R, base, dx);
      // (a shared-memory atomicAdd per tap compiles to an ATOMS.CAST.SPIN loop on sm_100a: no gain)
      float cur[RD + 1];
      bool ok[RD + 1];
#pragma unroll
      for (int k = 0; k <= RD; ++k) {                 // all shared-memory loads first
        ok[k] = (unsigned)(base + k) < width;
        cur[k] = ok[k] ? rowl[base + k] : 0.f;
      }
#pragma unroll
      for (int k = 0; k <= RD; ++k)
        cur[k] += BwdTap<float>::eval(k > 0 ? gt[k - 1] : 0.f, k < RD ? gt[k] : 0.f, k > 0, k < RD, dx);
#pragma unroll
      for (int k = 0; k <= RD; ++k)
        if (ok[k]) rowl[base + k] = cur[k];
    };
    const int T = batch.T;
#pragma unroll
    for (int u = 0; u < kDepth - 1; ++u)
      if (u < T) load(u, g[u], x[u]);
    for (int t0 = 0; t0 < T; t0 += kDepth) {
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        const int t = t0 + u;
        if (t < T) {
          if (t + kDepth - 1 < T) load(t + kDepth - 1, g[(u + kDepth - 1) % kDepth], x[(u + kDepth - 1) % kDepth]);
          scatter(g[u], x[u]);
        }
      }
    }
  }
  __syncthreads();                         // all scatters done before the fold reads the rows
  const int w0 = geo.width[0];
  for (int j = tid; j < w0; j += 128) {
    int top = 0;
    for (int lv = 1; lv < NL; ++lv) {
      if ((j >> lv) >= geo.width[lv]) break;
      top = lv;
    }
    const int i1 = geo.off[1] + (j >> 1), i2 = geo.off[2] + (j >> 2), i3 = geo.off[3] + (j >> 3);
    float* out = g0_out + ((long long)b * HW + hw0) * g0_pitch + j;
    for (int pp = 0; pp < npx; ++pp) {
      const float* rowp = s_rows + pp * geo.rowlen;
      float tsum = 0.f;
      if (top >= 3) tsum = rowp[i3];
      if (top >= 2) tsum = rowp[i2] + 0.5f * tsum;
      if (top >= 1) tsum = rowp[i1] + 0.5f * tsum;
      out[(long long)pp * g0_pitch] = rowp[j] + 0.5f * tsum;
    }
  }
}

// Third version: same warp = level / lane = pixel mapping and register streaming, but the tile is
// stored [column][pixel] -- word (off_l + col) * 32 + pixel -- so the bank of every access is the
// lane: the scatter is conflict-free WHATEVER the coordinates are.  In the [pixel][column] layout
// above the bank is (pixel * rowlen + column) with data-dependent columns, i.e. random: ncu counted
// 23.2 M shared-memory wavefronts for 10.7 M requests at config 2 and the L1/shared pipe, 62 % busy
// with dependent read-modify-write chains, is what bounds the kernel (profiles/r01/README.md).
// The price is the write-out: with lane = pixel a warp cannot write a row of the level-0 gradient
// as one line, so every lane folds FOUR consecutive columns of its own pixel and stores them as one
// 16-byte piece (32 lanes -> 32 rows -> 32 half sectors per instruction; the other half of each
// sector follows from the same thread one step later, L2 merges them before the write-back).
// Same arithmetic and accumulation order per element as the other versions => bit-identical.
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// PF: software L2 prefetch of the gradient / coordinate lines PF iterations ahead (prefetch.global.L2: no
// register, no scoreboard).  cfg2 fp32: 141 -> 120 us with PF = 2 (4: 122, 8: 126); the product uses
// PF = 2 where this kernel is its path (fp16 gradients, planes that are not 16-byte aligned).
// PROBE (cross-check builds, tools/variants_probe.py): 1 = stream the gradients but do not scatter
// (what the global loads alone cost), 2 = scatter constants without loading anything (what the
// shared-memory work alone costs).  0 = the real kernel.
template <typename GT, int NL, int PROBE = 0, int PF = 0>
__global__ void __launch_bounds__(128)
lookup_bwd_cp_kernel(const __grid_constant__ BwdBatch batch, const BatchedGeom geo,
                     float* __restrict__ g0_out, long long g0_pitch, int HW, int accumulate) {
  constexpr int R = 4, RD = 9, NC = NL * RD, kDepth = SMOE_BWD_LVL_DEPTH;
  extern __shared__ __align__(16) uint8_t s_raw[];
  float* s_t = reinterpret_cast<float*>(s_raw);                       // [geo.rowlen columns][kTilePx]

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int l = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  const int npx = min(kTilePx, HW - hw0);
  const int w0 = geo.width[0];

  for (int i = tid; i < geo.rowlen * (kTilePx / 4); i += 128)
    reinterpret_cast<float4*>(s_t)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate) {
    __syncthreads();
    if (lane < npx) {
      const float* src = g0_out + ((long long)b * HW + hw0 + lane) * g0_pitch;
      for (int j = l; j < w0; j += 4) s_t[j * kTilePx + lane] = src[j];
    }
  }
  __syncthreads();

  if (l < NL && lane < npx) {
    float* coll = s_t + geo.off[l] * kTilePx + lane;
    const unsigned width = (unsigned)geo.width[l];
    const float lscale = 1.0f / (float)(1 << l);
    const long long gofs = ((long long)b * NC + l * RD) * HW + hw0 + lane;
    float g[kDepth][RD], x[kDepth];
    float probe_acc = 0.f;
    // PF > 0: ask L2 for the lines of iteration t + PF (fire and forget: no register, no scoreboard)
    auto prefetch = [&](int t) {
      if (PF > 0 && t < batch.T) {
        const GT* gp = static_cast<const GT*>(batch.gout[t]) + gofs;
#pragma unroll
        for (int k = 0; k < RD; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(gp + (long long)k * HW));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0 + lane));
      }
    };
    auto load = [&](int t, float (&gt)[RD], float& xt) {
      prefetch(t + PF);
      if (PROBE == 2) {
#pragma unroll
        for (int k = 0; k < RD; ++k) gt[k] = 0.25f * (float)(k + t);
        xt = (float)((hw0 + lane) % w0) - 0.37f * (float)t;
        return;
      }
      const GT* gp = static_cast<const GT*>(batch.gout[t]) + gofs;
#pragma unroll
      for (int k = 0; k < RD; ++k) gt[k] = to_float(__ldg(gp + (long long)k * HW));
      xt = __ldg(batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0 + lane);
    };
    auto scatter = [&](const float (&gt)[RD], float xt) {
      if (PROBE == 1) {
#pragma unroll
        for (int k = 0; k < RD; ++k) probe_acc += gt[k];
        probe_acc += xt;
        return;
      }
      int base;
      float dx;
      split_coord(xt * lscale, R, base, dx);
      float* win = coll + base * kTilePx;
      float cur[RD + 1];
      bool ok[RD + 1];
#pragma unroll
      for (int k = 0; k <= RD; ++k) {                 // all shared-memory loads first
        ok[k] = (unsigned)(base + k) < width;
        cur[k] = ok[k] ? win[k * kTilePx] : 0.f;
      }
#pragma unroll
      for (int k = 0; k <= RD; ++k)
        cur[k] += BwdTap<float>::eval(k > 0 ? gt[k - 1] : 0.f, k < RD ? gt[k] : 0.f, k > 0, k < RD, dx);
#pragma unroll
      for (int k = 0; k <= RD; ++k)
        if (ok[k]) win[k * kTilePx] = cur[k];
    };
    const int T = batch.T;
    if (PF > 0) {
      for (int t = kDepth - 1; t < kDepth - 1 + PF; ++t) prefetch(t);
    }
#pragma unroll
    for (int u = 0; u < kDepth - 1; ++u)
      if (u < T) load(u, g[u], x[u]);
    for (int t0 = 0; t0 < T; t0 += kDepth) {
#pragma unroll
      for (int u = 0; u < kDepth; ++u) {
        const int t = t0 + u;
        if (t < T) {
          if (t + kDepth - 1 < T) load(t + kDepth - 1, g[(u + kDepth - 1) % kDepth], x[(u + kDepth - 1) % kDepth]);
          scatter(g[u], x[u]);
        }
      }
    }
    if (PROBE == 1) coll[0] = probe_acc;
  }
  __syncthreads();                         // all scatters done before the fold reads the tile
  // fold + write-out: lane = pixel, a thread handles 4 consecutive level-0 columns at a time (they
  // share G2[j/4], G3[j/8] and use G1[j/2], G1[j/2+1]); same chain as fold_grad_vec_kernel
  if (lane < npx) {
    const int w1 = NL > 1 ? geo.width[1] : 0, w2 = NL > 2 ? geo.width[2] : 0, w3 = NL > 3 ? geo.width[3] : 0;
    const float* c0 = s_t + lane;
    const float* c1 = s_t + geo.off[1] * kTilePx + lane;
    const float* c2p = s_t + geo.off[2] * kTilePx + lane;
    const float* c3 = s_t + geo.off[3] * kTilePx + lane;
    float* orow = g0_out + ((long long)b * HW + hw0 + lane) * g0_pitch;
    const bool vec = (g0_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(g0_out) & 15) == 0;
    for (int j = 4 * l; j < w0; j += 16) {
      const int m = j >> 1, q = j >> 2, o = j >> 3;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (j + e < w0) ? c0[(j + e) * kTilePx] : 0.f;
      const float g1a = m < w1 ? c1[m * kTilePx] : 0.f;
      const float g1b = m + 1 < w1 ? c1[(m + 1) * kTilePx] : 0.f;
      const float g2 = q < w2 ? c2p[q * kTilePx] : 0.f;
      const float g3 = (q < w2 && o < w3) ? c3[o * kTilePx] : 0.f;
      const float t2 = q < w2 ? g2 + 0.5f * g3 : 0.f;
      const float a0 = m < w1 ? g1a + 0.5f * t2 : 0.f;
      const float a1 = m + 1 < w1 ? g1b + 0.5f * t2 : 0.f;
      v[0] += 0.5f * a0; v[1] += 0.5f * a0; v[2] += 0.5f * a1; v[3] += 0.5f * a1;
      if (vec) {                           // (rows are padded to the pitch: a tail piece may be stored whole)
        *reinterpret_cast<float4*>(orow + j) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e < w0) orow[j + e] = v[e];
      }
    }
  }
}

// Fourth version: the [column][pixel] tile of lookup_bwd_cp_kernel, with the gradient stream moved
// from a register ring to a THREAD-PRIVATE cp.async ring in shared memory.  ncu shows the register
// version stalled on the long scoreboard (5.7 of 11.5 cycles per issued instruction) although four
// iterations of loads are nominally in flight per warp, and its time does not change with the ring
// depth: a warp has six scoreboards, the ring's load groups alias on them, and waiting for the
// oldest group waits for the youngest too -- the software pipeline collapses.  cp.async groups are
// counted, not scoreboarded: `cp.async.wait_group kStages-2` waits for exactly the oldest stage.
// Each lane copies the 4-byte values of ITS pixel (9 channels of its level + the coordinate; the 32
// lanes of a warp form one 128-byte line per copy instruction) and later reads back only what it
// copied itself, so no barrier or warp synchronisation is involved at all.  fp32 gradients only
// (cp.async moves >= 4 bytes); same arithmetic and order as the other versions => bit-identical.
constexpr int kAsyncStages = 6;
template <int NL>
__global__ void __launch_bounds__(128)
lookup_bwd_async_kernel(const __grid_constant__ BwdBatch batch, const BatchedGeom geo,
                        float* __restrict__ g0_out, long long g0_pitch, int HW, int accumulate) {
  constexpr int R = 4, RD = 9, NC = NL * RD, S = kAsyncStages;
  extern __shared__ __align__(16) uint8_t s_raw[];
  float* s_t = reinterpret_cast<float*>(s_raw);                       // [geo.rowlen columns][kTilePx]
  float* s_ring = s_t + geo.rowlen * kTilePx;                         // [4 warps][S][RD + 1][kTilePx]

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int l = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  const int npx = min(kTilePx, HW - hw0);
  const int w0 = geo.width[0];
  const bool worker = l < NL && lane < npx;
  float* ring = s_ring + (l * S) * (RD + 1) * kTilePx + lane;
  const long long gofs = ((long long)b * NC + l * RD) * HW + hw0 + lane;

  auto issue = [&](int t) {
    if (worker && t < batch.T) {
      const float* gp = static_cast<const float*>(batch.gout[t]) + gofs;
      const uint32_t dst = ptx_smem(ring + (t % S) * (RD + 1) * kTilePx);
#pragma unroll
      for (int k = 0; k < RD; ++k) cp_async4(dst + k * kTilePx * 4, gp + (long long)k * HW);
      cp_async4(dst + RD * kTilePx * 4, batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0 + lane);
    }
    cp_async_commit_group();                          // exactly one group per iteration slot
  };
#pragma unroll
  for (int t = 0; t < S - 1; ++t) issue(t);           // the stream starts before the tile is cleared

  for (int i = tid; i < geo.rowlen * (kTilePx / 4); i += 128)
    reinterpret_cast<float4*>(s_t)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate) {
    __syncthreads();
    if (lane < npx) {
      const float* src = g0_out + ((long long)b * HW + hw0 + lane) * g0_pitch;
      for (int j = l; j < w0; j += 4) s_t[j * kTilePx + lane] = src[j];
    }
  }
  __syncthreads();

  {
    float* coll = s_t + geo.off[l < NL ? l : 0] * kTilePx + lane;
    const unsigned width = (unsigned)geo.width[l < NL ? l : 0];
    const float lscale = 1.0f / (float)(1 << l);
    const int T = batch.T;
    for (int t = 0; t < T; ++t) {
      cp_async_wait_group<S - 2>();                   // groups 0..t have landed (t+1 .. t+S-2 may be in flight)
      float gt[RD], xt = 0.f;
      const float* src = ring + (t % S) * (RD + 1) * kTilePx;
      if (worker) {
#pragma unroll
        for (int k = 0; k < RD; ++k) gt[k] = src[k * kTilePx];
        xt = src[RD * kTilePx];
      }
      issue(t + S - 1);                               // refills the slot consumed in the previous iteration
      if (worker) {
        int base;
        float dx;
        split_coord(xt * lscale, R, base, dx);
        float* win = coll + base * kTilePx;
        float cur[RD + 1];
        bool ok[RD + 1];
#pragma unroll
        for (int k = 0; k <= RD; ++k) {
          ok[k] = (unsigned)(base + k) < width;
          cur[k] = ok[k] ? win[k * kTilePx] : 0.f;
        }
#pragma unroll
        for (int k = 0; k <= RD; ++k)
          cur[k] += BwdTap<float>::eval(k > 0 ? gt[k - 1] : 0.f, k < RD ? gt[k] : 0.f, k > 0, k < RD, dx);
#pragma unroll
        for (int k = 0; k <= RD; ++k)
          if (ok[k]) win[k * kTilePx] = cur[k];
      }
    }
  }
  cp_async_wait_group<0>();
  __syncthreads();                         // all scatters done before the fold reads the tile
  if (lane < npx) {
    const int w1 = NL > 1 ? geo.width[1] : 0, w2 = NL > 2 ? geo.width[2] : 0, w3 = NL > 3 ? geo.width[3] : 0;
    const float* c0 = s_t + lane;
    const float* c1 = s_t + geo.off[1] * kTilePx + lane;
    const float* c2p = s_t + geo.off[2] * kTilePx + lane;
    const float* c3 = s_t + geo.off[3] * kTilePx + lane;
    float* orow = g0_out + ((long long)b * HW + hw0 + lane) * g0_pitch;
    const bool vec = (g0_pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(g0_out) & 15) == 0;
    for (int j = 4 * l; j < w0; j += 16) {
      const int m = j >> 1, q = j >> 2, o = j >> 3;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = (j + e < w0) ? c0[(j + e) * kTilePx] : 0.f;
      const float g1a = m < w1 ? c1[m * kTilePx] : 0.f;
      const float g1b = m + 1 < w1 ? c1[(m + 1) * kTilePx] : 0.f;
      const float g2 = q < w2 ? c2p[q * kTilePx] : 0.f;
      const float g3 = (q < w2 && o < w3) ? c3[o * kTilePx] : 0.f;
      const float t2 = q < w2 ? g2 + 0.5f * g3 : 0.f;
      const float a0 = m < w1 ? g1a + 0.5f * t2 : 0.f;
      const float a1 = m + 1 < w1 ? g1b + 0.5f * t2 : 0.f;
      v[0] += 0.5f * a0; v[1] += 0.5f * a0; v[2] += 0.5f * a1; v[3] += 0.5f * a1;
      if (vec) {
        *reinterpret_cast<float4*>(orow + j) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e < w0) orow[j + e] = v[e];
      }
    }
  }
}

// Fifth version (the product's default for fp32 gradients): the [column][pixel] tile, fed by a TMA
// producer warp, written out through a transposing staging tile.
// What the probes of lookup_bwd_cp_kernel showed at config 2 (tools/variants_probe.py): streaming the
// 383 MB of gradients with per-lane loads takes 130 us on its own -- warps cannot keep more than
// about one iteration of loads in flight, whether the ring lives in registers (scoreboards) or in
// cp.async groups (three CTAs per SM) -- and scatter + fold + write-out take 75 us on their own.
// So:  (1) warp 4 is a PRODUCER: per (tile, iteration) ONE tensor-map TMA load brings the [NC
// channels x 32 pixels] gradient box (3-D map {H*W1, NC, B} over grad_out[t], encoded per launch;
// tails zero-filled) and one 128-byte bulk copy the coordinates, completion counted in bytes on
// an mbarrier, into a kStages-deep ring; kStages-1 iterations (4.7 KB each) are in flight per CTA
// with no register, scoreboard or LSU-queue cost, and the stream starts before the tile is
// cleared.  (One bulk copy PER ROW instead of the tensor map -- 37 requests per stage -- made the
// kernel 2x slower than the register ring, 297 us: ~29 cycles of TMA issue per 128-byte request.)  (2) warps 0..3 are CONSUMERS
// (warp = level, lane = pixel): wait for the stage, read their 10 values (lane == bank), release the
// stage with one arrive per warp, scatter conflict-free into the [column][pixel] tile.  (3) the
// write-out folds 4 columns at a time with lane = pixel, transposes 32x32 blocks through a padded
// staging tile (carved out of the ring, idle by then) and stores whole 128-byte row segments.
// Same arithmetic and accumulation order per element as every other version => bit-identical.
// Measured at BASELINE config 2 (profiles/r02/bwd_variants*.log, ncu in profiles/r02): 108 us vs 155 us
// for lookup_bwd_lvl_kernel (DRAM floor 72 us).  ncu: L1/shared data pipe 65 % busy (per tile: ring
// fill 814 + ring reads 880 + window read-modify-write 1760 + clear 345 + fold 345 + stores ~570
// wavefronts), DRAM 48 %, issue 50 %: the shared-memory pipe is the soft limit, every access queues.
// Variants that moved the same wavefronts measured no better: persistent CTAs with the producer
// running across tile boundaries (111), guard columns instead of per-tap predicates + a software
// pipeline over the ring (113), a register queue of 2-6 stages on top of the ring (117-124), two
// warps per level (115-123), L2 prefetch 8-32 stages ahead (122-142), helper warps for the clear /
// fold / write phases (117-120), shallower rings with four CTAs per SM (136), L2 prefetch by the idle
// lanes of the producer warp 2 / 4 / 8 stages ahead (109 / 109 / 123; profiles/r02/bwd_variants_tma_l2_prefetch.log).
constexpr int kTmaStages = 6;
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}

struct BwdTmaMaps {
  CUtensorMap gout[kMaxBatchedIters];     // grad_out[t] as {H*W1, NC, B} fp32, box {32, NC, 1}
};

template <int NL>
__global__ void __launch_bounds__(160)
lookup_bwd_tma_kernel(const __grid_constant__ BwdBatch batch, const __grid_constant__ BwdTmaMaps maps,
                      const BatchedGeom geo, float* __restrict__ g0_out, long long g0_pitch, int HW,
                      int accumulate) {
  constexpr int R = 4, RD = 9, NC = NL * RD, ROWS = NC + 1, S = kTmaStages;
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ __align__(8) uint64_t s_full[S], s_empty[S];
  // TMA tensor loads want 128-byte aligned destinations: align the dynamic part by hand -- on the
  // 32-bit shared-window address, so that every pointer below is still provably a shared-memory
  // pointer (rounding the generic address up through uintptr_t made ptxas emit generic LD/ST with
  // 64-bit address arithmetic for every tile access: 200 instructions per iteration instead of 110)
  float* s_t = reinterpret_cast<float*>(s_raw + ((128u - (ptx_smem(s_raw) & 127u)) & 127u));   // [geo.rowlen columns][kTilePx]
  float* s_ring = s_t + geo.rowlen * kTilePx;                         // [S][ROWS][kTilePx]; later staging

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  const int npx = min(kTilePx, HW - hw0);
  const int w0 = geo.width[0];
  const int T = batch.T;
  const uint32_t row_bytes = (uint32_t)npx * 4u;                      // a multiple of 16 (HW % 4 == 0)

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      mbar_init_(ptx_smem(&s_full[i]), 1);
      mbar_init_(ptx_smem(&s_empty[i]), NL);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto produce = [&](int t, int slot) {               // one lane of the producer warp
    if (lane == 0) {
      const uint32_t bar = ptx_smem(&s_full[slot]);
      const uint32_t dst = ptx_smem(s_ring + slot * (ROWS * kTilePx));
      mbar_expect_tx_(bar, NC * kTilePx * 4 + row_bytes);          // the box counts in full, tails included
      ptx::tma_load_3d(dst, &maps.gout[t], bar, hw0, 0, b);
      bulk_copy_g2s(dst + NC * kTilePx * 4, batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0,
                    row_bytes, bar);
    }
  };
  if (warp == 4) {
    for (int t = 0; t < S && t < T; ++t) produce(t, t);   // the stream starts before the tile is cleared
  }
  for (int i = tid; i < geo.rowlen * (kTilePx / 4); i += 160)
    reinterpret_cast<float4*>(s_t)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate) {
    __syncthreads();
    if (lane < npx && warp < 4) {
      const float* src = g0_out + ((long long)b * HW + hw0 + lane) * g0_pitch;
      for (int j = warp; j < w0; j += 4) s_t[j * kTilePx + lane] = src[j];
    }
  }
  __syncthreads();

  if (warp == 4) {
    int slot = 0;
    uint32_t parity = 0;                               // parity of the slot's previous release
    for (int t = S; t < T; ++t) {
      mbar_wait_(ptx_smem(&s_empty[slot]), parity);
      produce(t, slot);
      if (++slot == S) { slot = 0; parity ^= 1u; }
    }
  } else if (warp < NL) {
    const int l = warp;
    float* coll = s_t + geo.off[l] * kTilePx + lane;
    const int width = geo.width[l];
    const float lscale = 1.0f / (float)(1 << l);
    const bool worker = lane < npx;
    int slot = 0;
    uint32_t parity = 0;
    for (int t = 0; t < T; ++t) {
      mbar_wait_(ptx_smem(&s_full[slot]), parity);
      const float* src = s_ring + slot * (ROWS * kTilePx) + l * (RD * kTilePx) + lane;
      float gt[RD];
#pragma unroll
      for (int k = 0; k < RD; ++k) gt[k] = src[k * kTilePx];           // (idle lanes of a tail tile read zero-filled columns)
      const float xt = s_ring[slot * (ROWS * kTilePx) + NC * kTilePx + (worker ? lane : 0)];
      __syncwarp();
      if (lane == 0) mbar_arrive_(ptx_smem(&s_empty[slot]));       // this warp is done with the stage
      if (++slot == S) { slot = 0; parity ^= 1u; }
      int base;
      float dx;
      split_coord(xt * lscale, R, base, dx);
      float* win = coll + base * kTilePx;
      // the whole window inside the row for every pixel of the warp (the common case): no per-tap predicates
      const bool inside = base >= 0 && base + RD < width;
      if (__all_sync(0xffffffffu, inside || !worker)) {
        if (worker) {
          float cur[RD + 1];
#pragma unroll
          for (int k = 0; k <= RD; ++k) cur[k] = win[k * kTilePx];
#pragma unroll
          for (int k = 0; k <= RD; ++k)
            cur[k] += BwdTap<float>::eval(k > 0 ? gt[k - 1] : 0.f, k < RD ? gt[k] : 0.f, k > 0, k < RD, dx);
#pragma unroll
          for (int k = 0; k <= RD; ++k) win[k * kTilePx] = cur[k];
        }
      } else if (worker) {
        float cur[RD + 1];
        bool ok[RD + 1];
#pragma unroll
        for (int k = 0; k <= RD; ++k) {
          ok[k] = (unsigned)(base + k) < (unsigned)width;
          cur[k] = ok[k] ? win[k * kTilePx] : 0.f;
        }
#pragma unroll
        for (int k = 0; k <= RD; ++k)
          cur[k] += BwdTap<float>::eval(k > 0 ? gt[k - 1] : 0.f, k < RD ? gt[k] : 0.f, k > 0, k < RD, dx);
#pragma unroll
        for (int k = 0; k <= RD; ++k)
          if (ok[k]) win[k * kTilePx] = cur[k];
      }
    }
  }
  __syncthreads();                         // all scatters done, every bulk copy has been consumed
  // fold (same chain as fold_grad_vec_kernel, lane = pixel, 4 columns at a time) + transposing write-out
  if (warp < 4) {
    const int w1 = NL > 1 ? geo.width[1] : 0, w2 = NL > 2 ? geo.width[2] : 0, w3 = NL > 3 ? geo.width[3] : 0;
    const float* c0 = s_t + lane;
    const float* c1 = s_t + geo.off[1] * kTilePx + lane;
    const float* c2p = s_t + geo.off[2] * kTilePx + lane;
    const float* c3 = s_t + geo.off[3] * kTilePx + lane;
    float* stg = s_ring + warp * (kTilePx * 33);                       // [32 pixels][33]
    float* orow0 = g0_out + ((long long)b * HW + hw0) * g0_pitch;
    for (int j0 = 32 * warp; j0 < w0; j0 += 128) {
#pragma unroll 2
      for (int c = 0; c < 32; c += 4) {
        const int j = j0 + c;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < w0) {
          const int m = j >> 1, q = j >> 2, o = j >> 3;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (j + e < w0) ? c0[(j + e) * kTilePx] : 0.f;
          const float g1a = m < w1 ? c1[m * kTilePx] : 0.f;
          const float g1b = m + 1 < w1 ? c1[(m + 1) * kTilePx] : 0.f;
          const float g2 = q < w2 ? c2p[q * kTilePx] : 0.f;
          const float g3 = (q < w2 && o < w3) ? c3[o * kTilePx] : 0.f;
          const float t2 = q < w2 ? g2 + 0.5f * g3 : 0.f;
          const float a0 = m < w1 ? g1a + 0.5f * t2 : 0.f;
          const float a1 = m + 1 < w1 ? g1b + 0.5f * t2 : 0.f;
          v[0] += 0.5f * a0; v[1] += 0.5f * a0; v[2] += 0.5f * a1; v[3] += 0.5f * a1;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) stg[lane * 33 + c + e] = v[e];     // bank (lane + c + e) % 32: conflict-free
      }
      __syncwarp();
      const int j = j0 + lane;
      if (j < w0) {
        float* dst = orow0 + j;
        for (int p = 0; p < npx; ++p, dst += g0_pitch)                  // one 128-byte row segment per store
          *dst = stg[p * 33 + lane];
      }
      __syncwarp();
    }
  }
}

// ----------------------------------------------------------------------------------- fold
// G0[j] += 0.5*(G1[j/2] + 0.5*(G2[j/4] + 0.5*G3[j/8])), stopping at the first level whose index
// falls outside its width (the column an odd level dropped, core/corr.py:124).
// Vector kernel: one thread owns 4 consecutive level-0 columns (they share G2[j/4], G3[j/8] and
// use G1[j/2], G1[j/2+1]) and processes kFoldUnroll such groups with all loads issued up front.
constexpr int kFoldUnroll = 4;

__global__ void __launch_bounds__(256)
fold_grad_vec_kernel(PyrDev gp, unsigned total_vecs, unsigned vecs_per_row) {
  pdl_launch_dependents();
  pdl_wait();
  const int w1 = gp.num_levels > 1 ? gp.width[1] : 0;
  const int w2 = gp.num_levels > 2 ? gp.width[2] : 0;
  const int w3 = gp.num_levels > 3 ? gp.width[3] : 0;
  const unsigned base = blockIdx.x * (256u * kFoldUnroll) + threadIdx.x;
  float4 g0[kFoldUnroll];
  float2 g1[kFoldUnroll];
  float g2[kFoldUnroll], g3[kFoldUnroll];
  float4* dst[kFoldUnroll];
  int jj[kFoldUnroll];
#pragma unroll
  for (int u = 0; u < kFoldUnroll; ++u) {
    const unsigned idx = base + u * 256u;
    dst[u] = nullptr;
    if (idx < total_vecs) {
      const unsigned row = idx / vecs_per_row;
      const int j = (int)(idx - row * vecs_per_row) * 4;
      jj[u] = j;
      dst[u] = reinterpret_cast<float4*>(static_cast<float*>(gp.ptr[0]) + (long long)row * gp.pitch[0] + j);
      g0[u] = __ldcg(dst[u]);
      const int m = j >> 1, q = j >> 2, o = j >> 3;
      const float* p1 = static_cast<const float*>(gp.ptr[1]) + (long long)row * gp.pitch[1] + m;
      g1[u].x = m < w1 ? __ldg(p1) : 0.f;
      g1[u].y = m + 1 < w1 ? __ldg(p1 + 1) : 0.f;
      g2[u] = q < w2 ? __ldg(static_cast<const float*>(gp.ptr[2]) + (long long)row * gp.pitch[2] + q) : 0.f;
      g3[u] = (q < w2 && o < w3) ? __ldg(static_cast<const float*>(gp.ptr[3]) + (long long)row * gp.pitch[3] + o) : 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < kFoldUnroll; ++u) {
    if (dst[u] == nullptr) continue;
    const int m = jj[u] >> 1, q = jj[u] >> 2;
    const float c2 = q < w2 ? g2[u] + 0.5f * g3[u] : 0.f;            // t = g2 + 0.5*g3
    const float a0 = m < w1 ? 0.5f * (g1[u].x + 0.5f * c2) : 0.f;    // 0.5*(g1 + 0.5*t)
    const float a1 = m + 1 < w1 ? 0.5f * (g1[u].y + 0.5f * c2) : 0.f;
    float4 v = g0[u];
    v.x += a0; v.y += a0; v.z += a1; v.w += a1;
    __stcg(dst[u], v);
  }
}

// Scalar fallback for unaligned level-0 widths / pitches: one warp per pixel row.
__global__ void __launch_bounds__(256)
fold_grad_scalar_kernel(PyrDev gp, long long rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int w0 = gp.width[0];
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp0; row < rows; row += nwarps) {
    float* g0 = static_cast<float*>(gp.ptr[0]) + row * gp.pitch[0];
    for (int j = lane; j < w0; j += 32) {
      int top = 0;
      for (int l = 1; l < gp.num_levels; ++l) {
        if ((j >> l) >= gp.width[l]) break;
        top = l;
      }
      float t = 0.f;
      for (int l = top; l >= 1; --l)
        t = __ldg(static_cast<const float*>(gp.ptr[l]) + row * gp.pitch[l] + (j >> l)) + 0.5f * t;
      g0[j] += 0.5f * t;
    }
  }
}

// ----------------------------------------------------------------------------------- fill levels
// level i[row, j] = (level i-1[row, 2j] + level i-1[row, 2j+1]) * 0.5 -- the build epilogue's own
// arithmetic (fp32 sum, one rounding to the storage type), for the levels a build with
// SMOE_BUILD_SKIP_ODD_LEVELS left out.  One thread per output element, coalesced along j.
template <typename VT>
__global__ void __launch_bounds__(256)
fill_level_kernel(const VT* __restrict__ prev, long long prev_pitch, VT* __restrict__ dst, long long dst_pitch,
                  int width, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / width;
  const int j = (int)(i - row * width);
  const VT* p = prev + row * prev_pitch + 2 * j;
  dst[row * dst_pitch + j] = from_float<VT>((to_float(p[0]) + to_float(p[1])) * 0.5f);
}

// --------------------------------------------------------------------------------- dispatch
bool pyramid_vec_ok(const smoe_pyramid_t* p) {   // also used by lookup_conv.cu
  const int es = dtype_size(p->dtype);
  for (int i = 0; i < p->num_levels; ++i) {
    if (!aligned16(p->level_ptr[i])) return false;
    if ((p->pitch[i] * es) % 16 != 0) return false;
    const int per16 = 16 / es;
    // a tail chunk may be read/written whole: it must stay inside the row
    if (((int64_t)p->width[i] + per16 - 1) / per16 * per16 > p->pitch[i]) return false;
  }
  return true;
}

#ifdef SMOE_CROSSCHECK
static thread_local int g_fwd_kernel_tls = -1;       // smoe_corr_debug_set_lookup_kernel
static int fwd_kernel_choice() { return g_fwd_kernel_tls >= 0 ? g_fwd_kernel_tls : tuning().fwd_kernel; }
constexpr bool kXcheck = true;
#else
static int fwd_kernel_choice() { return 0; }
constexpr bool kXcheck = false;
#endif
constexpr int kStreamPxHint = 5;         // volumes >> L2: level-0 fills and output stores evict_first
constexpr int kDefaultBwdKernel = 4;     // Tuning::bwd_kernel: 1 = lookup_bwd_lvl_kernel, 2 = lookup_bwd_cp_kernel, 3 = async,
                                         // 4 = lookup_bwd_tma_kernel (fp32 gradients, 16-byte aligned planes; else falls back to 2)

// The 256-bit path reads whole 32-byte chunks aligned in MEMORY: the first chunk of a level may
// begin 16 bytes before the level and the last may end 16 bytes after its row.  That stays inside
// the pixel's own allocation iff level 0 is 32-byte aligned with a 32-byte-multiple pitch, and
// level 2 either is too or lies inside level 0's row span with the same pitch (the fused layout of
// smoe_corr_pyramid_layout, where padding follows every level).
bool pyramid_ld256_ok(const PyrDev& d, int NL) {   // also used by lookup_conv.cu
  auto a32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  if (!a32(d.ptr[0]) || (d.pitch[0] * 4) % 32 != 0) return false;
  if (d.pitch[0] < (d.width[0] + 7) / 8 * 8) return false;
  if (NL <= 2) return true;
  if (a32(d.ptr[2]) && (d.pitch[2] * 4) % 32 == 0 && d.pitch[2] >= (d.width[2] + 7) / 8 * 8) return true;
  const char* p0 = static_cast<const char*>(d.ptr[0]);
  const char* p2 = static_cast<const char*>(d.ptr[2]);
  return d.pitch[2] == d.pitch[0] && p2 >= p0 + 16 &&
         p2 + ((long long)d.width[2] + 4) * 4 <= p0 + d.pitch[0] * 4;
}

template <typename VT, typename OT, int NL>
static void launch_fwd_px(const PyrDev& d, const float* coords, long long cbs, float cs, void* out,
                          int HW, int B, cudaStream_t st, const CoordUpd& upd) {
  constexpr int NP = (NL + 1) / 2;
  const Tuning& tn = tuning();
  const int which = fwd_kernel_choice();
  const bool big = volume_exceeds_l2(d, (long long)HW * B, (int)sizeof(VT));
  // measured (tools/px_sweep.sh, bench.py): 32 pixels per CTA and one thread per level pair
  // everywhere -- beyond L2 (cfg2 10.2 -> 8.9 us, cfg3 fp16 18.6 -> 12.4 us), for fp16 volumes
  // (cfg1 3.40 -> 3.16 us) and, inside a real step (cold coordinates, distinct output buffers),
  // for L2-resident fp32 volumes too: 3.38 -> 3.10 us per lookup of the cfg1 bench step
  const bool split = NP > 1 && (!kXcheck || tn.px_split != 0);
  const int tpb = (kXcheck && (tn.px_tpb == 64 || tn.px_tpb == 128)) ? tn.px_tpb : 32;
  dim3 grid((HW + tpb - 1) / tpb, B);
  OT* o = static_cast<OT*>(out);
  if constexpr (sizeof(VT) == 4 && sizeof(OT) == 4) {
    // 256-bit segment loads for fp32 volumes that stream from DRAM (tools/px256_check.py, 32 distinct
    // outputs per graph: cfg2 12.3 -> 8.4 us, cfg3 24.2 -> 20.5 us; L2-resident cfg1 3.08 -> 3.15 us,
    // so those keep the 16-byte loads)
    const bool want256 = which == 4 || (which == 0 && (tn.px_ld256 >= 0 ? tn.px_ld256 == 1 : big));
    if (want256 && pyramid_ld256_ok(d, NL)) {
      constexpr bool kSplit = NP > 1;
      dim3 block(tpb, kSplit ? NP : 1);
      if (!big) {
        launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, false, kSplit, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
        return;
      }
#define SMOE_PX256(H) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, true, kSplit, true, H>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd)
      // L2 eviction priorities (variants_probe.py, 32 distinct outputs per graph): for a volume several
      // times the L2 (cfg3, 600 MB) marking the level-0 fills AND the output stores evict_first gives
      // 20.5 -> 17.0-17.9 us; at 166 MB (cfg2), where most of level 0 still finds room in L2, any
      // hint on level 0 costs 5-10 %, and at 2.2 GB (cfg4) nothing moves outside the noise.
      const bool huge = (long long)HW * B * d.pitch[0] * (long long)sizeof(VT) > (256ll << 20);
      const int hint = (kXcheck && tn.px_hint >= 0) ? tn.px_hint : (huge ? kStreamPxHint : 0);
      if constexpr (kXcheck) {
        switch (hint) {
          case 1: SMOE_PX256(1); break;
          case 2: SMOE_PX256(2); break;
          case 3: SMOE_PX256(3); break;
          case 4: SMOE_PX256(4); break;
          case 5: SMOE_PX256(5); break;
          case 6: SMOE_PX256(6); break;
          case 7: SMOE_PX256(7); break;
          default: SMOE_PX256(0); break;
        }
      } else {
        if (hint == kStreamPxHint) SMOE_PX256(kStreamPxHint); else SMOE_PX256(0);
      }
#undef SMOE_PX256
      return;
    }
  }
  if (split) {
    constexpr int NLS = NP > 1 ? NL : 4;      // (never launched for NP == 1; avoids instantiating it)
    dim3 block(tpb, NP);
    if (big) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NLS, true, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
    else launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NLS, false, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
  } else if constexpr (NP == 1 || kXcheck) {
    dim3 block(tpb);
    if (big) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, true, false>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
    else launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, false, false>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
  }
}

template <typename VT, typename OT>
static void dispatch_fwd_vec(int NL, const PyrDev& d, const float* coords, long long cbs, float cs,
                             void* out, int HW, int B, cudaStream_t st, const CoordUpd& upd) {
#ifdef SMOE_CROSSCHECK
  // the lane-group kernels, selectable as bit-exact cross-checks (smoe_corr_debug_set_lookup_kernel)
  const int which = fwd_kernel_choice();
  if (which == 2 || which == 3) {
    if (which == 2 && (NL == 2 || NL == 4)) {
      if (NL == 4) launch_fwd_pair<VT, OT, 2>(d, coords, cbs, cs, out, HW, B, st, upd);
      else launch_fwd_pair<VT, OT, 1>(d, coords, cbs, cs, out, HW, B, st, upd);
      return;
    }
    switch (NL) {
      case 1: launch_fwd_vec<VT, OT, 1>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      case 2: launch_fwd_vec<VT, OT, 2>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      case 3: launch_fwd_vec<VT, OT, 3>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      default: launch_fwd_vec<VT, OT, 4>(d, coords, cbs, cs, out, HW, B, st, upd); break;
    }
    return;
  }
#endif
  switch (NL) {
    case 1: launch_fwd_px<VT, OT, 1>(d, coords, cbs, cs, out, HW, B, st, upd); break;
    case 2: launch_fwd_px<VT, OT, 2>(d, coords, cbs, cs, out, HW, B, st, upd); break;
    case 3: launch_fwd_px<VT, OT, 3>(d, coords, cbs, cs, out, HW, B, st, upd); break;
    default: launch_fwd_px<VT, OT, 4>(d, coords, cbs, cs, out, HW, B, st, upd); break;
  }
}

template <typename GT, int NL>
static void launch_bwd_acc_vec(const PyrDev& d, const float* coords, long long cbs, float cs,
                               const void* gout, int HW, int B, cudaStream_t st) {
  dim3 grid((HW + kTilePx - 1) / kTilePx, B);
  launch_kernel(lookup_bwd_acc_vec_kernel<GT, 4, NL>, grid, dim3(kFwdThreads), 0, st, d, coords, cbs, cs,
                static_cast<const GT*>(gout), HW);
}

template <typename GT>
static void dispatch_bwd_acc_vec(int NL, const PyrDev& d, const float* coords, long long cbs,
                                 float cs, const void* gout, int HW, int B, cudaStream_t st) {
  switch (NL) {
    case 1: launch_bwd_acc_vec<GT, 1>(d, coords, cbs, cs, gout, HW, B, st); break;
    case 2: launch_bwd_acc_vec<GT, 2>(d, coords, cbs, cs, gout, HW, B, st); break;
    case 3: launch_bwd_acc_vec<GT, 3>(d, coords, cbs, cs, gout, HW, B, st); break;
    default: launch_bwd_acc_vec<GT, 4>(d, coords, cbs, cs, gout, HW, B, st); break;
  }
}

template <typename VT, typename GT>
static int launch_bwd_dense(const smoe_pyramid_t* gp, const PyrDev& d, const float* coords,
                            long long cbs, float cs, int R, const void* gout, int HW, int B,
                            cudaStream_t st) {
  dim3 grid((HW + kTilePx - 1) / kTilePx, B);
  const size_t smem = ((size_t)gp->num_levels * (2 * R + 1) * (kTilePx + 1) + kTilePx) * sizeof(float);
  SMOE_REQUIRE(smem <= 48 * 1024, SMOE_ERR_UNSUPPORTED, "lookup_bwd: radius %d too large", R);
  constexpr int VEC = 16 / sizeof(VT);
  bool vec = pyramid_vec_ok(gp);
  for (int i = 0; i < gp->num_levels; ++i) vec = vec && (gp->width[i] % VEC == 0);
  if (vec)
    launch_kernel(lookup_bwd_dense_kernel<VT, GT, VEC>, grid, dim3(kDenseThreads), smem, st, d, coords,
                  cbs, cs, R, static_cast<const GT*>(gout), HW);
  else
    launch_kernel(lookup_bwd_dense_kernel<VT, GT, 1>, grid, dim3(kDenseThreads), smem, st, d, coords,
                  cbs, cs, R, static_cast<const GT*>(gout), HW);
  return SMOE_OK;
}

}  // namespace smoe

namespace smoe {
// ------------------------------------------------------- forward into a strided / multicast view
// smoe_corr_lookup_to: the lookup of THIS rank's shard written straight into its place inside a larger
// (gathered) output -- element (b, ch, hw) goes to out + b*obs + ch*ocs + hw -- and, with MM, through
// an NVSwitch MULTICAST address: one multimem.st per value and the switch replicates it into the
// corresponding buffer of every GPU of the group.  That is lookup + all-gather in one kernel (the
// counterpart of nn.DataParallel's gather of the outputs, train_stereo_sceneflow.py:133): no NCCL
// kernel, no second pass over the result, the transfer overlaps the gathers of the other warps.
// Same gather (PxPair / PxPair256) and arithmetic as lookup_fwd_px_kernel => bit-identical values.
__device__ __forceinline__ void st_multimem(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_multimem4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w) : "memory");
}
// VEC: the 18 channels a thread produced are transposed through a 2.3 KB shared-memory tile per warp so
// that a lane stores 16 bytes (four consecutive pixels of one channel) -- 4.5 store instructions per
// thread instead of 18, each a 16-byte NVLink payload per lane instead of 4 (needs H*W1 % 4 == 0 and
// 16-byte aligned strides).
template <typename VT, int NL, bool STREAM, bool LD256, bool MM, bool VEC>
__global__ void __launch_bounds__(64)
lookup_fwd_px_to_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                        float coord_scale, float* __restrict__ out, long long obs, long long ocs, int HW) {
  using Pair = typename std::conditional<LD256, PxPair256<STREAM>, PxPair<VT, STREAM>>::type;
  constexpr int R = 4, RD = 2 * R + 1, NP = (NL + 1) / 2;
  static_assert(NP == 2 && (!LD256 || sizeof(VT) == 4), "3- or 4-level pyramids; 256-bit loads for fp32 volumes");
  __shared__ __align__(16) float s_tile[VEC ? 2 : 1][VEC ? 2 * RD : 1][VEC ? 32 : 1];
  pdl_wait();
  const int b = blockIdx.y;
  const int lane = threadIdx.x;
  const int hw0 = blockIdx.x * 32;
  const int hw = hw0 + lane;
  const int pr = threadIdx.y;                              // one thread (warp) per level pair
  const bool live = hw < HW;
  if constexpr (!VEC) {
    if (!live) return;
  }
  const CoordUpd none = {nullptr, nullptr, nullptr, nullptr, 0};
  const bool first = pr == 0;
  const int lf = 2 * pr;
  const bool has_coarse = lf + 1 < NL;
  float ec[RD], ef[RD];
#pragma unroll
  for (int k = 0; k < RD; ++k) ec[k] = ef[k] = 0.f;
  if (live) {
    const float x0 = fetch_coord(coords, coords_bstride, none, b, hw, false) * coord_scale;
    const long long row = (long long)b * HW + hw;
    const VT* fine_row = static_cast<const VT*>(first ? pyr.ptr[0] : pyr.ptr[2]) + row * (first ? pyr.pitch[0] : pyr.pitch[2]);
    Pair pair;
    pair.load(reinterpret_cast<const typename std::conditional<LD256, float, VT>::type*>(fine_row),
              first ? pyr.width[0] : pyr.width[2], x0 * (first ? 1.0f : 0.25f), 0);
    if (has_coarse) pair.template coarse<float>(first ? pyr.width[1] : pyr.width[3], ec);
    pair.template fine<float>(ef);
  }
  pdl_launch_dependents();
  float* dst = out + (long long)b * obs;
  if constexpr (!VEC) {
    auto put = [&](float* p, float v) {
      if (MM) st_multimem(p, v);
      else *p = v;
    };
    if (has_coarse) {
      float* d1 = dst + (long long)(lf + 1) * RD * ocs + hw;
#pragma unroll
      for (int k = 0; k < RD; ++k) put(d1 + (long long)k * ocs, ec[k]);
    }
    float* d0 = dst + (long long)lf * RD * ocs + hw;
#pragma unroll
    for (int k = 0; k < RD; ++k) put(d0 + (long long)k * ocs, ef[k]);
  } else {
    // tile rows: [0, RD) = level lf (fine), [RD, 2*RD) = level lf + 1 (coarse): consecutive output channels
    float (*t)[32] = s_tile[pr];
#pragma unroll
    for (int k = 0; k < RD; ++k) {
      t[k][lane] = ef[k];
      t[RD + k][lane] = ec[k];
    }
    __syncwarp();
    const int nch = has_coarse ? 2 * RD : RD;
    const int q = lane & 7, cr = lane >> 3;               // pixel quad, channel within a group of 4
    const int px = hw0 + 4 * q;
    if (px < HW) {                                        // (HW % 4 == 0: a quad is entirely inside or outside)
#pragma unroll
      for (int c0 = 0; c0 < 2 * RD; c0 += 4) {
        const int c = c0 + cr;
        if (c < nch) {
          const float4 v = *reinterpret_cast<const float4*>(&t[c][4 * q]);
          float* p = dst + (long long)(lf * RD + c) * ocs + px;
          if (MM) st_multimem4(p, v);
          else *reinterpret_cast<float4*>(p) = v;
        }
      }
    }
  }
}

static int lookup_impl(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                       float coord_scale, int radius, void* out, int out_dtype, smoe_stream_t stream,
                       const CoordUpd& upd) {
  if (int rc = validate_pyramid(pyr, "lookup")) return rc;
  SMOE_REQUIRE(radius >= 0 && radius <= 64, SMOE_ERR_INVALID_ARG, "lookup: radius %d outside [0,64]",
               radius);
  SMOE_REQUIRE(dtype_ok(out_dtype), SMOE_ERR_INVALID_ARG, "lookup: unknown out dtype %d", out_dtype);
  SMOE_REQUIRE(!(pyr->dtype == SMOE_DTYPE_F32 && out_dtype == SMOE_DTYPE_F16), SMOE_ERR_UNSUPPORTED,
               "lookup: fp16 output from an fp32 pyramid is not part of the reference API");
  const long long HW = (long long)pyr->H * pyr->W1;
  if (pyr->B == 0 || HW == 0) return SMOE_OK;
  SMOE_REQUIRE(coords && out, SMOE_ERR_INVALID_ARG, "lookup: NULL coords/out pointer");
  SMOE_REQUIRE(coords_batch_stride >= HW, SMOE_ERR_INVALID_ARG,
               "lookup: coords batch stride %lld < H*W1 %lld", (long long)coords_batch_stride, HW);
  SMOE_REQUIRE(pyr->B <= 65535, SMOE_ERR_UNSUPPORTED, "lookup: batch %d > 65535", pyr->B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PyrDev d = to_dev(pyr);
  const bool f16v = pyr->dtype == SMOE_DTYPE_F16, f16o = out_dtype == SMOE_DTYPE_F16;
  if (pyr->levels_missing & ((1 << pyr->num_levels) - 1)) {
    // levels skipped by the build (SMOE_BUILD_SKIP_ODD_LEVELS): only the thread-per-pixel kernel,
    // which reads levels 0 and 2 and recomputes 1 and 3, can serve the lookup
    SMOE_REQUIRE((pyr->levels_missing & 0x5) == 0, SMOE_ERR_INVALID_ARG,
                 "lookup: an even level is marked missing (levels_missing=0x%x)", pyr->levels_missing);
    SMOE_REQUIRE(radius == 4 && pyramid_vec_ok(pyr) &&
                     (fwd_kernel_choice() == 0 || fwd_kernel_choice() == 1 || fwd_kernel_choice() == 4),
                 SMOE_ERR_INVALID_ARG,
                 "lookup: levels 0x%x of the pyramid were not built; radius %d / this kernel choice "
                 "needs them (build without SMOE_BUILD_SKIP_ODD_LEVELS)", pyr->levels_missing, radius);
  }
  if (radius == 4 && pyramid_vec_ok(pyr)) {
    if (!f16v)
      dispatch_fwd_vec<float, float>(pyr->num_levels, d, coords, coords_batch_stride, coord_scale,
                                     out, (int)HW, pyr->B, st, upd);
    else if (f16o)
      dispatch_fwd_vec<__half, __half>(pyr->num_levels, d, coords, coords_batch_stride, coord_scale,
                                       out, (int)HW, pyr->B, st, upd);
    else
      dispatch_fwd_vec<__half, float>(pyr->num_levels, d, coords, coords_batch_stride, coord_scale,
                                      out, (int)HW, pyr->B, st, upd);
  } else {
    dim3 grid((unsigned)((HW + 127) / 128), pyr->B);
    if (!f16v)
      launch_kernel(lookup_fwd_generic_kernel<float, float>, grid, dim3(128), 0, st, d, coords,
                    (long long)coords_batch_stride, coord_scale, radius, static_cast<float*>(out), (int)HW, upd);
    else if (f16o)
      launch_kernel(lookup_fwd_generic_kernel<__half, __half>, grid, dim3(128), 0, st, d, coords,
                    (long long)coords_batch_stride, coord_scale, radius, static_cast<__half*>(out), (int)HW, upd);
    else
      launch_kernel(lookup_fwd_generic_kernel<__half, float>, grid, dim3(128), 0, st, d, coords,
                    (long long)coords_batch_stride, coord_scale, radius, static_cast<float*>(out), (int)HW, upd);
  }
  return check_launch("lookup forward");
}
}  // namespace smoe

extern "C" {

#ifdef SMOE_CROSSCHECK
void smoe_corr_debug_set_lookup_kernel(int which) { smoe::g_fwd_kernel_tls = which; }
#endif

int smoe_corr_lookup(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                     float coord_scale, int radius, void* out, int out_dtype,
                     smoe_stream_t stream) {
  const smoe::CoordUpd none = {nullptr, nullptr, nullptr, nullptr, 0};
  return smoe::lookup_impl(pyr, coords, coords_batch_stride, coord_scale, radius, out, out_dtype, stream, none);
}

int smoe_corr_lookup_to(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                        float coord_scale, int radius, void* out, int64_t out_batch_stride,
                        int64_t out_channel_stride, int out_dtype, unsigned out_flags, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(pyr, "lookup_to")) return rc;
  const long long HW = (long long)pyr->H * pyr->W1;
  if (pyr->B == 0 || HW == 0) return SMOE_OK;
  SMOE_REQUIRE(coords && out, SMOE_ERR_INVALID_ARG, "lookup_to: NULL coords/out pointer");
  SMOE_REQUIRE(coords_batch_stride >= HW, SMOE_ERR_INVALID_ARG,
               "lookup_to: coords batch stride %lld < H*W1 %lld", (long long)coords_batch_stride, HW);
  SMOE_REQUIRE((out_flags & ~(SMOE_OUT_MULTIMEM | SMOE_OUT_SCALAR)) == 0, SMOE_ERR_INVALID_ARG,
               "lookup_to: unknown flags 0x%x", out_flags);
  // the fused gather covers what a sharded CorrBlock1D produces: radius 4, 3 or 4 levels, fp32 outputs,
  // vector-aligned pyramid; everything else -> SMOE_ERR_UNSUPPORTED (use smoe_corr_lookup + a collective)
  SMOE_REQUIRE(radius == 4 && (pyr->num_levels == 3 || pyr->num_levels == 4) && out_dtype == SMOE_DTYPE_F32 &&
                   pyramid_vec_ok(pyr) && (pyr->levels_missing & 0x5) == 0 && pyr->B <= 65535,
               SMOE_ERR_UNSUPPORTED, "lookup_to: radius %d / %d levels / out dtype %d not covered", radius,
               pyr->num_levels, out_dtype);
  SMOE_REQUIRE(out_channel_stride >= HW && out_batch_stride >= out_channel_stride * pyr->num_levels * 9,
               SMOE_ERR_INVALID_ARG, "lookup_to: output strides (%lld, %lld) overlap", (long long)out_batch_stride,
               (long long)out_channel_stride);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PyrDev d = to_dev(pyr);
  const bool mm = (out_flags & SMOE_OUT_MULTIMEM) != 0;
  const bool f16v = pyr->dtype == SMOE_DTYPE_F16;
  const bool big = volume_exceeds_l2(d, HW * pyr->B, f16v ? 2 : 4);
  const bool ld256 = !f16v && big && pyramid_ld256_ok(d, pyr->num_levels);
  dim3 grid((unsigned)((HW + 31) / 32), pyr->B), block(32, 2);
  float* o = static_cast<float*>(out);
  const long long obs = out_batch_stride, ocs = out_channel_stride;
  // 16-byte stores (four pixels of a channel per lane) when every address they touch is 16-byte aligned
  const bool vec = HW % 4 == 0 && obs % 4 == 0 && ocs % 4 == 0 && aligned16(out) && (out_flags & SMOE_OUT_SCALAR) == 0;
#define SMOE_TO_K(VT, NLV, STR, L256, MMV, VECV)                                                             \
  launch_kernel_pdl(lookup_fwd_px_to_kernel<VT, NLV, STR, L256, MMV, VECV>, grid, block, 0, st, d, coords,   \
                    (long long)coords_batch_stride, coord_scale, o, obs, ocs, (int)HW)
#define SMOE_TO(VT, NLV, STR, L256)                                                                          \
  do {                                                                                                       \
    if (mm && vec) SMOE_TO_K(VT, NLV, STR, L256, true, true);                                                \
    else if (mm) SMOE_TO_K(VT, NLV, STR, L256, true, false);                                                 \
    else if (vec) SMOE_TO_K(VT, NLV, STR, L256, false, true);                                                \
    else SMOE_TO_K(VT, NLV, STR, L256, false, false);                                                        \
  } while (0)
#define SMOE_TO_NL(VT, STR, L256)                                                                            \
  do {                                                                                                       \
    if (pyr->num_levels == 4) SMOE_TO(VT, 4, STR, L256); else SMOE_TO(VT, 3, STR, L256);                     \
  } while (0)
  if (f16v) {
    if (big) SMOE_TO_NL(__half, true, false); else SMOE_TO_NL(__half, false, false);
  } else if (ld256) {
    SMOE_TO_NL(float, true, true);
  } else if (big) {
    SMOE_TO_NL(float, true, false);
  } else {
    SMOE_TO_NL(float, false, false);
  }
#undef SMOE_TO_NL
#undef SMOE_TO
#undef SMOE_TO_K
  return check_launch("lookup forward (strided / multicast output)");
}

int smoe_corr_lookup_update(const smoe_pyramid_t* pyr, const float* coords1, const float* delta_flow,
                            const float* coords0, float* coords1_out, float* flow_out, int radius,
                            void* out, int out_dtype, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(pyr, "lookup_update")) return rc;
  const long long plane = (long long)pyr->H * pyr->W1;
  if (pyr->B == 0 || plane == 0) return SMOE_OK;
  SMOE_REQUIRE(coords1 && coords1_out, SMOE_ERR_INVALID_ARG, "lookup_update: NULL coords1 / coords1_out");
  SMOE_REQUIRE(coords1_out != coords1, SMOE_ERR_INVALID_ARG,
               "lookup_update: coords1_out must not alias coords1 (several lanes read each coordinate)");
  SMOE_REQUIRE(flow_out == nullptr || coords0 != nullptr, SMOE_ERR_INVALID_ARG,
               "lookup_update: flow_out needs coords0");
  const CoordUpd upd = {delta_flow, coords0, coords1_out, flow_out, plane};
  return lookup_impl(pyr, coords1, 2 * plane, 1.0f, radius, out, out_dtype, stream, upd);
}

int smoe_corr_lookup_bwd(const smoe_pyramid_t* gp, const float* coords, int64_t coords_batch_stride,
                         float coord_scale, int radius, const void* grad_out, int grad_out_dtype,
                         int mode, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(gp, "lookup_bwd")) return rc;
  SMOE_REQUIRE(radius >= 0 && radius <= 64, SMOE_ERR_INVALID_ARG,
               "lookup_bwd: radius %d outside [0,64]", radius);
  SMOE_REQUIRE(dtype_ok(grad_out_dtype), SMOE_ERR_INVALID_ARG, "lookup_bwd: unknown grad dtype %d",
               grad_out_dtype);
  SMOE_REQUIRE(mode == SMOE_BWD_OVERWRITE || mode == SMOE_BWD_ACCUMULATE, SMOE_ERR_INVALID_ARG,
               "lookup_bwd: unknown mode %d", mode);
  const long long HW = (long long)gp->H * gp->W1;
  if (gp->B == 0 || HW == 0) return SMOE_OK;
  SMOE_REQUIRE(coords && grad_out, SMOE_ERR_INVALID_ARG, "lookup_bwd: NULL coords/grad pointer");
  SMOE_REQUIRE(coords_batch_stride >= HW, SMOE_ERR_INVALID_ARG,
               "lookup_bwd: coords batch stride %lld < H*W1 %lld", (long long)coords_batch_stride, HW);
  SMOE_REQUIRE(gp->B <= 65535, SMOE_ERR_UNSUPPORTED, "lookup_bwd: batch %d > 65535", gp->B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const PyrDev d = to_dev(gp);
  const bool f16g = grad_out_dtype == SMOE_DTYPE_F16;
  if (mode == SMOE_BWD_ACCUMULATE) {
    SMOE_REQUIRE(gp->dtype == SMOE_DTYPE_F32, SMOE_ERR_UNSUPPORTED,
                 "lookup_bwd: the accumulating gradient pyramid must be fp32");
    if (radius == 4 && pyramid_vec_ok(gp)) {
      if (f16g)
        dispatch_bwd_acc_vec<__half>(gp->num_levels, d, coords, coords_batch_stride, coord_scale,
                                     grad_out, (int)HW, gp->B, st);
      else
        dispatch_bwd_acc_vec<float>(gp->num_levels, d, coords, coords_batch_stride, coord_scale,
                                    grad_out, (int)HW, gp->B, st);
    } else {
      dim3 grid((unsigned)((HW + 127) / 128), gp->B);
      if (f16g)
        launch_kernel(lookup_bwd_acc_generic_kernel<__half>, grid, dim3(128), 0, st, d, coords,
                      (long long)coords_batch_stride, coord_scale, radius,
                      static_cast<const __half*>(grad_out), (int)HW);
      else
        launch_kernel(lookup_bwd_acc_generic_kernel<float>, grid, dim3(128), 0, st, d, coords,
                      (long long)coords_batch_stride, coord_scale, radius,
                      static_cast<const float*>(grad_out), (int)HW);
    }
  } else {
    int rc;
    if (gp->dtype == SMOE_DTYPE_F16) {
      SMOE_REQUIRE(f16g, SMOE_ERR_UNSUPPORTED,
                   "lookup_bwd: an fp16 gradient volume needs an fp16 upstream gradient");
      rc = launch_bwd_dense<__half, __half>(gp, d, coords, coords_batch_stride, coord_scale, radius,
                                            grad_out, (int)HW, gp->B, st);
    } else if (f16g) {
      rc = launch_bwd_dense<float, __half>(gp, d, coords, coords_batch_stride, coord_scale, radius,
                                           grad_out, (int)HW, gp->B, st);
    } else {
      rc = launch_bwd_dense<float, float>(gp, d, coords, coords_batch_stride, coord_scale, radius,
                                          grad_out, (int)HW, gp->B, st);
    }
    if (rc) return rc;
  }
  return check_launch("lookup backward");
}

int smoe_corr_lookup_bwd_batched(const smoe_pyramid_t* g0, int num_levels, int num_iters,
                                 const float* const* coords, const int64_t* coords_batch_stride,
                                 const void* const* grad_out, int grad_out_dtype, int radius,
                                 int accumulate, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(g0, "lookup_bwd_batched")) return rc;
  SMOE_REQUIRE(g0->dtype == SMOE_DTYPE_F32, SMOE_ERR_UNSUPPORTED, "lookup_bwd_batched: the gradient must be fp32");
  SMOE_REQUIRE(num_levels >= 1 && num_levels <= SMOE_MAX_LEVELS, SMOE_ERR_INVALID_ARG,
               "lookup_bwd_batched: num_levels=%d outside [1,%d]", num_levels, SMOE_MAX_LEVELS);
  SMOE_REQUIRE(radius == 4, SMOE_ERR_UNSUPPORTED, "lookup_bwd_batched: radius %d (only 4 is batched)", radius);
  SMOE_REQUIRE(num_iters >= 1 && num_iters <= kMaxBatchedIters, SMOE_ERR_UNSUPPORTED,
               "lookup_bwd_batched: %d iterations outside [1,%d]", num_iters, kMaxBatchedIters);
  SMOE_REQUIRE(dtype_ok(grad_out_dtype), SMOE_ERR_INVALID_ARG, "lookup_bwd_batched: unknown grad dtype %d",
               grad_out_dtype);
  const long long HW = (long long)g0->H * g0->W1;
  if (g0->B == 0 || HW == 0 || g0->width[0] == 0) return SMOE_OK;
  SMOE_REQUIRE(coords && coords_batch_stride && grad_out, SMOE_ERR_INVALID_ARG, "lookup_bwd_batched: NULL array");
  SMOE_REQUIRE(g0->B <= 65535, SMOE_ERR_UNSUPPORTED, "lookup_bwd_batched: batch %d > 65535", g0->B);
  const int ges = dtype_size(grad_out_dtype);
#ifdef SMOE_CROSSCHECK
  const bool staged = tuning().bwd_staged != 0;    // first version (cp.async-staged tiles, 16-byte rules)
#else
  constexpr bool staged = false;
#endif
  SMOE_REQUIRE(!staged || (HW * ges) % 16 == 0, SMOE_ERR_UNSUPPORTED,
               "lookup_bwd_batched: H*W1=%lld planes are not 16-byte multiples (cp.async staging)", HW);
  BwdBatch batch;
  batch.T = num_iters;
  for (int t = 0; t < num_iters; ++t) {
    SMOE_REQUIRE(coords[t] && grad_out[t], SMOE_ERR_INVALID_ARG, "lookup_bwd_batched: NULL pointer at iteration %d", t);
    SMOE_REQUIRE(coords_batch_stride[t] >= HW, SMOE_ERR_INVALID_ARG,
                 "lookup_bwd_batched: iteration %d coords batch stride < H*W1", t);
    SMOE_REQUIRE(!staged || (aligned16(coords[t]) && aligned16(grad_out[t]) && coords_batch_stride[t] % 4 == 0),
                 SMOE_ERR_UNSUPPORTED, "lookup_bwd_batched: iteration %d operands are not 16-byte aligned", t);
    batch.coords[t] = coords[t];
    batch.gout[t] = grad_out[t];
    batch.coords_bstride[t] = coords_batch_stride[t];
  }
  BatchedGeom geo;
  geo.num_levels = num_levels;
  int off = 0, w = g0->width[0];
  for (int l = 0; l < SMOE_MAX_LEVELS; ++l) {
    geo.width[l] = l < num_levels ? w : 0;
    geo.off[l] = off;
    if (l < num_levels) off += (w + 3) / 4 * 4;
    w /= 2;
  }
  // lane = pixel, so consecutive lanes hit rows `rowlen` apart at columns that advance by 2^-level
  // per pixel; rowlen % 32 == 2 makes level 0 conflict-free for smooth disparities (3 banks per
  // pixel) and keeps the coarser levels at <= 2-way.  (rowlen % 32 == 0, which W2=184 happened to
  // give with a fixed pad, made every scatter an 8-way bank conflict.)
  const int rowmod = kXcheck ? (tuning().bwd_rowmod & 31) : 2;
  geo.rowlen = off + ((rowmod - off % 32) + 32) % 32;
  size_t smem = (size_t)kTilePx * geo.rowlen * 4;
#ifdef SMOE_CROSSCHECK
  if (staged) {   // the staged kernel wants pixel rows 4 banks apart (the 8 pixels of a warp cover all banks)
    geo.rowlen = off + ((4 - off % 32) + 32) % 32;
    smem = (size_t)kTilePx * geo.rowlen * 4 +
           kBatchedStages * ((size_t)num_levels * 9 * (kTilePx + 32 / ges) * ges + kTilePx * 4);
  }
#endif
  // (the Python side bounds this with off + 32 words per row)
  SMOE_REQUIRE(smem <= 227 * 1024, SMOE_ERR_UNSUPPORTED,
               "lookup_bwd_batched: W2=%d needs %zu bytes of shared memory per CTA", g0->width[0], smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((HW + kTilePx - 1) / kTilePx), g0->B);
  float* out = static_cast<float*>(g0->level_ptr[0]);
  const int which = kXcheck ? tuning().bwd_kernel : kDefaultBwdKernel;   // 1 = [pixel][column], 2 = [column][pixel]
  bool async = which == 3 && !staged && grad_out_dtype == SMOE_DTYPE_F32;
  bool tma = which == 4 && !staged && grad_out_dtype == SMOE_DTYPE_F32 && HW % 4 == 0;
  for (int t = 0; tma && t < num_iters; ++t)       // bulk copies: 16-byte aligned rows
    tma = aligned16(coords[t]) && aligned16(grad_out[t]) && coords_batch_stride[t] % 4 == 0;
  if (which >= 2 && !staged) {
    geo.rowlen = off;                       // columns of the [column][pixel] tile
    smem = (size_t)kTilePx * off * 4;
    const size_t ring = (size_t)4 * kAsyncStages * 10 * kTilePx * 4;
    if (async && smem + ring <= 227 * 1024) smem += ring; else async = false;
    size_t tring = (size_t)kTmaStages * (num_levels * 9 + 1) * kTilePx * 4;
    if (tring < (size_t)4 * kTilePx * 33 * 4) tring = (size_t)4 * kTilePx * 33 * 4;   // doubles as the staging tiles
    if (tma && smem + tring + 256 <= 227 * 1024) smem += tring + 128; else tma = false;
  }
  BwdTmaMaps maps;
  if (tma) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(tensor_map_encoder_ptr());
    SMOE_REQUIRE(enc != nullptr, SMOE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const int nc = num_levels * 9;
    const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)nc, (cuuint64_t)g0->B};
    const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)nc * HW * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kTilePx, (cuuint32_t)nc, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    for (int t = 0; t < num_iters; ++t) {
      const CUresult r = enc(&maps.gout[t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(grad_out[t]), gdim,
                             gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      SMOE_REQUIRE(r == CUDA_SUCCESS, SMOE_ERR_CUDA, "lookup_bwd_batched: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
  }
#define SMOE_LAUNCH_TMA(NLV)                                                                             \
  do {                                                                                                   \
    SMOE_CUDA_OK(cudaFuncSetAttribute(lookup_bwd_tma_kernel<NLV>,                                        \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    SMOE_CUDA_OK(launch_kernel(lookup_bwd_tma_kernel<NLV>, grid, dim3(160), smem, st, batch, maps, geo,  \
                               out, (long long)g0->pitch[0], (int)HW, accumulate));                      \
  } while (0)
#define SMOE_LAUNCH_ASYNC(NLV)                                                                           \
  do {                                                                                                   \
    SMOE_CUDA_OK(cudaFuncSetAttribute(lookup_bwd_async_kernel<NLV>,                                      \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    SMOE_CUDA_OK(launch_kernel(lookup_bwd_async_kernel<NLV>, grid, dim3(128), smem, st, batch, geo, out, \
                               (long long)g0->pitch[0], (int)HW, accumulate));                           \
  } while (0)
#define SMOE_LAUNCH_FN(FN, THREADS)                                                                      \
  do {                                                                                                   \
    SMOE_CUDA_OK(cudaFuncSetAttribute(FN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    SMOE_CUDA_OK(launch_kernel(FN, grid, dim3(THREADS), smem, st, batch, geo, out,                       \
                               (long long)g0->pitch[0], (int)HW, accumulate));                           \
  } while (0)
#define SMOE_LAUNCH_KERN(KERN, GT, NLV, THREADS) SMOE_LAUNCH_FN((KERN<GT, NLV>), THREADS)
#ifdef SMOE_CROSSCHECK
#define SMOE_LAUNCH_BATCHED(GT, NLV)                                                                     \
  do {                                                                                                   \
    if (staged) SMOE_LAUNCH_KERN(lookup_bwd_batched_kernel, GT, NLV, kBatchedThreads);                   \
    else if (async) SMOE_LAUNCH_ASYNC(NLV);                                                              \
    else if (tma) SMOE_LAUNCH_TMA(NLV);                                                                  \
    else if (which == 11) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 1>), 128);                       \
    else if (which == 12) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 2>), 128);                       \
    else if (which == 21) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 0, 2>), 128);                    \
    else if (which == 22) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 0, 4>), 128);                    \
    else if (which == 23) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 0, 8>), 128);                    \
    else if (which == 24) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 1, 4>), 128);                    \
    else if (which >= 2) SMOE_LAUNCH_KERN(lookup_bwd_cp_kernel, GT, NLV, 128);                           \
    else SMOE_LAUNCH_KERN(lookup_bwd_lvl_kernel, GT, NLV, 128);                                          \
  } while (0)
#else
#define SMOE_LAUNCH_BATCHED(GT, NLV)                                                                     \
  do {                                                                                                   \
    if (async) SMOE_LAUNCH_ASYNC(NLV);                                                                   \
    else if (tma) SMOE_LAUNCH_TMA(NLV);                                                                  \
    else if (kDefaultBwdKernel >= 2) SMOE_LAUNCH_FN((lookup_bwd_cp_kernel<GT, NLV, 0, 2>), 128);         \
    else SMOE_LAUNCH_KERN(lookup_bwd_lvl_kernel, GT, NLV, 128);                                          \
  } while (0)
#endif
  const bool f16g = grad_out_dtype == SMOE_DTYPE_F16;
  switch (num_levels) {
    case 1: if (f16g) SMOE_LAUNCH_BATCHED(__half, 1); else SMOE_LAUNCH_BATCHED(float, 1); break;
    case 2: if (f16g) SMOE_LAUNCH_BATCHED(__half, 2); else SMOE_LAUNCH_BATCHED(float, 2); break;
    case 3: if (f16g) SMOE_LAUNCH_BATCHED(__half, 3); else SMOE_LAUNCH_BATCHED(float, 3); break;
    default: if (f16g) SMOE_LAUNCH_BATCHED(__half, 4); else SMOE_LAUNCH_BATCHED(float, 4); break;
  }
#undef SMOE_LAUNCH_BATCHED
#undef SMOE_LAUNCH_KERN
#undef SMOE_LAUNCH_FN
#undef SMOE_LAUNCH_ASYNC
#undef SMOE_LAUNCH_TMA
  return check_launch("lookup backward (batched)");
}

int smoe_corr_fill_levels(const smoe_pyramid_t* pyr, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(pyr, "fill_levels")) return rc;
  const int missing = pyr->levels_missing & ((1 << pyr->num_levels) - 1);
  SMOE_REQUIRE((missing & 1) == 0, SMOE_ERR_INVALID_ARG, "fill_levels: level 0 cannot be recomputed");
  const long long rows = (long long)pyr->B * pyr->H * pyr->W1;
  if (rows == 0 || missing == 0) return SMOE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int i = 1; i < pyr->num_levels; ++i) {
    if (!((missing >> i) & 1) || pyr->width[i] == 0) continue;
    SMOE_REQUIRE(pyr->width[i] <= pyr->width[i - 1] / 2, SMOE_ERR_INVALID_ARG,
                 "fill_levels: level %d width %d > floor(%d/2)", i, pyr->width[i], pyr->width[i - 1]);
    const long long total = rows * pyr->width[i];
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (pyr->dtype == SMOE_DTYPE_F16)
      launch_kernel(fill_level_kernel<__half>, dim3(blocks), dim3(256), 0, st,
                    static_cast<const __half*>(pyr->level_ptr[i - 1]), (long long)pyr->pitch[i - 1],
                    static_cast<__half*>(pyr->level_ptr[i]), (long long)pyr->pitch[i], pyr->width[i], total);
    else
      launch_kernel(fill_level_kernel<float>, dim3(blocks), dim3(256), 0, st,
                    static_cast<const float*>(pyr->level_ptr[i - 1]), (long long)pyr->pitch[i - 1],
                    static_cast<float*>(pyr->level_ptr[i]), (long long)pyr->pitch[i], pyr->width[i], total);
  }
  return check_launch("fill_levels");
}

int smoe_corr_fold_pyramid_grad(const smoe_pyramid_t* gp, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(gp, "fold_pyramid_grad")) return rc;
  SMOE_REQUIRE(gp->dtype == SMOE_DTYPE_F32, SMOE_ERR_UNSUPPORTED,
               "fold_pyramid_grad: gradient pyramid must be fp32");
  const long long rows = (long long)gp->B * gp->H * gp->W1;
  if (rows == 0 || gp->width[0] == 0 || gp->num_levels == 1) return SMOE_OK;
  const long long vpr = gp->width[0] / 4;
  const bool vec = gp->width[0] % 4 == 0 && aligned16(gp->level_ptr[0]) && (gp->pitch[0] % 4) == 0 &&
                   rows * vpr < (1ll << 31);
  if (vec) {
    const long long total = rows * vpr;
    const long long blocks = (total + 256 * kFoldUnroll - 1) / (256 * kFoldUnroll);
    launch_kernel(fold_grad_vec_kernel, dim3((unsigned)blocks), dim3(256), 0,
                  static_cast<cudaStream_t>(stream), to_dev(gp), (unsigned)total, (unsigned)vpr);
  } else {
    long long blocks = (rows + 7) / 8;          // 8 warps (rows) per block
    if (blocks > 148ll * 8) blocks = 148ll * 8;
    launch_kernel(fold_grad_scalar_kernel, dim3((unsigned)blocks), dim3(256), 0,
                  static_cast<cudaStream_t>(stream), to_dev(gp), rows);
  }
  return check_launch("fold_pyramid_grad");
}

}  // extern "C"
