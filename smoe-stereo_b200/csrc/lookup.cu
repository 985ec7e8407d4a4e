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
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "px_gather.cuh"

namespace smoe {

__device__ __forceinline__ uint32_t ptx_smem(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

constexpr int kTilePx = 32;       // pixels per CTA tile (one 128-byte line of fp32 outputs)
constexpr int kFwdThreads = 128;  // 4 lanes per pixel

// ----------------------------------------------------------------------------- forward (vector)
// Shared "window space" tile: for every (level, pixel) the 4*EPL interpolated values
// e[i] = lerp(v[i], v[i+1]) of the aligned chunks covering the pixel's window, plus the shift s
// (window start inside the first chunk).  Output channel k of that level is e[s + k]: the
// data-dependent shift is applied on the READ side, so the write side is one 16-byte store per
// chunk and the output loop is `LDS [row + s + k]` with immediate k.  Row stride WIN+4 words keeps
// the 16-byte alignment and spreads rows over banks (the kernel is issue/latency-bound at
// L2-resident sizes, so instruction count matters more than the residual 2-4 way conflicts:
// 416 -> ~290 SASS instructions per thread, profiles/r01/README.md).
template <int EPL> struct WinTile {
  static constexpr int kWin = 4 * EPL;
  static constexpr int kStride = kWin + 4;                 // words per (level,pixel) row
};

template <typename VT, typename OT, int R, int NL, int TPX, bool STREAM>
__global__ void __launch_bounds__(4 * TPX)
lookup_fwd_vec_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                      float coord_scale, OT* __restrict__ out, int HW, CoordUpd upd) {
  constexpr int EPL = Vec16<VT>::N;          // elements per 16-byte chunk
  constexpr int LOG_EPL = (EPL == 4) ? 2 : 3;
  constexpr int RD = 2 * R + 1;
  constexpr int QMAX = (EPL - 1 + 2 * R + 1) >> LOG_EPL;   // last chunk a window can reach
  using Tile = WinTile<EPL>;
  static_assert(EPL - 1 + 2 * R + 2 <= 4 * EPL, "window must fit the 4 chunks of a lane group");
  __shared__ __align__(16) float s_win[NL][TPX * Tile::kStride];
  __shared__ int s_shift[NL][TPX];

  pdl_wait();                // coords / pyramid come from earlier kernels in the stream
  const int tid = threadIdx.x;
  const int p = tid >> 2, q = tid & 3;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * TPX;
  const int hw = hw0 + p;
  const bool valid = hw < HW;
  const float x0 = valid ? fetch_coord(coords, coords_bstride, upd, b, hw, q == 0) * coord_scale : 0.f;
  const long long row = (long long)b * HW + hw;

  float v[NL][EPL];
  float dxs[NL];
  float lscale = 1.0f;
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    int base;
    split_coord(x0 * lscale, R, base, dxs[l]);
    lscale *= 0.5f;
    const int s = base & (EPL - 1);
    const int ci = (base >> LOG_EPL) + q;          // arithmetic shift == floor division
    if (q == 0) s_shift[l][p] = s;
    const int width = pyr.width[l];
    const unsigned nchunks = (unsigned)(width + EPL - 1) >> LOG_EPL;
    // chunk q intersects the window iff q*EPL <= s + 2R+1; one unsigned compare checks 0 <= ci < nchunks
    const bool need = valid && (q < QMAX || (q == QMAX && s + 2 * R + 1 >= QMAX * EPL)) &&
                      (unsigned)ci < nchunks;
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[l][j] = 0.f;
    if (need) {
      const VT* src = static_cast<const VT*>(pyr.ptr[l]) + row * pyr.pitch[l] + (long long)ci * EPL;
      Vec16<VT>::template load<STREAM>(src, v[l]);
      if (__builtin_expect((ci + 1) * EPL > width, 0)) {       // rare tail chunk: mask elements >= width
#pragma unroll
        for (int j = 0; j < EPL; ++j)
          if (ci * EPL + j >= width) v[l][j] = 0.f;
      }
    }
  }
  pdl_launch_dependents();   // all loads are in flight: let the successor's CTAs become resident
  float* my_row = &s_win[0][p * Tile::kStride + q * EPL];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const float nxt = __shfl_down_sync(0xffffffffu, v[l][0], 1);   // first element of chunk q+1
    const float dx = dxs[l];
    float e[EPL];
    LerpChunk<OT, EPL>::run(v[l], nxt, dx, e);
#pragma unroll
    for (int h = 0; h < EPL / 4; ++h)
      *reinterpret_cast<float4*>(my_row + l * (TPX * Tile::kStride) + 4 * h) =
          make_float4(e[4 * h], e[4 * h + 1], e[4 * h + 2], e[4 * h + 3]);
  }
  __syncthreads();
  // output phase: lane = pixel; a warp takes one (level, 32-pixel group) at a time, so every
  // store instruction writes one full 128-byte line of one output channel
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kGroups = TPX / 32, kWarps = TPX / 8;
  for (int item = warp; item < NL * kGroups; item += kWarps) {
    const int l = item % NL, pg = item / NL;
    const int pp = pg * 32 + lane;
    if (hw0 + pp < HW) {
      OT* dst = out + ((long long)b * (NL * RD) + l * RD) * HW + hw0 + pp;
      const float* wrow = &s_win[l][pp * Tile::kStride] + s_shift[l][pp];
#pragma unroll
      for (int k = 0; k < RD; ++k) dst[(long long)k * HW] = from_float<OT>(wrow[k]);
    }
  }
}

// ------------------------------------------------------------------------- forward (level pairs)
// Level i+1 is by construction the pair-average of level i (core/corr.py:122-125), and the 2r+2
// taps of a pixel at level i+1 cover level-i indices [2*base_c, 2*base_c + 4r+3], which CONTAIN
// the pixel's level-i window.  So the lookup reads ONE contiguous level-i segment (20 elements
// for r=4) per level pair (0,1) and (2,3) -- half the scattered DRAM accesses and fewer sectors
// than four separate 10-element windows -- and recomputes the coarse taps as (a+b)*0.5f, the same
// fp32 operation the build epilogue used, so the result is bit-identical to reading levels 1/3.
// 8 lanes per pixel: lane q fetches aligned chunk q of both segments (<= 6 chunks for fp32,
// <= 4 for fp16).  Window-space tiles + read-side shifts as in lookup_fwd_vec_kernel.
template <typename VT, typename OT, int R, int NP, int TPX, bool STREAM>
__global__ void __launch_bounds__(8 * TPX)
lookup_fwd_pair_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                       float coord_scale, OT* __restrict__ out, int HW, CoordUpd upd) {
  constexpr int EPL = Vec16<VT>::N;
  constexpr int LOG_EPL = (EPL == 4) ? 2 : 3;
  constexpr int RD = 2 * R + 1;
  constexpr int SEG = 4 * R + 4;                              // fine elements under the coarse window
  constexpr int NCH = (EPL - 2 + SEG + EPL - 1) / EPL;        // chunks a segment can span
  static_assert(NCH <= 8, "segment must fit the 8 lanes of a pixel");
  constexpr int FW = 8 * EPL;                                 // fine window-space words per pixel
  constexpr int FS = FW + 4;                                  // row stride (bank spreading)
  constexpr int CW = 4 * EPL;                                 // coarse window-space words per pixel
  constexpr int CS = CW + 4;
  __shared__ __align__(16) float s_fine[NP][TPX * FS];
  __shared__ __align__(16) float s_coarse[NP][TPX * CS];
  __shared__ int s_shift[2 * NP][TPX];

  pdl_wait();
  const int tid = threadIdx.x;
  const int p = tid >> 3, q = tid & 7;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * TPX;
  const int hw = hw0 + p;
  const bool valid = hw < HW;
  const float x0 = valid ? fetch_coord(coords, coords_bstride, upd, b, hw, q == 0) * coord_scale : 0.f;
  const long long row = (long long)b * HW + hw;

  float v[NP][EPL];
  float dxf[NP], dxc[NP];
  int ci0[NP];
  float lscale = 1.0f;
#pragma unroll
  for (int pr = 0; pr < NP; ++pr) {
    const int lf = 2 * pr;
    int basef, basec;
    split_coord(x0 * lscale, R, basef, dxf[pr]);
    split_coord(x0 * lscale * 0.5f, R, basec, dxc[pr]);
    lscale *= 0.25f;
    // the segment follows the coarse window unless the coordinate is out of the representable
    // range (split_coord parks both far outside, everything masks to zero)
    const int S = 2 * basec;
    const int cs = S >> LOG_EPL;                               // first chunk of the segment
    const int ci = cs + q;
    ci0[pr] = cs;
    if (q == 0) {
      s_shift[2 * pr][p] = basef - (cs << LOG_EPL);            // fine tap k  -> fine tile index
      s_shift[2 * pr + 1][p] = basec - (cs << (LOG_EPL - 1));  // coarse tap k -> coarse tile index
    }
    const int width = pyr.width[lf];
    const int nchunks = (width + EPL - 1) >> LOG_EPL;
    const bool need = valid && q < NCH && (ci << LOG_EPL) < S + SEG && ci >= 0 && ci < nchunks;
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[pr][j] = 0.f;
    if (need) {
      const VT* src = static_cast<const VT*>(pyr.ptr[lf]) + row * pyr.pitch[lf] + (long long)ci * EPL;
      Vec16<VT>::template load<STREAM>(src, v[pr]);
      if ((ci + 1) * EPL > width) {
#pragma unroll
        for (int j = 0; j < EPL; ++j)
          if (ci * EPL + j >= width) v[pr][j] = 0.f;
      }
    }
  }
  pdl_launch_dependents();
#pragma unroll
  for (int pr = 0; pr < NP; ++pr) {
    const int wc = (2 * pr + 1 < pyr.num_levels) ? pyr.width[2 * pr + 1] : 0;
    const int ci = ci0[pr] + q;
    // coarse values owned by this lane: pair averages, zero where the coarse level has no column
    float pc[EPL / 2];
#pragma unroll
    for (int j = 0; j < EPL / 2; ++j) {
      const int cidx = ci * (EPL / 2) + j;
      float a = (v[pr][2 * j] + v[pr][2 * j + 1]) * 0.5f;
      if (sizeof(VT) == 2) a = __half2float(__float2half_rn(a));     // stored level is fp16
      pc[j] = (cidx >= 0 && cidx < wc) ? a : 0.f;
    }
    const float nxt_f = __shfl_down_sync(0xffffffffu, v[pr][0], 1);
    const float nxt_c = __shfl_down_sync(0xffffffffu, pc[0], 1);
    float ef[EPL], ec[EPL / 2];
    LerpChunk<OT, EPL>::run(v[pr], nxt_f, dxf[pr], ef);
    LerpChunk<OT, EPL / 2>::run(pc, nxt_c, dxc[pr], ec);
#pragma unroll
    for (int h = 0; h < EPL / 4; ++h)
      *reinterpret_cast<float4*>(&s_fine[pr][p * FS + (q * (EPL / 4) + h) * 4]) =
          make_float4(ef[4 * h], ef[4 * h + 1], ef[4 * h + 2], ef[4 * h + 3]);
    if (EPL == 4) {
      *reinterpret_cast<float2*>(&s_coarse[pr][p * CS + q * 2]) = make_float2(ec[0], ec[1]);
    } else {
      *reinterpret_cast<float4*>(&s_coarse[pr][p * CS + q * 4]) =
          make_float4(ec[0], ec[1], ec[EPL == 4 ? 0 : 2], ec[EPL == 4 ? 1 : 3]);
    }
  }
  __syncthreads();
  // output phase: lane = pixel; a warp takes one (level, 32-pixel group): full-line stores
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kGroups = TPX / 32, kWarps = TPX / 4;
  const int NL = pyr.num_levels;
  for (int item = warp; item < NL * kGroups; item += kWarps) {
    const int l = item % NL, pg = item / NL;
    const int pp = pg * 32 + lane;
    if (hw0 + pp < HW) {
      OT* dst = out + ((long long)b * (NL * RD) + l * RD) * HW + hw0 + pp;
      const int sft = s_shift[l][pp];
      const float* wrow = (l & 1) ? &s_coarse[l >> 1][pp * CS] : &s_fine[l >> 1][pp * FS];
      const int lim = (l & 1) ? CW : FW;
#pragma unroll
      for (int k = 0; k < RD; ++k) {
        const int e = sft + k;                                  // parked coordinates: out of tile -> 0
        dst[(long long)k * HW] = from_float<OT>((e >= 0 && e < lim) ? wrow[e] : 0.f);
      }
    }
  }
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
// LD256: fp32 volumes fetched with 256-bit loads (PxPair256; opt-in, SMOE_CORR_PX_LD256=1).
template <typename VT, typename OT, int R, int NL, bool STREAM, bool SPLIT, bool LD256 = false>
__global__ void __launch_bounds__(256)
lookup_fwd_px_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                     float coord_scale, OT* __restrict__ out, int HW, CoordUpd upd) {
  static_assert(!LD256 || sizeof(VT) == 4, "256-bit segment loads are implemented for fp32 volumes");
  using Pair = typename std::conditional<LD256, PxPair256<STREAM>, PxPair<VT, STREAM>>::type;
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
                 first ? pyr.width[0] : pyr.width[2], x0 * (pr == 0 ? 1.0f : 0.25f));
  }
  pdl_launch_dependents();

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
      for (int k = 0; k < RD; ++k) d1[(long long)k * HW] = from_float<OT>(e[k]);
    }
    pair[i].template fine<OT>(e);
    OT* d0 = dst + (long long)lf * RD * HW;
#pragma unroll
    for (int k = 0; k < RD; ++k) d0[(long long)k * HW] = from_float<OT>(e[k]);
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

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// elements per grad-tile row: a 16-byte multiple (cp.async) that is not a multiple of 32 words
// elements per grad-tile row: 40 words for fp32 puts the 4 levels of a pixel (9 rows apart) 8 banks apart
// and the 8 pixels of a warp in between: conflict-free tile reads; fp16 rows only need 16-byte alignment
template <typename GT> struct BatchedTile { static constexpr int kStride = kTilePx + 32 / (int)sizeof(GT); };
constexpr int kBatchedThreads = 256;   // two threads per (pixel, level): taps 0-4 and 5-9

template <typename GT, int NL>
__global__ void __launch_bounds__(kBatchedThreads)
lookup_bwd_batched_kernel(const __grid_constant__ BwdBatch batch, const BatchedGeom geo,
                          float* __restrict__ g0_out, long long g0_pitch, int HW, int accumulate) {
  constexpr int R = 4, RD = 9;
  constexpr int NC = NL * RD;
  constexpr int kBatchedTileStride = BatchedTile<GT>::kStride;
  constexpr int kTileBytes = NC * kBatchedTileStride * (int)sizeof(GT);   // one iteration's grad tile
  extern __shared__ __align__(16) uint8_t s_raw[];
  float* s_rows = reinterpret_cast<float*>(s_raw);                    // [kTilePx][rowlen]
  uint8_t* s_tiles = s_raw + (size_t)kTilePx * geo.rowlen * 4;        // [stages][NC][kBatchedTileStride] GT
  float* s_x = reinterpret_cast<float*>(s_tiles + kBatchedStages * kTileBytes);   // [stages][kTilePx]

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kTilePx;
  const int npx = min(kTilePx, HW - hw0);

  auto issue = [&](int t, int buf) {
    // grad tile: NC rows of kTilePx elements; 16-byte pieces, zero-filled past the end of the plane
    constexpr int kPiecesPerRow = kTilePx * (int)sizeof(GT) / 16;
    constexpr int kElemsPerPiece = 16 / (int)sizeof(GT);
    const GT* g = static_cast<const GT*>(batch.gout[t]) + (long long)b * NC * HW + hw0;
    for (int i = tid; i < NC * kPiecesPerRow; i += kBatchedThreads) {
      const int c = i / kPiecesPerRow, e0 = (i - c * kPiecesPerRow) * kElemsPerPiece;
      const int bytes = max(0, min(16, (npx - e0) * (int)sizeof(GT)));
      const GT* src = g + (long long)c * HW + (bytes > 0 ? e0 : 0);
      cp_async16(ptx_smem(s_tiles + buf * kTileBytes + (c * kBatchedTileStride + e0) * sizeof(GT)), src, bytes);
    }
    if (tid < kTilePx / 4) {
      const int e0 = tid * 4;
      const int bytes = max(0, min(16, (npx - e0) * 4));
      const float* src = batch.coords[t] + (long long)b * batch.coords_bstride[t] + hw0 + (bytes > 0 ? e0 : 0);
      cp_async16(ptx_smem(s_x + buf * kTilePx + e0), src, bytes);
    }
    cp_async_commit();
  };

  for (int t = 0; t < kBatchedStages; ++t) {
    if (t < batch.T) issue(t, t); else cp_async_commit();     // exactly one group per slot
  }
  // rows start from zero (or from the existing level-0 gradient when accumulating)
  for (int i = tid; i < kTilePx * geo.rowlen / 4; i += kBatchedThreads)
    reinterpret_cast<float4*>(s_rows)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (accumulate) {
    __syncthreads();
    for (int pp = 0; pp < npx; ++pp)
      for (int j = tid; j < geo.width[0]; j += kBatchedThreads)
        s_rows[pp * geo.rowlen + j] = g0_out[((long long)b * HW + hw0 + pp) * g0_pitch + j];
  }

  // one thread per (pixel, level): its 2r+2 taps are contiguous in the shared-memory row, so there
  // is no 16-byte chunking (that only matters for global memory) and no idle slot
  const int p = (tid >> 2) & (kTilePx - 1), l = tid & 3, half = tid >> 7;   // half: taps 0-4 / 5-9
  const bool active = p < npx && l < NL;
  float* rowl = s_rows + p * geo.rowlen + geo.off[l];
  const int width = geo.width[l];
  const float lscale = 1.0f / (float)(1 << l);
  for (int t = 0; t < batch.T; ++t) {
    const int buf = t % kBatchedStages;
    // groups committed so far: kBatchedStages + (t-1) (the refill of this iteration comes after the
    // wait), so at most kBatchedStages-2 may still be pending for tile t to have landed
    cp_async_wait<kBatchedStages - 2>();
    __syncthreads();                       // tile t visible to everyone; everyone finished tile t-1
    // refill the buffer consumed in the PREVIOUS iteration (one barrier per iteration suffices)
    if (t >= 1) {
      if (t - 1 + kBatchedStages < batch.T) issue(t - 1 + kBatchedStages, (t - 1) % kBatchedStages);
      else cp_async_commit();
    }
    if (active) {
      const GT* tl = reinterpret_cast<const GT*>(s_tiles + buf * kTileBytes) + (l * RD) * kBatchedTileStride + p;
      int base;
      float dx;
      split_coord(s_x[buf * kTilePx + p] * lscale, R, base, dx);
      // this thread's 5 taps: all shared-memory loads first, then the math, then the stores, so
      // the read-modify-writes do not serialise on the shared-memory latency
      const int t0 = half * 5;
      float cur[5], g[6];
      bool ok[5];
      g[0] = half ? to_float(tl[(t0 - 1) * kBatchedTileStride]) : 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        g[k + 1] = (t0 + k < RD) ? to_float(tl[(t0 + k) * kBatchedTileStride]) : 0.f;
        const int idx = base + t0 + k;
        ok[k] = (unsigned)idx < (unsigned)width;
        cur[k] = ok[k] ? rowl[idx] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 5; ++k)
        cur[k] += BwdTap<float>::eval(g[k], g[k + 1], t0 + k > 0, t0 + k < RD, dx);
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (ok[k]) rowl[base + t0 + k] = cur[k];
    }
  }
  __syncthreads();                         // all scatters done before the fold reads the rows
  // fold the levels in shared memory and write level 0 (same chain as fold_grad_*_kernel);
  // which coarser levels cover column j depends on j only, so it is resolved once per column
  const int w0 = geo.width[0];
  for (int j = tid; j < w0; j += kBatchedThreads) {
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

// Pixels per CTA (32 / 64; 4 lanes per pixel).  SMOE_CORR_LOOKUP_TILE overrides for experiments.
static int pick_fwd_tile(long long px) {
  static const int forced = [] { const char* e = getenv("SMOE_CORR_LOOKUP_TILE"); return e ? atoi(e) : 0; }();
  if (forced == 32 || forced == 64) return forced;
  (void)px;
  return 32;   // measured best at every size tried (cfg1: 3.65/3.80/4.04 us, cfg3: 32.8/33.7/35.5 us)
}

// A pyramid that cannot stay L2-resident (126 MB L2, minus the lookup's own outputs) is read with
// the streaming load variant (Vec16).
static bool volume_exceeds_l2(const PyrDev& d, long long rows, int elem_bytes) {
  return rows * d.pitch[0] * elem_bytes > (96ll << 20);   // (64 MB tried: worse for a 90 MB fp16 volume)
}

template <typename VT, typename OT, int NL>
static void launch_fwd_vec(const PyrDev& d, const float* coords, long long cbs, float cs, void* out,
                           int HW, int B, cudaStream_t st, const CoordUpd& upd) {
  int tile = pick_fwd_tile((long long)HW * B);
  if (tile != 64) tile = 32;
  dim3 grid((HW + tile - 1) / tile, B);
  const bool big = volume_exceeds_l2(d, (long long)HW * B, (int)sizeof(VT));
#define SMOE_LAUNCH_FWD(TPX, STREAM)                                                               \
  launch_kernel(lookup_fwd_vec_kernel<VT, OT, 4, NL, TPX, STREAM>, grid, dim3(4 * TPX), 0, st, d,  \
                coords, cbs, cs, static_cast<OT*>(out), HW, upd)
  if (tile == 64) {
    if (big) SMOE_LAUNCH_FWD(64, true); else SMOE_LAUNCH_FWD(64, false);
  } else {
    if (big) SMOE_LAUNCH_FWD(32, true); else SMOE_LAUNCH_FWD(32, false);
  }
#undef SMOE_LAUNCH_FWD
}

static const bool g_no_pair = getenv("SMOE_CORR_NO_PAIR_LOOKUP") != nullptr;      // A/B switches
static const bool g_force_pair = getenv("SMOE_CORR_FORCE_PAIR_LOOKUP") != nullptr;

template <typename VT, typename OT, int NP>
static void launch_fwd_pair(const PyrDev& d, const float* coords, long long cbs, float cs, void* out,
                            int HW, int B, cudaStream_t st, const CoordUpd& upd) {
  constexpr int TPX = 32;
  dim3 grid((HW + TPX - 1) / TPX, B);
  const bool big = volume_exceeds_l2(d, (long long)HW * B, (int)sizeof(VT));
  if (big)
    launch_kernel(lookup_fwd_pair_kernel<VT, OT, 4, NP, TPX, true>, grid, dim3(8 * TPX), 0, st, d, coords,
                  cbs, cs, static_cast<OT*>(out), HW, upd);
  else
    launch_kernel(lookup_fwd_pair_kernel<VT, OT, 4, NP, TPX, false>, grid, dim3(8 * TPX), 0, st, d, coords,
                  cbs, cs, static_cast<OT*>(out), HW, upd);
}

// SMOE_CORR_LOOKUP_KERNEL=px|pair|vec forces one forward kernel (A/B experiments, parity tests)
static const int g_fwd_kernel_env = [] {
  const char* e = getenv("SMOE_CORR_LOOKUP_KERNEL");
  if (!e) return 0;
  return e[0] == 'p' && e[1] == 'x' ? 1 : (e[0] == 'p' ? 2 : (e[0] == 'v' ? 3 : 0));
}();
static thread_local int g_fwd_kernel_tls = -1;       // smoe_corr_debug_set_lookup_kernel
#define g_fwd_kernel (g_fwd_kernel_tls >= 0 ? g_fwd_kernel_tls : g_fwd_kernel_env)

// experiment switches for the thread-per-pixel kernel: pixels per CTA, one thread per level pair
static const int g_px_tpb = [] { const char* e = getenv("SMOE_CORR_PX_TPB"); return e ? atoi(e) : 0; }();
static const int g_px_split = [] { const char* e = getenv("SMOE_CORR_PX_SPLIT"); return e ? atoi(e) : -1; }();

template <typename VT, typename OT, int NL>
static void launch_fwd_px(const PyrDev& d, const float* coords, long long cbs, float cs, void* out,
                          int HW, int B, cudaStream_t st, const CoordUpd& upd) {
  constexpr int NP = (NL + 1) / 2;
  const bool big = volume_exceeds_l2(d, (long long)HW * B, (int)sizeof(VT));
  // measured (tools/px_sweep.sh, bench.py): 32 pixels per CTA and one thread per level pair
  // everywhere -- beyond L2 (cfg2 10.2 -> 8.9 us, cfg3 fp16 18.6 -> 12.4 us), for fp16 volumes
  // (cfg1 3.40 -> 3.16 us) and, inside a real step (cold coordinates, distinct output buffers),
  // for L2-resident fp32 volumes too: 3.38 -> 3.10 us per lookup of the cfg1 bench step, although
  // a lookups-only graph with warm coordinates had shown 3.02 vs 3.08 us
  const bool split = NP > 1 && (g_px_split >= 0 ? g_px_split != 0 : true);
  int tpb = (g_px_tpb == 32 || g_px_tpb == 64 || g_px_tpb == 128) ? g_px_tpb : 32;
  dim3 grid((HW + tpb - 1) / tpb, B);
  OT* o = static_cast<OT*>(out);
  if constexpr (sizeof(VT) == 4 && sizeof(OT) == 4) {
    // 256-bit segment loads for fp32 volumes that stream from DRAM (tools/px256_check.py, 32 distinct
    // outputs per graph: cfg2 12.3 -> 8.4 us, cfg3 24.2 -> 20.5 us; L2-resident cfg1 3.08 -> 3.15 us,
    // so those keep the 16-byte loads).  SMOE_CORR_PX_LD256=1 / 0 forces them on / off.  Level-0 rows
    // must be 32-byte aligned (the first chunk of a level may start 16 bytes before it, which is
    // only inside the pixel row for levels > 0).
    static const int env256 = [] { const char* e = getenv("SMOE_CORR_PX_LD256"); return e ? (e[0] != '0' ? 1 : 0) : -1; }();
    const bool ok256 = (reinterpret_cast<uintptr_t>(d.ptr[0]) & 31) == 0 && (d.pitch[0] * 4) % 32 == 0;
    const bool want256 = g_fwd_kernel == 4 || (g_fwd_kernel == 0 && (env256 >= 0 ? env256 == 1 : big));
    if (want256 && ok256) {
      constexpr bool kSplit = NP > 1;
      dim3 block(tpb, kSplit ? NP : 1);
      if (big) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, true, kSplit, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
      else launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, false, kSplit, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
      return;
    }
  }
  if (split) {
    constexpr int NLS = NP > 1 ? NL : 4;      // (never launched for NP == 1; avoids instantiating it)
    dim3 block(tpb, NP);
    if (big) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NLS, true, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
    else launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NLS, false, true>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
  } else {
    dim3 block(tpb);
    if (big) launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, true, false>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
    else launch_kernel_pdl(lookup_fwd_px_kernel<VT, OT, 4, NL, false, false>, grid, block, 0, st, d, coords, cbs, cs, o, HW, upd);
  }
}

template <typename VT, typename OT>
static void dispatch_fwd_vec(int NL, const PyrDev& d, const float* coords, long long cbs, float cs,
                             void* out, int HW, int B, cudaStream_t st, const CoordUpd& upd) {
  // Default: the thread-per-pixel kernel (faster than both lane-group kernels at every size and
  // dtype measured, tools/lookup_kernels_probe.py).  The lane-group kernels stay selectable
  // (smoe_corr_debug_set_lookup_kernel / SMOE_CORR_LOOKUP_KERNEL) as bit-exact cross-checks.
  if (g_fwd_kernel == 0 || g_fwd_kernel == 1 || g_fwd_kernel == 4) {
    switch (NL) {
      case 1: launch_fwd_px<VT, OT, 1>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      case 2: launch_fwd_px<VT, OT, 2>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      case 3: launch_fwd_px<VT, OT, 3>(d, coords, cbs, cs, out, HW, B, st, upd); break;
      default: launch_fwd_px<VT, OT, 4>(d, coords, cbs, cs, out, HW, B, st, upd); break;
    }
    return;
  }
  // Volumes that stream from DRAM: two contiguous segments per pixel instead of four windows
  // (cfg3: 31.3 -> 26.8 us).  L2-resident volumes stay on the 4-lane kernel (3.65 vs 4.04 us at
  // cfg1: that regime is latency-bound and the pair kernel does more work per pixel).
  const bool pair = (volume_exceeds_l2(d, (long long)HW * B, (int)sizeof(VT)) || g_force_pair ||
                     g_fwd_kernel == 2) && g_fwd_kernel != 3;
  if (!g_no_pair && pair && (NL == 2 || NL == 4)) {
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
    SMOE_REQUIRE(radius == 4 && pyramid_vec_ok(pyr) && (g_fwd_kernel == 0 || g_fwd_kernel == 1 || g_fwd_kernel == 4),
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

void smoe_corr_debug_set_lookup_kernel(int which) { smoe::g_fwd_kernel_tls = which; }

int smoe_corr_lookup(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                     float coord_scale, int radius, void* out, int out_dtype,
                     smoe_stream_t stream) {
  const smoe::CoordUpd none = {nullptr, nullptr, nullptr, nullptr, 0};
  return smoe::lookup_impl(pyr, coords, coords_batch_stride, coord_scale, radius, out, out_dtype, stream, none);
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
  // SMOE_CORR_BWD_BATCHED=staged selects the first version (cp.async-staged tiles, 16-byte rules)
  static const bool staged = [] { const char* e = getenv("SMOE_CORR_BWD_BATCHED"); return e && e[0] == 's'; }();
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
  // consecutive pixel rows must start 4 banks apart (rowlen % 32 == 4): the 8 pixels of a warp
  // then cover all 32 banks.  (rowlen % 32 == 0, which W2=184 happened to give with a fixed pad,
  // made every scatter an 8-way bank conflict: 810 us instead of ~200 us at config 2.)
  geo.rowlen = off + ((4 - off % 32) + 32) % 32;
  // warp-per-level kernel: lane = pixel, so consecutive lanes hit rows `rowlen` apart at columns
  // that advance by 2^-level per pixel; rowlen % 32 == 2 makes level 0 conflict-free (3 banks per
  // pixel) and keeps the coarser levels at <= 2-way
  if (!staged) geo.rowlen = off + ((2 - off % 32) + 32) % 32;
  const size_t smem = (size_t)kTilePx * geo.rowlen * 4 +
                      (staged ? kBatchedStages * ((size_t)num_levels * 9 * (kTilePx + 32 / ges) * ges + kTilePx * 4) : 0);
  // (the Python side bounds this with off + 32 words per row)
  SMOE_REQUIRE(smem <= 227 * 1024, SMOE_ERR_UNSUPPORTED,
               "lookup_bwd_batched: W2=%d needs %zu bytes of shared memory per CTA", g0->width[0], smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((HW + kTilePx - 1) / kTilePx), g0->B);
  float* out = static_cast<float*>(g0->level_ptr[0]);
#define SMOE_LAUNCH_BATCHED(GT, NLV)                                                                     \
  do {                                                                                                   \
    if (staged) {                                                                                        \
      SMOE_CUDA_OK(cudaFuncSetAttribute(lookup_bwd_batched_kernel<GT, NLV>,                              \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
      SMOE_CUDA_OK(launch_kernel(lookup_bwd_batched_kernel<GT, NLV>, grid, dim3(kBatchedThreads), smem,  \
                                 st, batch, geo, out, (long long)g0->pitch[0], (int)HW, accumulate));     \
    } else {                                                                                             \
      SMOE_CUDA_OK(cudaFuncSetAttribute(lookup_bwd_lvl_kernel<GT, NLV>,                                  \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
      SMOE_CUDA_OK(launch_kernel(lookup_bwd_lvl_kernel<GT, NLV>, grid, dim3(128), smem, st, batch, geo,  \
                                 out, (long long)g0->pitch[0], (int)HW, accumulate));                     \
    }                                                                                                    \
  } while (0)
  const bool f16g = grad_out_dtype == SMOE_DTYPE_F16;
  switch (num_levels) {
    case 1: if (f16g) SMOE_LAUNCH_BATCHED(__half, 1); else SMOE_LAUNCH_BATCHED(float, 1); break;
    case 2: if (f16g) SMOE_LAUNCH_BATCHED(__half, 2); else SMOE_LAUNCH_BATCHED(float, 2); break;
    case 3: if (f16g) SMOE_LAUNCH_BATCHED(__half, 3); else SMOE_LAUNCH_BATCHED(float, 3); break;
    default: if (f16g) SMOE_LAUNCH_BATCHED(__half, 4); else SMOE_LAUNCH_BATCHED(float, 4); break;
  }
#undef SMOE_LAUNCH_BATCHED
  return check_launch("lookup backward (batched)");
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
