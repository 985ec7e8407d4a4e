// xcheck_lookup.cuh -- SUPERSEDED lookup kernels, compiled only into the cross-check library
// (-DSMOE_CROSSCHECK -> tests/_build/libsmoecorr_xcheck.so; never part of libsmoecorr.so).
// They are the earlier designs of the forward lookup (4 and 8 lanes per pixel) and of the batched
// backward (cp.async-staged tiles); tests/ compare the product kernels with them bit for bit, and
// tools/ use them for before/after measurements (profiles/r01/README.md has their history).
// Included by lookup.cu inside namespace smoe, after the shared helpers.
#pragma once

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

// Pixels per CTA (32 / 64; 4 lanes per pixel).  SMOE_CORR_LOOKUP_TILE overrides for experiments.
static int pick_fwd_tile(long long px) {
  (void)px;
  const int forced = tuning().lookup_tile;
  return forced == 64 ? 64 : 32;   // measured best at every size tried (cfg1: 3.65/3.80/4.04 us, cfg3: 32.8/33.7/35.5 us)
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
