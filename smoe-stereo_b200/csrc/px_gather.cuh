// px_gather.cuh -- one thread gathers the lookup windows of one level PAIR of one pixel.
// Shared by lookup_fwd_px_kernel (lookup.cu) and lookup_conv_px_kernel (lookup_conv.cu).
//
// Level i+1 is by construction the pair average of level i (core/corr.py:122-125), and the 2r+2
// taps of a pixel at level i+1 cover level-i indices [2*base_c, 2*base_c + 4r+3], which contain the
// pixel's level-i window.  So ONE contiguous level-i segment (20 elements for r = 4) per level pair
// holds both windows: the thread issues the NCH aligned 16-byte loads of the segment up front
// (load()), later recomputes the coarse taps as (a+b)*0.5f -- the operation the build epilogue
// used, so the values are bit-identical to the stored level -- and moves the data-dependent window
// start (one of EPL positions inside the aligned segment) to a fixed register position with
// log2(EPL) rounds of selects: no shared memory, no shuffles, no dynamic register indexing.
#pragma once
#include "common.cuh"

namespace smoe {

template <int N, int BITS>
__device__ __forceinline__ void shift_down(float (&a)[N], int s) {   // a[j] <- a[j + s], 0 <= s < 2^BITS
#pragma unroll
  for (int bit = BITS - 1; bit >= 0; --bit) {
    const bool on = (s >> bit) & 1;
#pragma unroll
    for (int j = 0; j + (1 << bit) < N; ++j) a[j] = on ? a[j + (1 << bit)] : a[j];
  }
}

template <typename VT, bool STREAM>
struct PxPair {
  static constexpr int R = 4, RD = 2 * R + 1;
  static constexpr int EPL = Vec16<VT>::N;
  static constexpr int LOG_EPL = (EPL == 4) ? 2 : 3;
  static constexpr int SEG = 4 * R + 4;                              // fine elements under the coarse window
  static constexpr int NCH = (EPL - 2 + SEG + EPL - 1) / EPL;        // chunks a segment can span
  static constexpr int NV = NCH * EPL;

  float v[NV];
  float dxf, dxc;
  int sh;       // S - EPL*cs: where the segment starts inside its first chunk (even)
  int odd;      // basef - S - R: 0 or 1 (-1: parked coordinate, everything is zero)
  int cs;       // first chunk of the segment

  // xl: the pixel's x coordinate in the FINE level of the pair; `fine` points at that level's row
  __device__ __forceinline__ void load(const VT* __restrict__ fine_row, int width, float xl, int /*hint*/ = 0) {
    int basef, basec;
    split_coord(xl, R, basef, dxf);
    split_coord(xl * 0.5f, R, basec, dxc);
    const int S = 2 * basec;
    cs = S >> LOG_EPL;
    sh = S - (cs << LOG_EPL);
    const int o = basef - S - R;
    odd = (o == 0 || o == 1) ? o : -1;
    const int nchunks = (width + EPL - 1) >> LOG_EPL;
    const VT* src = fine_row + (long long)cs * EPL;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int ci = cs + j;
      float t[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) t[e] = 0.f;
      // the last chunk is only needed when the segment does not start at the chunk boundary
      const bool need = odd >= 0 && (j < NCH - 1 || sh + SEG > (NCH - 1) * EPL) &&
                        (unsigned)ci < (unsigned)nchunks;
      if (need) Vec16<VT>::template load<STREAM>(src + j * EPL, t);
#pragma unroll
      for (int e = 0; e < EPL; ++e) v[j * EPL + e] = t[e];
    }
    if (__builtin_expect((cs + NCH) * EPL > width, 0)) {       // rare: segment reaches the row tail
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if (cs * EPL + k >= width) v[k] = 0.f;
    }
  }

  // Interpolated taps of the coarse level of the pair (width wc), in LT's arithmetic.  Call BEFORE
  // fine(): fine() shifts v in place.
  template <typename LT>
  __device__ __forceinline__ void coarse(int wc, float (&out)[RD]) const {
    // pc[j] is coarse column cs*EPL/2 + j; columns >= wc do not exist (negative ones come from
    // chunks that were never loaded: already zero)
    float pc[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
      float a = (v[2 * j] + v[2 * j + 1]) * 0.5f;
      if (sizeof(VT) == 2) a = __half2float(__float2half_rn(a));     // stored level is fp16
      pc[j] = a;
    }
    if (__builtin_expect((cs + NCH) * (EPL / 2) > wc, 0)) {
#pragma unroll
      for (int j = 0; j < NV / 2; ++j)
        if (cs * (EPL / 2) + j >= wc) pc[j] = 0.f;
    }
    shift_down<NV / 2, LOG_EPL - 1>(pc, sh >> 1);
    LerpTaps<LT, RD>::run(pc, dxc, out);
  }

  // Interpolated taps of the fine level: the window starts R + odd elements after the segment start.
  template <typename LT>
  __device__ __forceinline__ void fine(float (&out)[RD]) {
    shift_down<NV, LOG_EPL>(v, sh + (odd > 0 ? 1 : 0));
    LerpTaps<LT, RD>::run(v + R, dxf, out);
  }
};

// fp32 volumes, 256-bit loads (LDG.E.256, sm_100): the segment is fetched as up to four 32-byte
// chunks aligned to 32 bytes in MEMORY (level starts are only 16-byte aligned, so the first chunk
// may begin 4 floats before the level: those elements are masked).  3.25 load instructions per
// segment on average instead of 5.5 -- a third fewer L1 wavefronts, the busiest unit of the
// streaming lookup (profiles/r01/README.md) -- for 12 more registers per pair.
template <bool STREAM>
struct PxPair256 {
  static constexpr int R = 4, RD = 2 * R + 1;
  static constexpr int SEG = 4 * R + 4;
  static constexpr int NCH = 4, NV = 32;

  float v[NV];
  float dxf, dxc;
  int sh;       // where the segment starts inside the first 32-byte chunk: 0, 2, 4 or 6 floats
  int odd;      // basef - S - R: 0 or 1 (-1: parked coordinate, everything is zero)
  int i0;       // fine index of v[0] (may be negative)

  // hint (STREAM only): L2 eviction priority of the fill -- 0 normal, 1 evict_first (single-use level-0
  // segments of a volume that streams from DRAM), 2 evict_last (the 16x smaller level 2, which can stay
  // L2-resident across the iterations of a GRU loop).  SASS: LDG.E.NA.{EN,EF,EL}L2.LTC64B.256.
  __device__ __forceinline__ static void load32(const float* p, float (&t)[8], int hint) {
    if (STREAM) {
      if (hint == 1)
        asm volatile("ld.global.L1::no_allocate.L2::evict_first.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]), "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7])
                     : "l"(p));
      else if (hint == 2)
        asm volatile("ld.global.L1::no_allocate.L2::evict_last.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]), "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7])
                     : "l"(p));
      else
        asm volatile("ld.global.L1::no_allocate.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]), "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7])
                     : "l"(p));
    } else {
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(t[0]), "=f"(t[1]), "=f"(t[2]), "=f"(t[3]), "=f"(t[4]), "=f"(t[5]), "=f"(t[6]), "=f"(t[7])
                   : "l"(p));
    }
  }

  __device__ __forceinline__ void load(const float* __restrict__ fine_row, int width, float xl, int hint = 0) {
    int basef, basec;
    split_coord(xl, R, basef, dxf);
    split_coord(xl * 0.5f, R, basec, dxc);
    const int S = 2 * basec;
    const int o = basef - S - R;
    odd = (o == 0 || o == 1) ? o : -1;
    const int m = (int)((reinterpret_cast<uintptr_t>(fine_row) >> 2) & 7);   // 0 or 4 (16-byte aligned rows)
    const int c8 = (S + m) >> 3;                      // arithmetic shift == floor division
    sh = (S + m) - (c8 << 3);
    i0 = (c8 << 3) - m;
    const float* src = fine_row + i0;                 // 32-byte aligned
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const int ij = i0 + 8 * j;
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = 0.f;
      // the last chunk is only needed when the segment starts 6 floats into the first one
      const bool need = odd >= 0 && (j < NCH - 1 || sh + SEG > 8 * (NCH - 1)) && ij + 8 > 0 && ij < width;
      if (need) load32(src + 8 * j, t, hint);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[8 * j + e] = t[e];
    }
    if (__builtin_expect(i0 < 0 || i0 + NV > width, 0)) {      // rare: the chunks reach outside [0, width)
#pragma unroll
      for (int k = 0; k < NV; ++k)
        if ((unsigned)(i0 + k) >= (unsigned)width) v[k] = 0.f;
    }
  }

  template <typename LT>
  __device__ __forceinline__ void coarse(int wc, float (&out)[RD]) const {
    float pc[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) pc[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
    const int c0 = i0 >> 1;                           // coarse column of pc[0] (i0 is even)
    if (__builtin_expect(c0 < 0 || c0 + NV / 2 > wc, 0)) {
#pragma unroll
      for (int j = 0; j < NV / 2; ++j)
        if ((unsigned)(c0 + j) >= (unsigned)wc) pc[j] = 0.f;
    }
    shift_down<NV / 2, 2>(pc, sh >> 1);
    LerpTaps<LT, RD>::run(pc, dxc, out);
  }

  template <typename LT>
  __device__ __forceinline__ void fine(float (&out)[RD]) {
    shift_down<NV, 3>(v, sh + (odd > 0 ? 1 : 0));
    LerpTaps<LT, RD>::run(v + R, dxf, out);
  }
};

}  // namespace smoe
