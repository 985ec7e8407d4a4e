// xcheck_lookup_conv.cuh -- FIRST version of the fused lookup + 1x1 convolution (4 lanes per pixel,
// lane-group gather like lookup_fwd_vec_kernel).  Compiled only into the cross-check library
// (-DSMOE_CROSSCHECK); the product launches lookup_conv_px_kernel.  Included by lookup_conv.cu
// inside namespace smoe.
#pragma once

// VT: pyramid element type; LT: the type the unfused lookup would return (float for CorrBlock1D,
// the volume dtype for CorrBlockFast1D -- its interpolation arithmetic is reproduced);
// CT: output type of the convolution.
template <typename VT, typename LT, typename CT, int NL, bool STREAM>
__global__ void __launch_bounds__(kFcThreads)
lookup_conv_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                   float coord_scale, const float* __restrict__ weight, const float* __restrict__ bias,
                   int cout, int relu, uint32_t tmem_cols, CT* __restrict__ out, int HW, CoordUpd upd) {
  constexpr int R = 4;
  constexpr int RD = 2 * R + 1;
  constexpr int K = NL * RD;                          // input channels of the convolution
  constexpr int KMMA = (K + 7) / 8 * 8;               // K covered by the MMAs (zero padded)
  constexpr int EPL = Vec16<VT>::N;
  constexpr int LOG_EPL = (EPL == 4) ? 2 : 3;
  constexpr int QMAX = (EPL - 1 + 2 * R + 1) >> LOG_EPL;
  static_assert(KMMA <= kFcKPad, "K must fit the padded operand");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[kFcMaxCout];
  const uint32_t sA = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;    // swizzle atoms: 1 KiB aligned
  const uint32_t sB = sA + kFcABytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int p = tid >> 2, q = tid & 3;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kFcPx;
  const int hw = hw0 + p;
  const bool valid = hw < HW;

  if (warp == 0) {      // (allocating after the gathers are issued measured slower: 9.8 vs 8.8 us)
    ptx::tmem_alloc(ptx::smem_u32(&s_tmem), tmem_cols);
    ptx::tmem_relinquish();
  }
  if (tid == 32) {
    ptx::mbar_init(ptx::smem_u32(&s_bar), 1);
    ptx::fence_barrier_init();
  }
  pdl_wait();
  // ---- gather: issue the window loads of all levels first
  const float x0 = valid ? fetch_coord(coords, coords_bstride, upd, b, hw, q == 0) * coord_scale : 0.f;
  const long long row = (long long)b * HW + hw;
  float v[NL][EPL];
  float dxs[NL];
  int shift[NL];
  float lscale = 1.0f;
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    int base;
    split_coord(x0 * lscale, R, base, dxs[l]);
    lscale *= 0.5f;
    const int s = base & (EPL - 1);
    const int ci = (base >> LOG_EPL) + q;
    shift[l] = s;
    const int width = pyr.width[l];
    const unsigned nchunks = (unsigned)(width + EPL - 1) >> LOG_EPL;
    const bool need = valid && (q < QMAX || (q == QMAX && s + 2 * R + 1 >= QMAX * EPL)) &&
                      (unsigned)ci < nchunks;
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[l][j] = 0.f;
    if (need) {
      const VT* src = static_cast<const VT*>(pyr.ptr[l]) + row * pyr.pitch[l] + (long long)ci * EPL;
      Vec16<VT>::template load<STREAM>(src, v[l]);
      if (__builtin_expect((ci + 1) * EPL > width, 0)) {
#pragma unroll
        for (int j = 0; j < EPL; ++j)
          if (ci * EPL + j >= width) v[l][j] = 0.f;
      }
    }
  }
  // ---- weights -> B operand (tf32, zero padded to 64 columns) while the gathers are in flight
  for (int idx = tid; idx < cout * (kFcKPad / 4); idx += kFcThreads) {
    const int n = idx / (kFcKPad / 4), piece = idx % (kFcKPad / 4);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = piece * 4 + j;
      w[j] = k < K ? to_tf32(__ldg(weight + (long long)n * K + k)) : 0u;
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sB + kmajor_off(n, piece * 4, cout)),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
  if (tid < cout) s_bias[tid] = bias != nullptr ? __ldg(bias + tid) : 0.f;
  // zero the K padding the MMAs read (columns K .. KMMA-1 of every pixel row)
  if (q == 0) {
#pragma unroll
    for (int k = K; k < KMMA; ++k)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sA + kmajor_off(p, k, kFcPx)), "r"(0u) : "memory");
  }
  // ---- interpolate; tap k of level l sits at window position shift+k of the 4-chunk span
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const float nxt = __shfl_down_sync(0xffffffffu, v[l][0], 1);
    float e[EPL];
    LerpChunk<LT, EPL>::run(v[l], nxt, dxs[l], e);
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
      const int k = q * EPL + i - shift[l];
      if (k >= 0 && k < RD)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sA + kmajor_off(p, l * RD + k, kFcPx)),
                     "r"(to_tf32(e[i]))
                     : "memory");
    }
  }
  pdl_launch_dependents();
  ptx::fence_proxy_async_smem();          // generic-proxy operand writes -> tensor-core (async) proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    // c_format f32 | tf32 x tf32 | A, B K-major | N | M
    const uint32_t idesc = (1u << 4) | (ptx::kFmtTF32 << 7) | (ptx::kFmtTF32 << 10) |
                           (((uint32_t)cout >> 3) << 17) | ((uint32_t)(kFcPx >> 4) << 24);
#pragma unroll
    for (int ks = 0; ks < KMMA / 8; ++ks) {
      const int atom = (ks * 8) >> 5, kin = (ks * 8) & 31;
      // K-major SW128: 8-row groups 1 KiB apart (SBO), 32 bytes per UMMA_K inside the 128-byte row
      const uint64_t adesc = ptx::make_smem_desc(sA + atom * kFcPx * 128 + kin * 4, 16, 1024, ptx::kLayoutSwizzle128B);
      const uint64_t bdesc = ptx::make_smem_desc(sB + atom * cout * 128 + kin * 4, 16, 1024, ptx::kLayoutSwizzle128B);
      ptx::mma_tf32(tmem, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
    }
    ptx::tc_commit(ptx::smem_u32(&s_bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&s_bar), 0);
  ptx::tc_fence_after_sync();
  // ---- epilogue: TMEM lane == pixel; the 4 warps of a lane quadrant split the output channels
  const int quad = warp & 3, part = warp >> 2;
  const int px = quad * 32 + lane;
  const bool store = hw0 + px < HW;
  CT* obase = out + (long long)b * cout * HW + hw0 + px;
  for (int c0 = part * 16; c0 < cout; c0 += 64) {
    uint32_t acc[16];
    tmem_ld_32x32_x16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, acc);
    ptx::tmem_ld_wait();
    if (store) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float r = __uint_as_float(acc[j]) + s_bias[c0 + j];
        if (relu) r = fmaxf(r, 0.f);
        obase[(long long)(c0 + j) * HW] = from_float<CT>(r);
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem, tmem_cols);
  }
}

