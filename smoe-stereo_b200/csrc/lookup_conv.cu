// lookup_conv.cu -- pyramid lookup fused with the 1x1 convolution (+bias, +ReLU) that consumes it
// (SURVEY.md 8f row f2): the first layer of the reference's motion encoder,
//     cor = F.relu(self.convc1(corr))          core/update.py:70,77   (Conv2d(L*(2r+1), 64, 1))
// so the L*(2r+1)-channel lookup result never goes to global memory.
//
// One CTA = 128 pixels = one UMMA M tile.  512 threads gather exactly like lookup_fwd_vec_kernel
// (4 lanes per pixel, one aligned 16-byte chunk per lane and level, neighbour exchange by shuffle),
// but instead of an output tile they write the interpolated taps, rounded to tf32, straight into
// the K-major SWIZZLE_128B operand layout of tcgen05 (A = [128 px x K], K = L*9 padded to 64).
// The weights (B = [Cout x K]) are staged the same way, one thread issues K/8 tcgen05.mma
// (kind::tf32, fp32 accumulate in TMEM), and the epilogue reads TMEM with lane == pixel: for a
// fixed output channel the 32 lanes of a warp hold 32 consecutive pixels, i.e. one 128-byte line
// of the channel-major output -- the accumulator layout does the transpose.
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "px_gather.cuh"

namespace smoe {

constexpr int kFcPx = 128;                 // pixels per CTA == UMMA M
constexpr int kFcThreads = 4 * kFcPx;      // 4 lanes per pixel
constexpr int kFcKPad = 64;                // K padded to two 32-element swizzle atoms
constexpr int kFcABytes = 2 * kFcPx * 128;
constexpr int kFcMaxCout = 256;

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
// byte offset of element (row, k) in a K-major SWIZZLE_128B operand of `rows` rows: 32-element
// (128-byte) atoms along K, rows 128 bytes apart, 16-byte pieces XORed with (row & 7)
__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int rows) {
  const int atom = k >> 5, kk = k & 31;
  return (uint32_t)(atom * rows * 128 + row * 128 + ((((kk >> 2) ^ (row & 7)) << 4) | ((kk & 3) << 2)));
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// VT: pyramid element type; LT: the type the unfused lookup would return (float for CorrBlock1D,
// the volume dtype for CorrBlockFast1D -- its interpolation arithmetic is reproduced);
// CT: output type of the convolution.
#ifdef SMOE_CROSSCHECK
#include "xcheck_lookup_conv.cuh"     // first version (4 lanes per pixel), cross-check library only
#endif

// Second version: thread = pixel (PxPair gather, px_gather.cuh).  128 threads per CTA: each thread
// gathers its pixel's two segments (12 independent 16-byte loads), stages its share of the weights
// while they are in flight, writes its own operand row (ten conflict-free 16-byte stores into the
// K-major SWIZZLE_128B tile), one thread issues the MMAs, and the same thread then reads its own
// TMEM lane (lane == pixel) for the epilogue.  A quarter of the threads and ~3x fewer instructions
// than the lane-group version, four CTAs per SM instead of two.
// LD256: fp32 volumes beyond L2 fetch their segments with 256-bit loads (PxPair256), level-0 fills
// evict_first, like the plain lookup.
template <typename VT, typename LT, typename CT, int NL, bool STREAM, bool LD256 = false>
__global__ void __launch_bounds__(kFcPx)
lookup_conv_px_kernel(PyrDev pyr, const float* __restrict__ coords, long long coords_bstride,
                      float coord_scale, const float* __restrict__ weight, const float* __restrict__ bias,
                      int cout, int relu, uint32_t tmem_cols, CT* __restrict__ out, int HW, CoordUpd upd) {
  static_assert(!LD256 || (sizeof(VT) == 4 && STREAM), "256-bit segment loads: streaming fp32 volumes");
  using Pair = typename std::conditional<LD256, PxPair256<STREAM>, PxPair<VT, STREAM>>::type;
  using ET = typename std::conditional<LD256, float, VT>::type;
  constexpr int RD = Pair::RD;
  constexpr int NP = (NL + 1) / 2;
  constexpr int K = NL * RD;                          // input channels of the convolution
  constexpr int KMMA = (K + 7) / 8 * 8;               // K covered by the MMAs (zero padded)
  constexpr int KP4 = (KMMA + 3) / 4;                 // 16-byte pieces of an operand row that are read
  static_assert(KMMA <= kFcKPad && NP <= 2, "K must fit the padded operand");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[kFcMaxCout];
  const uint32_t sA = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;    // swizzle atoms: 1 KiB aligned
  const uint32_t sB = sA + kFcABytes;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * kFcPx;
  const int hw = hw0 + tid;
  const bool valid = hw < HW;

  if (warp == 0) {
    ptx::tmem_alloc(ptx::smem_u32(&s_tmem), tmem_cols);
    ptx::tmem_relinquish();
  }
  if (tid == 32) {
    ptx::mbar_init(ptx::smem_u32(&s_bar), 1);
    ptx::fence_barrier_init();
  }
  pdl_wait();
  // ---- gather: issue the segment loads of both level pairs first
  const float x0 = valid ? fetch_coord(coords, coords_bstride, upd, b, hw, true) * coord_scale : 0.f;
  const long long row = (long long)b * HW + (valid ? hw : 0);
  Pair pair[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const VT* fine_row = static_cast<const VT*>(i == 0 ? pyr.ptr[0] : pyr.ptr[2]) +
                         row * (i == 0 ? pyr.pitch[0] : pyr.pitch[2]);
    // an invalid (tail) pixel gathers with a non-finite coordinate: every load is skipped
    pair[i].load(reinterpret_cast<const ET*>(fine_row), i == 0 ? pyr.width[0] : pyr.width[2],
                 valid ? x0 * (i == 0 ? 1.0f : 0.25f) : __int_as_float(0x7fc00000), i == 0 ? 1 : 0);
  }
  // ---- weights -> B operand (tf32, zero padded) while the gathers are in flight.  Row n of the
  // weights is K contiguous floats; thread (n = idx / KP4, piece) handles one 16-byte piece.  All
  // loads of a thread are issued before the first is used (cout = 64: 5 pieces per thread).
  {
    constexpr int kMaxPieces = (kFcMaxCout * KP4 + kFcPx - 1) / kFcPx;
    const int total = cout * KP4;
#pragma unroll 1
    for (int i0 = 0; i0 < kMaxPieces; i0 += 5) {
      float wv[5][4];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int idx = tid + (i0 + u) * kFcPx;
        const int n = idx / KP4, piece = idx - n * KP4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = piece * 4 + j;
          wv[u][j] = (idx < total && k < K) ? __ldg(weight + (long long)n * K + k) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int idx = tid + (i0 + u) * kFcPx;
        const int n = idx / KP4, piece = idx - n * KP4;
        if (idx < total)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sB + kmajor_off(n, piece * 4, cout)),
                       "r"(to_tf32(wv[u][0])), "r"(to_tf32(wv[u][1])), "r"(to_tf32(wv[u][2])),
                       "r"(to_tf32(wv[u][3]))
                       : "memory");
      }
      if (tid + (i0 + 5) * kFcPx >= total) break;
    }
  }
  for (int c = tid; c < cout; c += kFcPx) s_bias[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
  // ---- interpolate and write this pixel's operand row (channel = level*RD + tap), zero padded
  {
    float row_vals[KP4 * 4];
#pragma unroll
    for (int k = 0; k < KP4 * 4; ++k) row_vals[k] = 0.f;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float e[RD];
      if (2 * i + 1 < NL) {
        pair[i].template coarse<LT>((i == 0 || NL < 4) ? pyr.width[1] : pyr.width[3], e);
#pragma unroll
        for (int k = 0; k < RD; ++k) row_vals[(2 * i + 1) * RD + k] = e[k];
      }
      pair[i].template fine<LT>(e);
#pragma unroll
      for (int k = 0; k < RD; ++k) row_vals[2 * i * RD + k] = e[k];
    }
#pragma unroll
    for (int pc = 0; pc < KP4; ++pc)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sA + kmajor_off(tid, pc * 4, kFcPx)),
                   "r"(to_tf32(row_vals[4 * pc])), "r"(to_tf32(row_vals[4 * pc + 1])),
                   "r"(to_tf32(row_vals[4 * pc + 2])), "r"(to_tf32(row_vals[4 * pc + 3]))
                   : "memory");
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
      const uint64_t adesc = ptx::make_smem_desc(sA + atom * kFcPx * 128 + kin * 4, 16, 1024, ptx::kLayoutSwizzle128B);
      const uint64_t bdesc = ptx::make_smem_desc(sB + atom * cout * 128 + kin * 4, 16, 1024, ptx::kLayoutSwizzle128B);
      ptx::mma_tf32(tmem, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
    }
    ptx::tc_commit(ptx::smem_u32(&s_bar));
  }
  ptx::mbar_wait(ptx::smem_u32(&s_bar), 0);
  ptx::tc_fence_after_sync();
  // ---- epilogue: TMEM lane == pixel == thread; every store instruction is one 128-byte line
  CT* obase = out + (long long)b * cout * HW + hw;
  for (int c0 = 0; c0 < cout; c0 += 16) {
    uint32_t acc[16];
    tmem_ld_32x32_x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, acc);
    ptx::tmem_ld_wait();
    if (valid) {
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

bool pyramid_vec_ok(const smoe_pyramid_t* p);   // lookup.cu
bool pyramid_ld256_ok(const PyrDev& d, int NL);   // lookup.cu

template <typename VT, typename LT, typename CT>
static int launch_lookup_conv(const smoe_pyramid_t* pyr, const float* coords, long long cbs, float cs,
                              const float* weight, const float* bias, int cout, int relu, void* out,
                              cudaStream_t st, const CoordUpd& upd) {
  const PyrDev d = to_dev(pyr);
  const int HW = pyr->H * pyr->W1;
  const size_t smem = (size_t)kFcABytes + (size_t)cout * 256 + 1024;
  uint32_t cols = 32;
  while ((int)cols < cout) cols <<= 1;
  const bool big = (long long)pyr->B * HW * d.pitch[0] * (long long)sizeof(VT) > (96ll << 20);
  dim3 grid((HW + kFcPx - 1) / kFcPx, pyr->B);
  // Thread-per-pixel kernel everywhere (cfg1 8.98 -> 6.35 us vs the lane-group version, fp16
  // volumes 10.7 -> 6.4 us and 72 -> 57 us at B=8 96x312); fp32 volumes beyond L2 use its 256-bit
  // gather.  The lane-group version lives in the cross-check library (SMOE_CORR_LOOKUP_CONV=lanes).
  const bool missing = (pyr->levels_missing & ((1 << pyr->num_levels) - 1)) != 0;   // odd levels not built
  SMOE_REQUIRE((pyr->levels_missing & 0x5) == 0, SMOE_ERR_INVALID_ARG,
               "lookup_conv1x1: levels 0x%x of the pyramid were not built", pyr->levels_missing);
#ifdef SMOE_CROSSCHECK
  const bool lanes = tuning().conv_lanes != 0;
  SMOE_REQUIRE(!(missing && lanes), SMOE_ERR_INVALID_ARG,
               "lookup_conv1x1: the lane-group kernel needs all levels (levels_missing=0x%x)", pyr->levels_missing);
#else
  (void)missing;
#endif
  const bool ld256 = big && sizeof(VT) == 4 && pyramid_ld256_ok(d, pyr->num_levels) && tuning().px_ld256 != 0;
#define SMOE_FC_PX(NLV, STREAMV, LD)                                                              \
  do {                                                                                            \
    auto kern = lookup_conv_px_kernel<VT, LT, CT, NLV, STREAMV, LD>;                              \
    SMOE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    SMOE_CUDA_OK(launch_kernel_pdl(kern, grid, dim3(kFcPx), smem, st, d, coords, cbs, cs, weight, \
                                   bias, cout, relu, cols, static_cast<CT*>(out), HW, upd));      \
  } while (0)
#ifdef SMOE_CROSSCHECK
#define SMOE_FC_LANES(NLV, STREAMV)                                                               \
  if (lanes) {                                                                                    \
    auto kern = lookup_conv_kernel<VT, LT, CT, NLV, STREAMV>;                                     \
    SMOE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    SMOE_CUDA_OK(launch_kernel(kern, grid, dim3(kFcThreads), smem, st, d, coords, cbs, cs, weight, \
                               bias, cout, relu, cols, static_cast<CT*>(out), HW, upd));          \
    break;                                                                                        \
  }
#else
#define SMOE_FC_LANES(NLV, STREAMV)
#endif
#define SMOE_FC_LAUNCH(NLV, STREAMV)                                                              \
  do {                                                                                            \
    SMOE_FC_LANES(NLV, STREAMV)                                                                   \
    if constexpr (STREAMV && sizeof(VT) == 4) {                                                   \
      if (ld256) { SMOE_FC_PX(NLV, true, true); break; }                                          \
    }                                                                                             \
    SMOE_FC_PX(NLV, STREAMV, false);                                                              \
  } while (0)
#define SMOE_FC_LEVELS(STREAMV)                                                                   \
  switch (pyr->num_levels) {                                                                      \
    case 1: SMOE_FC_LAUNCH(1, STREAMV); break;                                                    \
    case 2: SMOE_FC_LAUNCH(2, STREAMV); break;                                                    \
    case 3: SMOE_FC_LAUNCH(3, STREAMV); break;                                                    \
    default: SMOE_FC_LAUNCH(4, STREAMV); break;                                                   \
  }
  if (big) { SMOE_FC_LEVELS(true) } else { SMOE_FC_LEVELS(false) }
#undef SMOE_FC_LEVELS
#undef SMOE_FC_LAUNCH
#undef SMOE_FC_LANES
#undef SMOE_FC_PX
  return check_launch("lookup_conv_kernel");
}

}  // namespace smoe

extern "C" int smoe_corr_lookup_conv1x1(const smoe_pyramid_t* pyr, const float* coords,
                                        int64_t coords_batch_stride, float coord_scale, int radius,
                                        int lookup_dtype, const float* weight, const float* bias,
                                        int cout, int relu, void* out, int out_dtype,
                                        smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(pyr, "lookup_conv1x1")) return rc;
  SMOE_REQUIRE(dtype_ok(lookup_dtype) && dtype_ok(out_dtype), SMOE_ERR_INVALID_ARG,
               "lookup_conv1x1: unknown dtype (lookup %d, out %d)", lookup_dtype, out_dtype);
  SMOE_REQUIRE(!(pyr->dtype == SMOE_DTYPE_F32 && lookup_dtype == SMOE_DTYPE_F16), SMOE_ERR_UNSUPPORTED,
               "lookup_conv1x1: fp16 lookup arithmetic on an fp32 pyramid is not part of the reference API");
  SMOE_REQUIRE(radius == 4, SMOE_ERR_UNSUPPORTED,
               "lookup_conv1x1: radius %d (the fused kernel is built for radius 4; run lookup + conv)", radius);
  SMOE_REQUIRE(cout >= 16 && cout <= kFcMaxCout && cout % 16 == 0, SMOE_ERR_UNSUPPORTED,
               "lookup_conv1x1: %d output channels (need a multiple of 16 in [16,%d])", cout, kFcMaxCout);
  SMOE_REQUIRE(pyramid_vec_ok(pyr), SMOE_ERR_UNSUPPORTED,
               "lookup_conv1x1: pyramid levels must be 16-byte aligned with 16-byte-multiple pitches");
  const long long HW = (long long)pyr->H * pyr->W1;
  if (pyr->B == 0 || HW == 0) return SMOE_OK;
  SMOE_REQUIRE(coords && out && weight, SMOE_ERR_INVALID_ARG, "lookup_conv1x1: NULL coords/weight/out pointer");
  SMOE_REQUIRE(coords_batch_stride >= HW, SMOE_ERR_INVALID_ARG,
               "lookup_conv1x1: coords batch stride %lld < H*W1 %lld", (long long)coords_batch_stride, HW);
  SMOE_REQUIRE(pyr->B <= 65535 && HW < (1ll << 31), SMOE_ERR_UNSUPPORTED, "lookup_conv1x1: batch/plane too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const CoordUpd none = {nullptr, nullptr, nullptr, nullptr, 0};
  const bool f16v = pyr->dtype == SMOE_DTYPE_F16, f16l = lookup_dtype == SMOE_DTYPE_F16,
             f16o = out_dtype == SMOE_DTYPE_F16;
#define SMOE_FC_GO(VT, LT)                                                                           \
  return f16o ? launch_lookup_conv<VT, LT, __half>(pyr, coords, coords_batch_stride, coord_scale,     \
                                                   weight, bias, cout, relu, out, st, none)           \
              : launch_lookup_conv<VT, LT, float>(pyr, coords, coords_batch_stride, coord_scale,      \
                                                  weight, bias, cout, relu, out, st, none)
  if (!f16v) { SMOE_FC_GO(float, float); }
  if (f16l) { SMOE_FC_GO(__half, __half); }
  SMOE_FC_GO(__half, float);
#undef SMOE_FC_GO
}
