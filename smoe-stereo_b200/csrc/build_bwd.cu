// build_bwd.cu -- backward of the correlation volume build on tcgen05 / TMA (sm_100a).
//
// With G = d loss / d level0 (fp32, fused-pyramid layout, already folded), per image row (b,h):
//   which = 0:  dF1^T [W1 x C] = G   [W1 x W2] * F2^T [W2 x C]   (M = w1, K = w2)
//   which = 1:  dF2^T [W2 x C] = G^T [W2 x W1] * F1^T [W1 x C]   (M = w2, K = w1)
// both scaled by 1/sqrt(C)  (autograd of core/corr.py:154-156).
//
// Operand layouts, all read in place by TMA:
//   G   as A of GEMM 0: K-major (w2 contiguous)          -> SWIZZLE_128B tiles, box {32 k, 128 m}
//   G^T as A of GEMM 1: MN-major (m = w2 contiguous)      -> 128B_ATOM_32B slabs, box {32 m, 32 k}
//   F^T as B (N = c, K = w): K-major (w contiguous in NCHW) -> SWIZZLE_128B, box {32 k, 1, 256 c, 1}
// The accumulator tile is [128 (w) x 256 (c)]: TMEM lane = w, column = c, so for a fixed channel
// the 32 lanes of a warp hold 32 consecutive w -- exactly the contiguous axis of the NCHW output:
// a warp's [32 c x 32 w] sub-tile is staged in shared memory with conflict-free stores (lane = w)
// and written by ONE TMA tensor store (box {32 w, 1, 32 c, 1} on the NCHW gradient, tails clipped).
// (A first version issued 32 scalar global stores per chunk with per-store 64-bit address math:
// ~740 instructions per chunk on one warp per scheduler, 3x slower -- profiles/r01/README.md.)
// Same warp roles / mbarrier pipeline as corr_build_kernel (build.cu).
#include <cuda.h>
#include <math.h>

#include "common.cuh"
#include "ptx.cuh"

namespace smoe {

constexpr int kBwdM = 128, kBwdN = 256, kBwdKT = 32, kBwdStages = 4;
constexpr int kBwdABytes = kBwdM * kBwdKT * 4;    // 16 KB
constexpr int kBwdBBytes = kBwdN * kBwdKT * 4;    // 32 KB
constexpr int kBwdStageBytes = kBwdABytes + kBwdBBytes;
constexpr int kBwdThreads = 192;
constexpr int kBwdStgBytes = 32 * 128;             // staging tile: 32 channels x 32 w (128 bytes)
constexpr size_t kBwdSmemBytes = (size_t)kBwdStages * kBwdStageBytes + 4 * 2 * kBwdStgBytes + 1024;

struct BwdDims {
  int W1, W2, C, H, BH;
  int MT1, MT2, NT;          // M tiles of GEMM 0 / GEMM 1, N tiles over C
  float scale;               // 1/sqrt(C)
};

struct BwdItem {
  int bh, which, mt, nt;
};

__device__ __forceinline__ BwdItem decode_item(int item, const BwdDims& d) {
  const int per_row = (d.MT1 + d.MT2) * d.NT;
  BwdItem it;
  it.bh = item / per_row;
  int r = item - it.bh * per_row;
  it.nt = r % d.NT;
  r /= d.NT;
  it.which = r >= d.MT1 ? 1 : 0;
  it.mt = it.which ? r - d.MT1 : r;
  return it;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
corr_build_bwd_kernel(const __grid_constant__ CUtensorMap map_gk,   // G, K-major tiles (GEMM 0 A)
                      const __grid_constant__ CUtensorMap map_gm,   // G, MN-major slabs (GEMM 1 A)
                      const __grid_constant__ CUtensorMap map_f1,   // fmap1 as B of GEMM 1
                      const __grid_constant__ CUtensorMap map_f2,   // fmap2 as B of GEMM 0
                      const __grid_constant__ CUtensorMap map_d1,   // d fmap1 (W1,H,C,B), store
                      const __grid_constant__ CUtensorMap map_d2,   // d fmap2 (W2,H,C,B), store
                      const BwdDims dims) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kBwdStages], s_empty[kBwdStages], s_tfull[2], s_tempty[2];
  __shared__ uint32_t s_tmem_base;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = dims.BH * (dims.MT1 + dims.MT2) * dims.NT;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_gk);
    ptx::prefetch_tensormap(&map_gm);
    ptx::prefetch_tensormap(&map_f1);
    ptx::prefetch_tensormap(&map_f2);
    ptx::prefetch_tensormap(&map_d1);
    ptx::prefetch_tensormap(&map_d2);
    for (int i = 0; i < kBwdStages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&s_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&s_tfull[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_tempty[i]), 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&s_tmem_base), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const BwdItem it = decode_item(item, dims);
        const int b = it.bh / dims.H, h = it.bh - b * dims.H;
        const int K = it.which ? dims.W1 : dims.W2;
        const int num_k = (K + kBwdKT - 1) / kBwdKT;
        const int n_valid = min(kBwdN, dims.C - it.nt * kBwdN);
        // B rows beyond C are zero-filled by TMA but still counted in the transaction bytes
        const uint32_t tx_bytes = (uint32_t)kBwdABytes + (uint32_t)kBwdN * kBwdKT * 4;
        (void)n_valid;
        for (int ks = 0; ks < num_k; ++ks) {
          ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1u);
          const uint32_t full = ptx::smem_u32(&s_full[stage]);
          ptx::mbar_expect_tx(full, tx_bytes);
          const uint32_t sa = smem_base + stage * kBwdStageBytes;
          const uint32_t sb = sa + kBwdABytes;
          if (it.which == 0) {
            // A[m = w1][k = w2]: one K-major tile {32 k, 128 m}
            ptx::tma_load_3d(sa, &map_gk, full, ks * kBwdKT, it.mt * kBwdM, it.bh);
            ptx::tma_load_4d(sb, &map_f2, full, ks * kBwdKT, h, it.nt * kBwdN, b);
          } else {
            // A[m = w2][k = w1]: four MN-major slabs {32 m, 32 k}
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ptx::tma_load_3d(sa + j * 4096, &map_gm, full, it.mt * kBwdM + j * 32, ks * kBwdKT, it.bh);
            ptx::tma_load_4d(sb, &map_f1, full, ks * kBwdKT, h, it.nt * kBwdN, b);
          }
          if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const BwdItem it = decode_item(item, dims);
        const int K = it.which ? dims.W1 : dims.W2;
        const int num_k = (K + kBwdKT - 1) / kBwdKT;
        const int n_valid = min(kBwdN, dims.C - it.nt * kBwdN);
        const uint32_t n_mma = (uint32_t)((n_valid + 15) & ~15);
        // c_format f32 | tf32 x tf32 | a_major = which (0: K, 1: MN) | b_major = K | N | M
        const uint32_t idesc = (1u << 4) | (ptx::kFmtTF32 << 7) | (ptx::kFmtTF32 << 10) |
                               ((uint32_t)it.which << 15) | ((n_mma >> 3) << 17) |
                               ((uint32_t)(kBwdM >> 4) << 24);
        ptx::mbar_wait(ptx::smem_u32(&s_tempty[buf]), bphase ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kBwdN;
        for (int ks = 0; ks < num_k; ++ks) {
          ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = smem_base + stage * kBwdStageBytes;
          const uint32_t sb = sa + kBwdABytes;
#pragma unroll
          for (int kk = 0; kk < kBwdKT / 8; ++kk) {
            // K-major SW128: 8-row atoms of 1 KiB (SBO), advance 32 bytes per UMMA_K inside the row.
            // MN-major 128B_BASE32B: slabs 4 KiB apart (LBO), 4-k-row atoms (SBO 512), 1 KiB per UMMA_K.
            const uint64_t adesc =
                it.which == 0
                    ? ptx::make_smem_desc(sa + kk * 32, 16, 1024, ptx::kLayoutSwizzle128B)
                    : ptx::make_smem_desc(sa + kk * 1024, 4096, 512, ptx::kLayoutSwizzle128B_Base32B);
            const uint64_t bdesc = ptx::make_smem_desc(sb + kk * 32, 16, 1024, ptx::kLayoutSwizzle128B);
            ptx::mma_tf32(d_tmem, adesc, bdesc, idesc, (ks | kk) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(ptx::smem_u32(&s_empty[stage]));
          if (++stage == kBwdStages) { stage = 0; phase ^= 1u; }
        }
        ptx::tc_commit(ptx::smem_u32(&s_tfull[buf]));
        buf ^= 1;
        if (buf == 0) bphase ^= 1u;
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    const int quad = warp & 3;
    const uint32_t stg_base = smem_base + kBwdStages * kBwdStageBytes + (uint32_t)(warp - 2) * 2 * kBwdStgBytes;
    int buf = 0;
    uint32_t bphase = 0;
    int stg_sel = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const BwdItem it = decode_item(item, dims);
      const int b = it.bh / dims.H, h = it.bh - b * dims.H;
      const int Wm = it.which ? dims.W2 : dims.W1;
      const int m0 = it.mt * kBwdM + quad * 32;          // first w of this warp's 32 rows
      const int n_valid = min(kBwdN, dims.C - it.nt * kBwdN);
      const CUtensorMap* omap = it.which ? &map_d2 : &map_d1;
      ptx::mbar_wait(ptx::smem_u32(&s_tfull[buf]), bphase);
      ptx::tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * kBwdN;
      for (int c0 = 0; c0 < n_valid; c0 += 32) {
        const uint32_t stg = stg_base + (uint32_t)stg_sel * kBwdStgBytes;
        if (lane == 0) ptx::tma_store_wait_read<1>();     // staging buffer of two chunks ago drained
        __syncwarp();
        uint32_t acc[32];
        ptx::tmem_ld_32x32(taddr + c0, acc);
        ptx::tmem_ld_wait();
        // staging tile [32 c][32 w]: lane = w, so each store instruction covers one 128-byte row
#pragma unroll
        for (int j = 0; j < 32; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + j * 128 + lane * 4),
                       "f"(__uint_as_float(acc[j]) * dims.scale)
                       : "memory");
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (m0 < Wm) ptx::tma_store_4d(omap, stg, m0, h, it.nt * kBwdN + c0, b);
          ptx::tma_store_commit();
        }
        stg_sel ^= 1;
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(ptx::smem_u32(&s_tempty[buf]));
      buf ^= 1;
      if (buf == 0) bphase ^= 1u;
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder();   // build.cu

static int encode(CUtensorMap* out, int rank, const void* base, const cuuint64_t* gdim,
                  const cuuint64_t* gstr, const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what,
                  bool tf32 = true) {
  EncodeTiledFn enc = tensor_map_encoder();
  SMOE_REQUIRE(enc != nullptr, SMOE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = enc(out, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         (cuuint32_t)rank, const_cast<void*>(base),
                         gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SMOE_REQUIRE(r == CUDA_SUCCESS, SMOE_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what,
               (int)r);
  return SMOE_OK;
}

}  // namespace smoe

extern "C" {

int smoe_corr_build_bwd(const void* fmap1, const void* fmap2, int in_dtype, int B, int C, int H, int W1,
                        int W2, int64_t row_pitch1, int64_t row_pitch2, const smoe_pyramid_t* grad_pyr,
                        float denom, void* dfmap1, void* dfmap2, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(grad_pyr, "build_bwd")) return rc;
  SMOE_REQUIRE(in_dtype == SMOE_DTYPE_F32 && grad_pyr->dtype == SMOE_DTYPE_F32, SMOE_ERR_UNSUPPORTED,
               "build_bwd: feature maps and gradient pyramid must be fp32");
  SMOE_REQUIRE(B >= 0 && C >= 1 && H >= 0 && W1 >= 0 && W2 >= 0, SMOE_ERR_INVALID_ARG,
               "build_bwd: bad dims B=%d C=%d H=%d W1=%d W2=%d", B, C, H, W1, W2);
  SMOE_REQUIRE(grad_pyr->B == B && grad_pyr->H == H && grad_pyr->W1 == W1 && grad_pyr->width[0] == W2,
               SMOE_ERR_INVALID_ARG, "build_bwd: gradient pyramid dims do not match the feature maps");
  SMOE_REQUIRE(denom > 0.f, SMOE_ERR_INVALID_ARG, "build_bwd: denom must be positive");
  if (B == 0 || H == 0 || W1 == 0 || W2 == 0) return SMOE_OK;
  SMOE_REQUIRE(fmap1 && fmap2 && dfmap1 && dfmap2, SMOE_ERR_INVALID_ARG, "build_bwd: NULL pointer");
  SMOE_REQUIRE(row_pitch1 >= W1 && row_pitch2 >= W2 && row_pitch1 % 4 == 0 && row_pitch2 % 4 == 0,
               SMOE_ERR_INVALID_ARG,
               "build_bwd: row pitches %lld / %lld must be multiples of 4 elements (TMA 16-byte strides) "
               "and >= W1=%d / W2=%d", (long long)row_pitch1, (long long)row_pitch2, W1, W2);
  SMOE_REQUIRE(aligned16(fmap1) && aligned16(fmap2) && aligned16(grad_pyr->level_ptr[0]) &&
                   (grad_pyr->pitch[0] * 4) % 16 == 0,
               SMOE_ERR_INVALID_ARG, "build_bwd: operands must be 16-byte aligned");

  CUtensorMap gk, gm, f1m, f2m;
  {
    const cuuint64_t gdim[3] = {(cuuint64_t)W2, (cuuint64_t)W1, (cuuint64_t)B * H};
    const cuuint64_t gstr[2] = {(cuuint64_t)grad_pyr->pitch[0] * 4, (cuuint64_t)W1 * grad_pyr->pitch[0] * 4};
    const cuuint32_t box_k[3] = {32u, 128u, 1u};
    const cuuint32_t box_m[3] = {32u, 32u, 1u};
    if (int rc = encode(&gk, 3, grad_pyr->level_ptr[0], gdim, gstr, box_k, CU_TENSOR_MAP_SWIZZLE_128B, "G/K"))
      return rc;
    if (int rc = encode(&gm, 3, grad_pyr->level_ptr[0], gdim, gstr, box_m,
                        CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, "G/MN"))
      return rc;
  }
  auto fmap_map = [&](CUtensorMap* m, const void* base, int W, int64_t Wp, const char* what) {
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)Wp * 4, (cuuint64_t)H * Wp * 4, (cuuint64_t)C * H * Wp * 4};
    const cuuint32_t box[4] = {32u, 1u, (cuuint32_t)kBwdN, 1u};
    return encode(m, 4, base, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
  };
  if (int rc = fmap_map(&f1m, fmap1, W1, row_pitch1, "fmap1")) return rc;
  if (int rc = fmap_map(&f2m, fmap2, W2, row_pitch2, "fmap2")) return rc;
  CUtensorMap d1m, d2m;
  auto out_map = [&](CUtensorMap* m, void* base, int W, int64_t Wp, const char* what) {
    const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)Wp * 4, (cuuint64_t)H * Wp * 4, (cuuint64_t)C * H * Wp * 4};
    const cuuint32_t box[4] = {32u, 1u, 32u, 1u};
    return encode(m, 4, base, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_NONE, what, false);
  };
  SMOE_REQUIRE(aligned16(dfmap1) && aligned16(dfmap2), SMOE_ERR_INVALID_ARG,
               "build_bwd: gradient outputs must be 16-byte aligned");
  if (int rc = out_map(&d1m, dfmap1, W1, row_pitch1, "dfmap1")) return rc;
  if (int rc = out_map(&d2m, dfmap2, W2, row_pitch2, "dfmap2")) return rc;

  BwdDims d;
  d.W1 = W1; d.W2 = W2; d.C = C; d.H = H; d.BH = B * H;
  d.MT1 = (W1 + kBwdM - 1) / kBwdM;
  d.MT2 = (W2 + kBwdM - 1) / kBwdM;
  d.NT = (C + kBwdN - 1) / kBwdN;
  d.scale = 1.0f / denom;
  const long long items = (long long)d.BH * (d.MT1 + d.MT2) * d.NT;
  SMOE_REQUIRE(items < (1ll << 31), SMOE_ERR_UNSUPPORTED, "build_bwd: too many work items");

  SMOE_CUDA_OK(cudaFuncSetAttribute(corr_build_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)kBwdSmemBytes));
  int dev = 0, sms = 0;
  SMOE_CUDA_OK(cudaGetDevice(&dev));
  SMOE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = (int)(items < sms ? items : sms);
  SMOE_CUDA_OK(launch_kernel(corr_build_bwd_kernel, dim3(grid), dim3(kBwdThreads), kBwdSmemBytes,
                             static_cast<cudaStream_t>(stream), gk, gm, f1m, f2m, d1m, d2m, d));
  return check_launch("corr_build_bwd_kernel");
}

}  // extern "C"
