// build.cu -- all-pairs 1-D correlation volume + fused W-axis pyramid on tcgen05 / TMA (sm_100a).
//
//   level0[(b,h,w1), w2] = (sum_c f1[b,c,h,w1] * f2[b,c,h,w2]) / sqrt(C)       (core/corr.py:148-156)
//   level(i+1)[., j]     = (level i[., 2j] + level i[., 2j+1]) / 2            (core/corr.py:122-125)
//
// For a fixed (b,h) this is a [W1 x C] * [C x W2] GEMM whose operands are both "MN-major": in NCHW
// memory the W axis is contiguous and the contraction axis C has stride H*W.  Work item =
// (b*H+h, 128-row M tile, <=256-column N tile); a persistent CTA walks items round-robin so
// that the M/N tiles of one image row run concurrently and share its operands through L2.
//
// Warp roles (192 threads):
//   warp 0      TMA producer: per K step loads 128B-wide W slabs x KT channels of both operands
//               with 4-D tensor maps (W,H,C,B) straight from NCHW memory into the swizzled
//               MN-major layout tcgen05 expects (no transpose pass, TMA zero-fills W/C tails).
//   warp 1      allocates TMEM (2 x 256 fp32 columns) and issues tcgen05.mma
//               (kind::tf32 for fp32 inputs, kind::f16 for fp16 inputs), fp32 accumulate.
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time, divide, build the three pooled levels
//               in registers (pooling runs along N == along a thread's own TMEM row), transpose
//               through swizzled shared-memory tiles and write whole 128-byte row segments with
//               coalesced 16-byte stores; the narrow rows of the coarse levels are gathered over
//               several chunks first (GroupPlan).  Accumulators are double buffered so this
//               overlaps the next item's loads and MMAs.
// A CTA-pair variant (kPair, tcgen05 cta_group::2) is selectable; see the kernel comment.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace smoe {

// ------------------------------------------------------------------------------ tile constants
constexpr int kBlockM = 128;          // UMMA M (TMEM lanes)
constexpr int kBlockN = 256;          // max UMMA N / accumulator columns per buffer
// Probe build (SMOE_NVCC_DEFS=-DSMOE_BUILD_PROBES python smoe-stereo_b200/build.py --force): the
// environment variable SMOE_CORR_BUILD_DEBUG then switches off parts of the kernel (1: operand
// loads, 2: result stores, 4: MMAs) to time the rest alone.  Results are garbage; never shipped.
#ifdef SMOE_BUILD_PROBES
constexpr bool kProbes = true;
#else
constexpr bool kProbes = false;
#endif
constexpr int kMaxStages = 8;              // ring depth is chosen per launch (BuildDims::stages)
constexpr int kEpiWarps = 4;                // one per TMEM lane quadrant
constexpr int kStageABytes = 16 * 1024;   // 128 rows x KT channels
constexpr int kStageBBytes = 32 * 1024;   // 256 cols x KT channels at most; a launch whose N tiles
                                          // are narrower packs its stages tighter (stage_bytes)
constexpr int kMaxSmemBytes = 226 * 1024; // dynamic shared memory budget (227 KiB opt-in limit on
                                          // sm_100 minus the kernel's static barriers)
constexpr int kBuildThreads = 64 + 32 * kEpiWarps;
constexpr int kTmemCols = 512;
// epilogue staging: per warp one 4 KiB tile per pyramid level (32 rows x up to 128 bytes).
// Level 0 leaves as 128-byte row segments per chunk.  A coarser level i only produces 128>>i
// bytes per row and chunk; rows that narrow are expensive to store (a 16-byte row is half a DRAM
// sector: stored per chunk, level 3 alone cost as much as level 0 -- profiles/r01/README.md),
// so the chunks of a level are gathered over up to 2^i chunks and leave as one wider segment
// (see GroupPlan).
constexpr int kStgLevelBytes = 4096;
constexpr int kStgBytes = SMOE_MAX_LEVELS * kStgLevelBytes;
constexpr int kStgPerWarp = kStgBytes;
constexpr int kBuildSmemFixed = kEpiWarps * kStgPerWarp + 1024;      // staging + alignment slack

template <typename T> struct OperandCfg;
template <> struct OperandCfg<float> {
  static constexpr int kSlab = 32;          // W elements per 128-byte slab
  static constexpr int kKT = 32;            // channels per pipeline stage
  static constexpr int kUmmaK = 8;          // K per tcgen05.mma (32 bytes)
  static constexpr uint32_t kSBO = 512;     // 4 k-rows per swizzle atom
  static constexpr uint64_t kLayout = ptx::kLayoutSwizzle128B_Base32B;
  static constexpr uint32_t kFmt = ptx::kFmtTF32;
};
template <> struct OperandCfg<__half> {
  static constexpr int kSlab = 64;
  static constexpr int kKT = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kSBO = 1024;    // 8 k-rows per swizzle atom
  static constexpr uint64_t kLayout = ptx::kLayoutSwizzle128B;
  static constexpr uint32_t kFmt = ptx::kFmtF16;
};

struct BuildDims {
  int W1, W2, C, H, BH;      // BH = B*H image rows
  int MT, NT;                // tiles per image row
  int BN;                    // columns per N tile: W2 split evenly, rounded up to 32, <= kBlockN
  float denom;               // level0 = sum / denom   (core/corr.py:156: corr / sqrt(D))
  float rcp;                 // 1/denom, used when it is exact (denom a power of two)
  int rcp_exact;
  int stages;                // operand ring: depth and bytes per stage (A 16 KiB + the B slabs
  int stage_bytes;           // one N tile needs), sized to fill shared memory
  int pair;                  // CTA-pair kernel (MT counts 256-row tiles)
  int kmajor;                // channels-last feature maps (SMOE_BUILD_CHANNELS_LAST)
  int store_mask;            // bit i: level i is written (SMOE_BUILD_SKIP_ODD_LEVELS clears 1 and 3)
  int debug;                 // probe builds only (kProbes)
};

// Division exactly as the reference performs it: a true fp32 division, replaced by the (bit-
// identical) multiplication when the denominator is a power of two (C = 64, 256, ...).
struct Scaler {
  float denom, rcp;
  int exact;
  // Whole-chunk form: ONE uniform branch, then 32 independent operations.  (A per-element
  // `exact ? mul : div` compiles to 32 serialized division regions; with a single epilogue warp
  // per scheduler that cost ~6 us per 128x240 tile, profiles/r01/README.md.)
  __device__ __forceinline__ void apply(const uint32_t (&acc)[32], float (&v)[32]) const {
    if (exact) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * rcp;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __fdiv_rn(__uint_as_float(acc[j]), denom);
    }
  }
  __device__ __forceinline__ void apply_rounded_half(const uint32_t (&acc)[32], float (&v)[32]) const {
    if (exact) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = __half2float(__float2half_rn(__half2float(__float2half_rn(__uint_as_float(acc[j]))) * rcp));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = __half2float(__float2half_rn(
            __fdiv_rn(__half2float(__float2half_rn(__uint_as_float(acc[j]))), denom)));
    }
  }
};

// ------------------------------------------------------------------------------------ epilogue
// Staging tile of width class d for one warp: 32 rows x (128 >> d) bytes, dense, with an XOR
// swizzle of the 16-byte piece index (the 128B / 64B / 32B patterns TMA uses) so that the 32
// lanes (one row each) write without bank conflicts.
__device__ __forceinline__ uint32_t stg_addr(uint32_t tile, int d, int row, int byte_off) {
  const uint32_t lin = (uint32_t)row * (128u >> d) + (uint32_t)byte_off;
  const uint32_t mask = (8u >> d) - 1u;
  return tile + (lin ^ (((lin >> 7) & mask) << 4));
}
// Where chunk c of a tile with `chunks` chunks puts its level-i output.  Chunks are gathered into
// aligned power-of-two groups of at most 2^i chunks that lie inside the tile (buddy
// decomposition: 5 chunks at level 3 -> groups of 4 + 1); a group of 2^k chunks is stored as one
// tile of width class d = i - k (row = 128 >> d bytes) once its last chunk has been staged.
struct GroupPlan {
  int d[SMOE_MAX_LEVELS];       // width class of the group's staging tile
  int off[SMOE_MAX_LEVELS];     // byte offset of this chunk inside a staged row
  int first[SMOE_MAX_LEVELS];   // first chunk of the group
  bool flush[SMOE_MAX_LEVELS];  // this chunk completes the group
  __device__ __forceinline__ GroupPlan(int c, int chunks) {
#pragma unroll
    for (int i = 0; i < SMOE_MAX_LEVELS; ++i) {
      int k = i;
      while (k > 0 && ((c >> k) << k) + (1 << k) > chunks) --k;
      const int s = (c >> k) << k;
      d[i] = i - k;
      first[i] = s;
      off[i] = (c - s) * (128 >> i);
      flush[i] = (c - s) == (1 << k) - 1;
    }
  }
};
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                              uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void st_shared_v2u(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One thread owns one output row and 32 consecutive level-0 columns (sub-chunk `sub` of the
// 128-byte-wide staging chunk).  Produces 32/16/8/4 values of levels 0..3.
template <typename OT> struct ChunkStage;

template <> struct ChunkStage<float> {
  static constexpr int kSubChunks = 1;     // 32 fp32 columns == 128 bytes
  // store_mask: bit i set = level i is staged and stored (SMOE_BUILD_SKIP_ODD_LEVELS clears 1 and 3;
  // a skipped level is still computed when a finer stored level is pooled from it)
  template <bool>
  __device__ static void run(const uint32_t (&acc)[32], uint32_t stg, int row, int /*sub*/,
                             const Scaler& sc, int num_levels, const GroupPlan& gp, int store_mask) {
    float v[32];
    sc.apply(acc, v);
#pragma unroll
    for (int g = 0; g < 8; ++g)
      st_shared_v4(stg_addr(stg, 0, row, 16 * g), v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    if (num_levels < 2 || (store_mask >> 1) == 0) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
    if (store_mask & 2) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        st_shared_v4(stg_addr(stg + kStgLevelBytes, gp.d[1], row, gp.off[1] + 16 * g), v[4 * g], v[4 * g + 1],
                     v[4 * g + 2], v[4 * g + 3]);
    }
    if (num_levels < 3 || (store_mask >> 2) == 0) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
    if (store_mask & 4) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
        st_shared_v4(stg_addr(stg + 2 * kStgLevelBytes, gp.d[2], row, gp.off[2] + 16 * g), v[4 * g],
                     v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
    if (num_levels < 4 || (store_mask & 8) == 0) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
    st_shared_v4(stg_addr(stg + 3 * kStgLevelBytes, gp.d[3], row, gp.off[3]), v[0], v[1], v[2], v[3]);
  }
};

// fp16 pyramid: every level is rounded to half and the next level is pooled from the ROUNDED
// values with an fp32 sum, which is what F.avg_pool2d does on a half tensor (core/corr.py:42).
// 64 fp16 columns == 128 bytes, processed as two 32-column sub-chunks.  Everything stays packed:
// one F2FP per value pair, the power-of-two division as an exact HMUL2, pooling as
// unpack -> fp32 (a+b)*0.5 -> pack (one rounding, like avg_pool2d's fp32 accumulator).
template <> struct ChunkStage<__half> {
  static constexpr int kSubChunks = 2;
  __device__ static __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
  template <int N>   // N pooled half2 outputs from 2N half2 inputs
  __device__ static __forceinline__ void pool(const __half2* in, __half2* out) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float2 a = __half22float2(in[2 * j]), b = __half22float2(in[2 * j + 1]);
      out[j] = __floats2half2_rn((a.x + a.y) * 0.5f, (b.x + b.y) * 0.5f);
    }
  }
  // kEinsumIsHalf: fp16 feature maps -- the reference's einsum output is already fp16 and the
  // division rounds a second time (reg_cuda path).  Otherwise (fp32 feature maps, fp16 STORAGE
  // opted into by the caller) the fp32 quotient is rounded once.
  template <bool kEinsumIsHalf>
  __device__ static void run(const uint32_t (&acc)[32], uint32_t stg, int row, int sub,
                             const Scaler& sc, int num_levels, const GroupPlan& gp, int store_mask) {
    __half2 h0[16];
    if (!kEinsumIsHalf) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float a = __uint_as_float(acc[2 * j]), b = __uint_as_float(acc[2 * j + 1]);
        h0[j] = sc.exact ? __floats2half2_rn(a * sc.rcp, b * sc.rcp)
                         : __floats2half2_rn(__fdiv_rn(a, sc.denom), __fdiv_rn(b, sc.denom));
      }
    } else if (sc.exact) {
      const __half2 r2 = __float2half2_rn(sc.rcp);                 // power of two: exact
#pragma unroll
      for (int j = 0; j < 16; ++j)
        h0[j] = __hmul2_rn(__floats2half2_rn(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1])), r2);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 f = __half22float2(
            __floats2half2_rn(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1])));
        h0[j] = __floats2half2_rn(__fdiv_rn(f.x, sc.denom), __fdiv_rn(f.y, sc.denom));
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      st_shared_v4u(stg_addr(stg, 0, row, 64 * sub + 16 * g), as_u32(h0[4 * g]), as_u32(h0[4 * g + 1]),
                    as_u32(h0[4 * g + 2]), as_u32(h0[4 * g + 3]));
    if (num_levels < 2 || (store_mask >> 1) == 0) return;
    __half2 h1[8];
    pool<8>(h0, h1);
    if (store_mask & 2) {
#pragma unroll
      for (int g = 0; g < 2; ++g)
        st_shared_v4u(stg_addr(stg + kStgLevelBytes, gp.d[1], row, gp.off[1] + 32 * sub + 16 * g), as_u32(h1[4 * g]), as_u32(h1[4 * g + 1]),
                      as_u32(h1[4 * g + 2]), as_u32(h1[4 * g + 3]));
    }
    if (num_levels < 3 || (store_mask >> 2) == 0) return;
    __half2 h2[4];
    pool<4>(h1, h2);
    if (store_mask & 4)
      st_shared_v4u(stg_addr(stg + 2 * kStgLevelBytes, gp.d[2], row, gp.off[2] + 16 * sub), as_u32(h2[0]), as_u32(h2[1]), as_u32(h2[2]),
                    as_u32(h2[3]));
    if (num_levels < 4 || (store_mask & 8) == 0) return;
    __half2 h3[2];
    pool<2>(h2, h3);
    st_shared_v2u(stg_addr(stg + 3 * kStgLevelBytes, gp.d[3], row, gp.off[3] + 8 * sub), as_u32(h3[0]), as_u32(h3[1]));
  }
};

__device__ __forceinline__ uint4 ld_shared_v4u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// Copy one staged tile (width class d: 32 rows x (128 >> d) bytes) to global memory with plain
// coalesced 16-byte stores: 8 >> d lanes cover one row, so every warp instruction writes
// 4 << d whole row segments.  `grow0` points at (first row of the warp, column 0) of the level,
// col0 (elements, a multiple of 16 bytes) is where the tile starts.  Columns at or beyond the
// level width are dropped per 16-byte piece; the pitch rule of smoe_corr_build guarantees that a
// piece which starts inside the row ends inside its padding.
template <typename OT>
__device__ __forceinline__ void store_staged_tile(uint32_t tile, int d, OT* __restrict__ grow0,
                                                  long long pitch, int col0, int width, int rows_valid,
                                                  int lane) {
  constexpr int kPer16 = 16 / (int)sizeof(OT);
  const int ppr = 8 >> d;                       // 16-byte pieces per staged row
  const int piece = lane & (ppr - 1);
  const int rsub = lane >> (3 - d);             // row inside one pass
  const int col = col0 + piece * kPer16;
  const bool col_ok = col < width;
  uint4 v[8];
#pragma unroll
  for (int p = 0; p < 8; ++p)
    if (p < ppr) v[p] = ld_shared_v4u(stg_addr(tile, d, p * (4 << d) + rsub, piece * 16));
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int row = p * (4 << d) + rsub;
    if (p < ppr && col_ok && row < rows_valid)
      *reinterpret_cast<uint4*>(grow0 + (long long)row * pitch + col) = v[p];
  }
}

// -------------------------------------------------------------------------------------- kernel
// kPair: the kernel runs as clusters of two CTAs (one TPC) that compute one 256-row M tile with
// tcgen05.mma.cta_group::2.  Each CTA loads its own 128 rows of A and only HALF of the B columns
// -- the tensor cores read both halves through the pair's shared memory -- so a B row slab is
// fetched from L2 once per 256 output rows instead of once per 128.  Rank 0 (the leader) owns
// the `full` and `accumulator empty` barriers and issues the MMAs; both CTAs run their own
// producer and epilogue; `empty` / `accumulator full` arrive in both CTAs by multicast commit.
//
// kKMajor: the feature maps are channels-last, (B,H,W,C) in memory, so both operands are K-major
// (the contraction axis C is contiguous): the canonical UMMA layout.  One TMA box {128 bytes of C,
// rows of W} per operand and stage lands as 8-row SWIZZLE_128B atoms; there is no 16-byte rule on
// W any more (the W stride is C*sizeof(T)), so odd widths need no pitch-padding pre-pass
// (the layout a channels-last producer convolution emits).
template <typename IT, typename OT, bool kPair, bool kKMajor>
__global__ void __launch_bounds__(kBuildThreads, 1)
corr_build_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const PyrDev out, const int num_levels, const BuildDims dims,
                  long long* __restrict__ timeline) {
  using Cfg = OperandCfg<IT>;
  constexpr int kSlabBytes = Cfg::kKT * 128;                 // one W slab x KT channels
  constexpr int kSlabsA = kBlockM / Cfg::kSlab;
  constexpr int kKSteps = Cfg::kKT / Cfg::kUmmaK;            // MMAs per stage
  constexpr uint32_t kKStepBytes = Cfg::kUmmaK * 128;        // smem advance per MMA along K
  static_assert(kSlabsA * kSlabBytes == kStageABytes, "A stage size");
  static_assert((kBlockN / Cfg::kSlab) * kSlabBytes == kStageBBytes, "B stage size");
  static_assert(!(kPair && kKMajor), "the CTA-pair kernel exists for NCHW operands only");
  static_assert(Cfg::kKT * sizeof(IT) == 128 && Cfg::kUmmaK * sizeof(IT) == 32, "K-major row / K step");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_tfull[2], s_tempty[2];
  __shared__ uint32_t s_tmem_base;

  // debug timeline (smoe_corr_debug_set_timeline): clock64 stamps, 16 slots per role per CTA
  const long long t_origin = clock64();
  auto stamp = [&](int role, int slot) {
    if (timeline != nullptr && slot < 16)
      timeline[((long long)blockIdx.x * 3 + role) * 16 + slot] = clock64() - t_origin;
  };
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzle atoms: 1 KiB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = dims.BH * dims.MT * dims.NT;
  const int num_k = (dims.C + Cfg::kKT - 1) / Cfg::kKT;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;             // CTA within the pair
  const int worker = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // item walker (CTA or pair)
  const int workers = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    for (int i = 0; i < dims.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(&s_full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&s_tfull[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&s_tempty[i]), kEpiWarps * (kPair ? 2 : 1));   // lane 0 of each warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if (kPair) {
      ptx::tmem_alloc_pair(ptx::smem_u32(&s_tmem_base), kTmemCols);
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(ptx::smem_u32(&s_tmem_base), kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before_sync();
  if (kPair) ptx::cluster_sync(); else __syncthreads();   // peer barriers are initialised too
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the
  // tail of the previous kernel in the stream; global memory is touched only after the wait.
  // This kernel does NOT trigger its own dependents early: a lookup launched with the PDL attribute
  // that became resident while the build was still running cost 7.4 us (tools/step_probe.py: build +
  // first lookup 42.6 us vs 35.2 us); without the trigger it launches when the build has finished.
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      stamp(0, 0);
      for (int item = worker; item < total_items; item += workers, ++it) {
        const int nt = item % dims.NT;
        const int mt = (item / dims.NT) % dims.MT;
        const int bh = item / (dims.NT * dims.MT);
        const int b = bh / dims.H, h = bh - b * dims.H;
        const int n_valid = min(dims.BN, dims.W2 - nt * dims.BN);
        // pair: this CTA's half of the N tile (whole slabs), leader expects both CTAs' bytes
        const int slabs_b = kPair ? (n_valid + 2 * Cfg::kSlab - 1) / (2 * Cfg::kSlab)
                                  : (n_valid + Cfg::kSlab - 1) / Cfg::kSlab;
        // K-major: one box of kBlockM rows (A) and one of BN rows (B), 128 bytes of C each; rows
        // beyond W are zero-filled by TMA but count in the transaction bytes
        const uint32_t tx_bytes = kKMajor ? (uint32_t)(kBlockM + dims.BN) * 128u
                                          : (uint32_t)(kSlabsA + slabs_b) * kSlabBytes * (kPair ? 2u : 1u);
        const int a_row0 = (kPair ? mt * 2 + (int)rank : mt) * kBlockM;
        const int b_col0 = nt * dims.BN + (kPair ? (int)rank * slabs_b * Cfg::kSlab : 0);
        for (int ks = 0; ks < num_k; ++ks) {
          ptx::mbar_wait(ptx::smem_u32(&s_empty[stage]), phase ^ 1u);
          const uint32_t full_local = ptx::smem_u32(&s_full[stage]);
          if (kProbes && (dims.debug & 1)) {
            if (rank == 0) ptx::mbar_arrive(full_local);
            if (++stage == dims.stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          if (rank == 0) ptx::mbar_expect_tx(full_local, tx_bytes);
          const uint32_t full = kPair ? ptx::map_to_cta(full_local, 0) : full_local;
          const uint32_t sa = smem_base + stage * dims.stage_bytes;
          const uint32_t sb = sa + kStageABytes;
          if (kKMajor) {
            ptx::tma_load_4d(sa, &tmap_a, full, ks * Cfg::kKT, a_row0, h, b);
            ptx::tma_load_4d(sb, &tmap_b, full, ks * Cfg::kKT, b_col0, h, b);
          } else if (kPair) {
#pragma unroll
            for (int j = 0; j < kSlabsA; ++j)
              ptx::tma_load_4d_pair(sa + j * kSlabBytes, &tmap_a, full, a_row0 + j * Cfg::kSlab, h,
                                    ks * Cfg::kKT, b);
            for (int j = 0; j < slabs_b; ++j)
              ptx::tma_load_4d_pair(sb + j * kSlabBytes, &tmap_b, full, b_col0 + j * Cfg::kSlab, h,
                                    ks * Cfg::kKT, b);
          } else {
#pragma unroll
            for (int j = 0; j < kSlabsA; ++j)
              ptx::tma_load_4d(sa + j * kSlabBytes, &tmap_a, full, a_row0 + j * Cfg::kSlab, h,
                               ks * Cfg::kKT, b);
            for (int j = 0; j < slabs_b; ++j)
              ptx::tma_load_4d(sb + j * kSlabBytes, &tmap_b, full, b_col0 + j * Cfg::kSlab, h,
                               ks * Cfg::kKT, b);
          }
          if (++stage == dims.stages) { stage = 0; phase ^= 1u; }
        }
        stamp(0, 1 + it);                 // all loads of item `it` issued
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int buf = 0;
      uint32_t bphase = 0;
      int it = 0;
      for (int item = worker; item < total_items; item += workers, ++it) {
        const int nt = item % dims.NT;
        const int n_valid = min(dims.BN, dims.W2 - nt * dims.BN);
        const uint32_t n_mma =
            kPair ? (uint32_t)((n_valid + 2 * Cfg::kSlab - 1) / (2 * Cfg::kSlab) * (2 * Cfg::kSlab))
                  : (uint32_t)((n_valid + 15) & ~15);
        const uint32_t idesc = kKMajor ? ptx::make_idesc_k_major(Cfg::kFmt, kBlockM, n_mma)
                                       : ptx::make_idesc_mn_major(Cfg::kFmt, kPair ? 2 * kBlockM : kBlockM, n_mma);
        if (kPair) ptx::mbar_wait_cluster(ptx::smem_u32(&s_tempty[buf]), bphase ^ 1u);
        else ptx::mbar_wait(ptx::smem_u32(&s_tempty[buf]), bphase ^ 1u);
        ptx::tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)buf * kBlockN;
        for (int ks = 0; ks < num_k; ++ks) {
          ptx::mbar_wait(ptx::smem_u32(&s_full[stage]), phase);
          ptx::tc_fence_after_sync();
          if (ks == 0) stamp(1, 3 * it);            // first stage of the item landed
          if (ks == num_k - 1) stamp(1, 3 * it + 1);  // last stage landed
          const uint32_t sa = smem_base + stage * dims.stage_bytes;
          const uint32_t sb = sa + kStageABytes;
#pragma unroll
          for (int kk = 0; kk < kKSteps; ++kk) {
            if (kProbes && (dims.debug & 4)) break;
            // MN-major: slabs kSlabBytes apart (LBO), one UMMA_K = kKStepBytes further down.
            // K-major SWIZZLE_128B: 8-row atoms of 1 KiB (SBO), 32 bytes per UMMA_K inside the row.
            const uint64_t adesc =
                kKMajor ? ptx::make_smem_desc(sa + kk * 32, 16, 1024, ptx::kLayoutSwizzle128B)
                        : ptx::make_smem_desc(sa + kk * kKStepBytes, kSlabBytes, Cfg::kSBO, Cfg::kLayout);
            const uint64_t bdesc =
                kKMajor ? ptx::make_smem_desc(sb + kk * 32, 16, 1024, ptx::kLayoutSwizzle128B)
                        : ptx::make_smem_desc(sb + kk * kKStepBytes, kSlabBytes, Cfg::kSBO, Cfg::kLayout);
            if (kPair)
              ptx::mma_pair<sizeof(IT) == 4>(d_tmem, adesc, bdesc, idesc, (ks | kk) != 0 ? 1u : 0u);
            else if (sizeof(IT) == 4)
              ptx::mma_tf32(d_tmem, adesc, bdesc, idesc, (ks | kk) != 0 ? 1u : 0u);
            else
              ptx::mma_f16(d_tmem, adesc, bdesc, idesc, (ks | kk) != 0 ? 1u : 0u);
          }
          // smem slot free (in both CTAs of a pair) when the MMAs retire
          if (kPair) ptx::tc_commit_pair(ptx::smem_u32(&s_empty[stage]));
          else ptx::tc_commit(ptx::smem_u32(&s_empty[stage]));
          if (++stage == dims.stages) { stage = 0; phase ^= 1u; }
        }
        if (kPair) ptx::tc_commit_pair(ptx::smem_u32(&s_tfull[buf]));   // accumulator ready
        else ptx::tc_commit(ptx::smem_u32(&s_tfull[buf]));
        stamp(1, 3 * it + 2);
        buf ^= 1;
        if (buf == 0) bphase ^= 1u;
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue
    constexpr int kSub = ChunkStage<OT>::kSubChunks;       // 32-column sub-chunks per 128-byte chunk
    constexpr int kChunkCols = 32 * kSub;
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const Scaler sc{dims.denom, dims.rcp, dims.rcp_exact};
    const uint32_t stg_base = smem_base + dims.stages * dims.stage_bytes + (uint32_t)(warp - 2) * kStgPerWarp;
    int buf = 0;
    uint32_t bphase = 0;
    int it = 0;
    long long probe_sum[5] = {0, 0, 0, 0, 0};
    for (int item = worker; item < total_items; item += workers, ++it) {
      const int nt = item % dims.NT;
      const int mt = (item / dims.NT) % dims.MT;
      const int bh = item / (dims.NT * dims.MT);
      const int n_valid = min(dims.BN, dims.W2 - nt * dims.BN);
      const int m0 = (kPair ? mt * 2 + (int)rank : mt) * kBlockM + quad * 32;   // first output row of this warp
      ptx::mbar_wait(ptx::smem_u32(&s_tfull[buf]), bphase);
      ptx::tc_fence_after_sync();
      if (warp == 2 && lane == 0) stamp(2, 2 * it);          // accumulator of the item complete
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * kBlockN;
      const int chunks = (n_valid + kChunkCols - 1) / kChunkCols;
      for (int c = 0; c < chunks; ++c) {
        const GroupPlan gp(c, chunks);
        long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
        if (kProbes) t0 = clock64();
        if (kProbes) t1 = clock64();
#pragma unroll
        for (int sub = 0; sub < kSub; ++sub) {
          if (c * kChunkCols + sub * 32 < n_valid || sub == 0) {
            uint32_t acc[32];
            ptx::tmem_ld_32x32(taddr + c * kChunkCols + sub * 32, acc);
            ptx::tmem_ld_wait();
            if (kProbes && sub == 0) t2 = clock64();
            ChunkStage<OT>::template run<sizeof(IT) == 2>(acc, stg_base, lane, sub, sc, num_levels, gp,
                                                          dims.store_mask);
          }
        }
        __syncwarp();                                // staged tiles visible to the whole warp
        if (kProbes) t3 = clock64();
        const int rows_valid = dims.W1 - m0;
        if (rows_valid > 0 && !(kProbes && (dims.debug & 2))) {
          const int n0 = nt * dims.BN;
          const long long grow = (long long)bh * dims.W1 + m0;
#pragma unroll
          for (int i = 0; i < SMOE_MAX_LEVELS; ++i)
            if (i < num_levels && gp.flush[i] && ((dims.store_mask >> i) & 1))
              store_staged_tile<OT>(stg_base + i * kStgLevelBytes, i == 0 ? 0 : gp.d[i],
                                    static_cast<OT*>(out.ptr[i]) + grow * out.pitch[i], out.pitch[i],
                                    (n0 + gp.first[i] * kChunkCols) >> i, out.width[i], rows_valid, lane);
        }
        __syncwarp();                                // tiles may be overwritten by the next chunk
        if (kProbes) {      // per-phase cycle sums: store drain wait, TMEM load, stage, store issue
          const long long t4 = clock64();
          probe_sum[0] += t1 - t0; probe_sum[1] += t2 - t1; probe_sum[2] += t3 - t2; probe_sum[3] += t4 - t3;
          probe_sum[4] += 1;
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {                               // this warp has drained the accumulator
        if (kPair) ptx::mbar_arrive_cluster(ptx::map_to_cta(ptx::smem_u32(&s_tempty[buf]), 0));
        else ptx::mbar_arrive(ptx::smem_u32(&s_tempty[buf]));
      }
      if (warp == 2 && lane == 0) stamp(2, 2 * it + 1);      // epilogue of the item issued
      buf ^= 1;
      if (buf == 0) bphase ^= 1u;
    }
    if (warp == 2 && lane == 0) stamp(2, 15);
    if (kProbes && warp == 2 && lane == 0 && timeline != nullptr)
      for (int i = 0; i < 5; ++i) timeline[((long long)blockIdx.x * 3 + 2) * 16 + 9 + i] = probe_sum[i];
  }

  ptx::tc_fence_before_sync();
  if (kPair) ptx::cluster_sync(); else __syncthreads();   // no CTA leaves while its peer may signal it
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    if (kPair) ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    else ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------- pitch-padding pre-pass
// TMA needs 16-byte global strides AND 16-byte-aligned box start addresses; feature maps whose
// W*sizeof(T) is not a multiple of 16 are copied once into a pitch-padded workspace (zero tail)
// and the tensor map points there.  (A flat (H*W,C,B) map over the unpadded tensor satisfies the
// stride rule when H*W*sizeof(T) % 16 == 0 but faults on rows that start 8 bytes off -- tried.)
template <typename T>
__global__ void __launch_bounds__(256)
pad_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, int W, int Wp) {
  pdl_launch_dependents();
  pdl_wait();
  // one warp per row: coalesced, no per-element division
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (long long r = warp0; r < rows; r += (long long)gridDim.x * 8) {
    const T* s_ = src + r * W;
    T* d_ = dst + r * Wp;
    for (int w = lane; w < Wp; w += 32) d_[w] = w < W ? s_[w] : T(0);
  }
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder();   // shared with build_bwd.cu
static EncodeTiledFn encode_fn() { return tensor_map_encoder(); }
void* tensor_map_encoder_ptr() { return reinterpret_cast<void*>(tensor_map_encoder()); }   // for lookup.cu
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

static int padded_width(int W, int es) {
  const int per16 = 16 / es;
  return (W + per16 - 1) / per16 * per16;
}

// 4-D map over an NCHW feature map (dims W,H,C,B; row pitch Wp elements), box = slab x 1 x KT x 1.
static int make_fmap_tmap(CUtensorMap* out, const void* base, int dtype, bool tf32_round, int B, int C,
                          int H, int W, int Wp) {
  EncodeTiledFn enc = encode_fn();
  SMOE_REQUIRE(enc != nullptr, SMOE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int es = dtype_size(dtype);
  const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)Wp * es, (cuuint64_t)H * Wp * es, (cuuint64_t)C * H * Wp * es};
  const cuuint32_t box[4] = {(cuuint32_t)(128 / es), 1u, (cuuint32_t)(128 / es), 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt =
      dtype == SMOE_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                              : (tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  const CUtensorMapSwizzle sw =
      dtype == SMOE_DTYPE_F16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  const CUresult r = enc(out, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SMOE_REQUIRE(r == CUDA_SUCCESS, SMOE_ERR_CUDA,
               "cuTensorMapEncodeTiled failed with CUresult %d (W=%d H=%d C=%d B=%d pitch=%d)", (int)r, W,
               H, C, B, Wp);
  return SMOE_OK;
}

// 4-D map over a channels-last feature map (memory (B,H,W,C); dims C,W,H,B), box = 128 bytes of C
// x `rows` pixels of one image row.
static int make_fmap_tmap_cl(CUtensorMap* out, const void* base, int dtype, bool tf32_round, int B, int C,
                             int H, int W, int rows) {
  EncodeTiledFn enc = encode_fn();
  SMOE_REQUIRE(enc != nullptr, SMOE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const int es = dtype_size(dtype);
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  const cuuint32_t box[4] = {(cuuint32_t)(128 / es), (cuuint32_t)rows, 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt =
      dtype == SMOE_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                              : (tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  const CUresult r = enc(out, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SMOE_REQUIRE(r == CUDA_SUCCESS, SMOE_ERR_CUDA,
               "cuTensorMapEncodeTiled (channels-last) failed with CUresult %d (W=%d H=%d C=%d B=%d rows=%d)",
               (int)r, W, H, C, B, rows);
  return SMOE_OK;
}

static thread_local long long* g_timeline = nullptr;   // smoe_corr_debug_set_timeline

template <typename IT, typename OT, bool kPair, bool kKMajor = false>
static int launch_build_impl(const CUtensorMap& ta, const CUtensorMap& tb, const PyrDev& out, int num_levels, const BuildDims& dims, cudaStream_t st) {
  auto kern = corr_build_kernel<IT, OT, kPair, kKMajor>;
  const size_t smem_bytes = (size_t)dims.stages * dims.stage_bytes + kBuildSmemFixed;
  SMOE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  int dev = 0, sms = 0;
  SMOE_CUDA_OK(cudaGetDevice(&dev));
  SMOE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long items = (long long)dims.BH * dims.MT * dims.NT;
  const int workers = kPair ? sms / 2 : sms;            // persistent CTAs (or CTA pairs)
  const int grid = (int)(items < workers ? items : workers) * (kPair ? 2 : 1);
  SMOE_CUDA_OK(launch_kernel_cluster(kern, dim3(grid), dim3(kBuildThreads), smem_bytes, st,
                                     kPair ? 2u : 1u, ta, tb, out, num_levels, dims, g_timeline));
  return check_launch("corr_build_kernel");
}
template <typename IT, typename OT>
static int launch_build(const CUtensorMap& ta, const CUtensorMap& tb, const PyrDev& out, int num_levels, const BuildDims& dims, cudaStream_t st) {
  if (dims.kmajor) return launch_build_impl<IT, OT, false, true>(ta, tb, out, num_levels, dims, st);
#ifdef SMOE_CROSSCHECK      // the CTA-pair variant (cta_group::2) measured no faster: cross-check build only
  if (dims.pair) return launch_build_impl<IT, OT, true>(ta, tb, out, num_levels, dims, st);
#endif
  return launch_build_impl<IT, OT, false>(ta, tb, out, num_levels, dims, st);
}

}  // namespace smoe

extern "C" {

#ifdef SMOE_CROSSCHECK
void smoe_corr_debug_set_timeline(void* device_buffer) {
  smoe::g_timeline = static_cast<long long*>(device_buffer);
}
#endif

int64_t smoe_corr_build_workspace_bytes(int B, int C, int H, int W1, int W2, int in_dtype) {
  using namespace smoe;
  if (!dtype_ok(in_dtype) || B < 0 || C < 0 || H < 0 || W1 < 0 || W2 < 0) return 0;
  const int es = dtype_size(in_dtype);
  int64_t bytes = 0;
  const int p1 = padded_width(W1, es), p2 = padded_width(W2, es);
  if (p1 != W1) bytes += ((int64_t)B * C * H * p1 * es + 255) / 256 * 256;
  if (p2 != W2) bytes += ((int64_t)B * C * H * p2 * es + 255) / 256 * 256;
  return bytes;
}

int smoe_corr_build(const void* fmap1, const void* fmap2, int in_dtype, int B, int C, int H, int W1,
                    int W2, const smoe_pyramid_t* pyr, float denom, void* workspace,
                    int64_t workspace_bytes, unsigned flags, smoe_stream_t stream) {
  using namespace smoe;
  if (int rc = validate_pyramid(pyr, "build")) return rc;
  SMOE_REQUIRE(dtype_ok(in_dtype), SMOE_ERR_INVALID_ARG, "build: unknown input dtype %d", in_dtype);
  SMOE_REQUIRE(B >= 0 && C >= 1 && H >= 0 && W1 >= 0 && W2 >= 0, SMOE_ERR_INVALID_ARG,
               "build: bad dims B=%d C=%d H=%d W1=%d W2=%d", B, C, H, W1, W2);
  SMOE_REQUIRE(denom > 0.f && denom < 3.0e38f, SMOE_ERR_INVALID_ARG, "build: denom must be positive");
  SMOE_REQUIRE(pyr->B == B && pyr->H == H && pyr->W1 == W1 && pyr->width[0] == W2, SMOE_ERR_INVALID_ARG,
               "build: pyramid dims (%d,%d,%d,W2=%d) do not match the feature maps (%d,%d,%d,W2=%d)",
               pyr->B, pyr->H, pyr->W1, pyr->width[0], B, H, W1, W2);
  for (int i = 1; i < pyr->num_levels; ++i)
    SMOE_REQUIRE(pyr->width[i] == pyr->width[i - 1] / 2, SMOE_ERR_INVALID_ARG,
                 "build: level %d width %d is not floor(%d/2)", i, pyr->width[i], pyr->width[i - 1]);
  if (B == 0 || H == 0 || W1 == 0 || W2 == 0) return SMOE_OK;
  SMOE_REQUIRE(fmap1 && fmap2, SMOE_ERR_INVALID_ARG, "build: NULL feature map pointer");
  SMOE_REQUIRE(aligned16(fmap1) && aligned16(fmap2), SMOE_ERR_INVALID_ARG,
               "build: feature maps must be 16-byte aligned");
  const int oes = dtype_size(pyr->dtype);
  for (int i = 0; i < pyr->num_levels; ++i) {
    const int per16 = 16 / oes;
    SMOE_REQUIRE(aligned16(pyr->level_ptr[i]) && (pyr->pitch[i] * oes) % 16 == 0 &&
                     ((int64_t)pyr->width[i] + per16 - 1) / per16 * per16 <= pyr->pitch[i],
                 SMOE_ERR_INVALID_ARG,
                 "build: level %d must be 16-byte aligned with a 16-byte-multiple pitch (TMA store) "
                 "(use smoe_corr_pyramid_layout)", i);
  }
  SMOE_REQUIRE((long long)B * H * ((W1 + kBlockM - 1) / kBlockM) * ((W2 + kBlockN - 1) / kBlockN) <
                   (1ll << 31),
               SMOE_ERR_UNSUPPORTED, "build: too many work items");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int es = dtype_size(in_dtype);

  const bool kmajor = (flags & SMOE_BUILD_CHANNELS_LAST) != 0;
  SMOE_REQUIRE(!kmajor || ((long long)C * es) % 16 == 0, SMOE_ERR_UNSUPPORTED,
               "build: channels-last feature maps need C*sizeof(T) %% 16 == 0 (C=%d)", C);

  // pitch-padded staging for W that violates TMA's 16-byte stride rule (NCHW operands only)
  const int p1 = kmajor ? W1 : padded_width(W1, es), p2 = kmajor ? W2 : padded_width(W2, es);
  const void* a_base = fmap1;
  const void* b_base = fmap2;
  if (p1 != W1 || p2 != W2) {
    const int64_t need = smoe_corr_build_workspace_bytes(B, C, H, W1, W2, in_dtype);
    SMOE_REQUIRE(workspace != nullptr && workspace_bytes >= need && aligned16(workspace),
                 SMOE_ERR_INVALID_ARG,
                 "build: W1=%d/W2=%d need a %lld-byte 16-byte-aligned workspace (got %lld)", W1, W2,
                 (long long)need, (long long)workspace_bytes);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const long long rows = (long long)B * C * H;
    auto pad = [&](const void* src, int W, int Wp) -> const void* {
      void* dst = ws;
      const long long total = rows * Wp;
      long long blocks = (rows + 7) / 8;
      if (blocks > 148ll * 8) blocks = 148ll * 8;
      (void)total;
      if (es == 4)
        launch_kernel(pad_rows_kernel<float>, dim3((unsigned)blocks), dim3(256), 0, st,
                      static_cast<const float*>(src), static_cast<float*>(dst), rows, W, Wp);
      else
        launch_kernel(pad_rows_kernel<__half>, dim3((unsigned)blocks), dim3(256), 0, st,
                      static_cast<const __half*>(src), static_cast<__half*>(dst), rows, W, Wp);
      ws += (rows * Wp * es + 255) / 256 * 256;
      return dst;
    };
    if (p1 != W1) a_base = pad(fmap1, W1, p1);
    if (p2 != W2) b_base = pad(fmap2, W2, p2);
    if (int rc = check_launch("pad_rows_kernel")) return rc;
  }

  const bool tf32_round = (flags & SMOE_BUILD_TF32_TRUNCATE) == 0;

  BuildDims dims;
  dims.W1 = W1; dims.W2 = W2; dims.C = C; dims.H = H; dims.BH = B * H;
  dims.kmajor = kmajor ? 1 : 0;
  dims.store_mask = (flags & SMOE_BUILD_SKIP_ODD_LEVELS) ? 0x5 : 0xF;
#ifdef SMOE_CROSSCHECK
  dims.pair = kmajor ? 0 : tuning().build_pair;
#else
  dims.pair = 0;
#endif
  dims.MT = (W1 + (dims.pair ? 2 : 1) * kBlockM - 1) / ((dims.pair ? 2 : 1) * kBlockM);
  dims.NT = (W2 + kBlockN - 1) / kBlockN;
  {
    int cc = 128 / oes;   // columns per 128-byte epilogue chunk: 32 (fp32 out) / 64 (fp16 out)
    if (dims.pair && cc < 2 * (128 / es)) cc = 2 * (128 / es);   // two whole B slabs per pair
    dims.BN = ((W2 + dims.NT - 1) / dims.NT + cc - 1) / cc * cc;
    if (dims.BN > kBlockN) dims.BN = kBlockN;
  }
  dims.debug = 0;
  if (kProbes) dims.debug = tuning().build_debug;
  {
    // B slabs (128 bytes of W2 each) one N tile needs -> bytes per stage -> ring depth
    const int slab = 128 / es;
    const int bn_cta = dims.pair ? dims.BN / 2 : dims.BN;          // B columns one CTA stages
    dims.stage_bytes = kStageABytes + (bn_cta + slab - 1) / slab * (kStageBBytes / (kBlockN / slab));
    dims.stages = (kMaxSmemBytes - kBuildSmemFixed) / dims.stage_bytes;
    if (dims.stages > kMaxStages) dims.stages = kMaxStages;
    const int cap = tuning().build_stages;                  // experiment switch (cross-check build)
    if (cap >= 2 && cap < dims.stages) dims.stages = cap;
  }
  CUtensorMap ta, tb;
  if (kmajor) {
    if (int rc = make_fmap_tmap_cl(&ta, a_base, in_dtype, tf32_round, B, C, H, W1, kBlockM)) return rc;
    if (int rc = make_fmap_tmap_cl(&tb, b_base, in_dtype, tf32_round, B, C, H, W2, dims.BN)) return rc;
  } else {
    if (int rc = make_fmap_tmap(&ta, a_base, in_dtype, tf32_round, B, C, H, W1, p1)) return rc;
    if (int rc = make_fmap_tmap(&tb, b_base, in_dtype, tf32_round, B, C, H, W2, p2)) return rc;
  }
  dims.denom = denom;
  dims.rcp = 1.0f / denom;
  {
    int e = 0;
    dims.rcp_exact = (frexpf(denom, &e) == 0.5f) ? 1 : 0;   // power of two
  }
  // levels whose width has shrunk to 0 have nothing to write
  int levels = pyr->num_levels;
  while (levels > 1 && pyr->width[levels - 1] == 0) --levels;
  const PyrDev pd = to_dev(pyr);
  if (in_dtype == SMOE_DTYPE_F32 && pyr->dtype == SMOE_DTYPE_F16)      // opt-in fp16 storage
    return launch_build<float, __half>(ta, tb, pd, levels, dims, st);
  if (in_dtype == SMOE_DTYPE_F32) return launch_build<float, float>(ta, tb, pd, levels, dims, st);
  if (pyr->dtype == SMOE_DTYPE_F16) return launch_build<__half, __half>(ta, tb, pd, levels, dims, st);
  return launch_build<__half, float>(ta, tb, pd, levels, dims, st);
}

}  // extern "C"
