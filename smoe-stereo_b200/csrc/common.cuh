// common.cuh -- error plumbing and small device helpers shared by the libsmoecorr kernels.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/smoe_corr.h"
#ifdef SMOE_CROSSCHECK
#include "xcheck_api.h"
#endif

namespace smoe {

// ---------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);   // thread-local message, defined in api.cu

#define SMOE_REQUIRE(cond, code, ...)            \
  do {                                           \
    if (!(cond)) {                               \
      ::smoe::set_error(__VA_ARGS__);            \
      return (code);                             \
    }                                            \
  } while (0)

#define SMOE_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t err__ = (expr);                                                         \
    if (err__ != cudaSuccess) {                                                         \
      ::smoe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),      \
                        __FILE__, __LINE__);                                            \
      return SMOE_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(err));
    return SMOE_ERR_CUDA;
  }
  return SMOE_OK;
}

// Launch-time tuning, read ONCE when the library is loaded (api.cu) -- never inside an entry point.
// The product library honours a single documented switch, SMOE_CORR_PDL; every other field is a
// compile-time default there.  The cross-check build (-DSMOE_CROSSCHECK, tests/tools only: it also
// carries the superseded kernels as bit-exact checkers) lets the environment override them for A/B
// measurements (tools/README.md).
struct Tuning {
  // Programmatic dependent launch: the kernel may become resident while its predecessor in the
  // stream is still draining; it must execute pdl_wait() before touching global memory.  On B200
  // back-to-back lane-group lookups measured ~10% SLOWER with PDL, the thread-per-pixel lookup ~6%
  // faster (2.82 vs 3.01 us at cfg1; profiles/r01/README.md), so by default only launches that ask
  // for it (prefer_pdl) get the attribute.  0 never, 1 always, 2 where preferred.
  int pdl = 2;
  int fwd_kernel = 0;      // 0 default (px) | 1 px, 16-byte loads | 2 pair | 3 vec | 4 px, 256-bit loads
  int px_tpb = 32;         // pixels per CTA of the thread-per-pixel lookup
  int px_split = 1;        // one thread per level pair
  int px_ld256 = -1;       // 256-bit segment loads: -1 = for fp32 volumes beyond L2, 0 never, 1 always
  int px_hint = -1;        // L2 eviction priorities of the streaming lookup (lookup.cu): -1 = by volume size
  int lookup_tile = 32;    // lane-group kernels: pixels per CTA
  int bwd_staged = 0;      // first version of the batched backward (cross-check build only)
  int bwd_kernel = 1;      // batched backward: 1 = [pixel][column] tile (lvl), 2 = [column][pixel] tile (cp)
  int bwd_rowmod = 2;      // shared-memory row length of that kernel, modulo 32 banks
  int build_pair = 0;      // CTA-pair build kernel (cross-check build only)
  int build_stages = 0;    // cap of the operand ring depth (0 = fill shared memory)
  int build_debug = 0;     // probe builds (-DSMOE_BUILD_PROBES)
  int conv_lanes = 0;      // lane-group version of the fused lookup+conv (cross-check build only)
};
const Tuning& tuning();    // api.cu
inline int pdl_mode() { return tuning().pdl; }

// cluster_x > 1 launches thread-block clusters of that many CTAs along x (grid.x a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, unsigned cluster_x, bool prefer_pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  const int mode = pdl_mode();
  if (mode == 1 || (mode == 2 && prefer_pdl)) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, unsigned cluster_x, Args... args) {
  return launch_kernel_ex(kern, grid, block, smem, st, cluster_x, false, args...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t st, Args... args) {
  return launch_kernel_ex(kern, grid, block, smem, st, 1u, false, args...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args... args) {
  return launch_kernel_ex(kern, grid, block, smem, st, 1u, true, args...);
}

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline int dtype_size(int dtype) { return dtype == SMOE_DTYPE_F16 ? 2 : 4; }
inline bool dtype_ok(int dtype) { return dtype == SMOE_DTYPE_F32 || dtype == SMOE_DTYPE_F16; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int validate_pyramid(const smoe_pyramid_t* p, const char* who);   // api.cu

// ------------------------------------------------------------------------------ device side
// Kernel-side copy of the pyramid description (passed by value as a kernel parameter).
struct PyrDev {
  void* ptr[SMOE_MAX_LEVELS];
  long long pitch[SMOE_MAX_LEVELS];
  int width[SMOE_MAX_LEVELS];
  int num_levels;
};

inline PyrDev to_dev(const smoe_pyramid_t* p) {
  PyrDev d;
  for (int i = 0; i < SMOE_MAX_LEVELS; ++i) {
    d.ptr[i] = i < p->num_levels ? p->level_ptr[i] : nullptr;
    d.pitch[i] = i < p->num_levels ? p->pitch[i] : 0;
    d.width[i] = i < p->num_levels ? p->width[i] : 0;
  }
  d.num_levels = p->num_levels;
  return d;
}

// 16-byte vector of volume elements -> floats.
// The windows are 40-byte pieces scattered over 2-KB pixel rows.  With a plain ld.global /
// ld.global.nc / .cg the L2 fills the whole 128-byte line from DRAM on a miss (measured 3.96
// sectors per request, 144 MB of DRAM reads per cfg3 launch); the explicit L2::64B prefetch size
// halves the fill (2.39 sectors per request, 87 MB) and L1::no_allocate keeps these
// single-use lines out of L1 (profiles/r01/README.md, "load variants").
__device__ __forceinline__ uint4 ld_window16(const void* p) {
  uint4 t;
  asm volatile("ld.global.L1::no_allocate.L2::64B.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
               : "l"(p));
  return t;
}
// STREAM=true : volume larger than L2 -> 64-byte fills, no L1 allocation (less DRAM traffic).
// STREAM=false: L2-resident volume    -> read-only path; measured faster there (3.65 vs 4.50 us
//               per cfg1 lookup) because nothing goes to DRAM anyway.
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  template <bool STREAM>
  __device__ static void load(const float* p, float (&v)[4]) {
    const uint4 u = STREAM ? ld_window16(p) : __ldg(reinterpret_cast<const uint4*>(p));
    const float4 t = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z),
                                 __uint_as_float(u.w));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  template <bool STREAM>
  __device__ static void load(const __half* p, float (&v)[8]) {
    const uint4 t = STREAM ? ld_window16(p) : __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
};

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// Optional coordinate-update glue fused into the lookup prologue: the GRU
// loop's  delta_flow[:,1] = 0; coords1 = coords1 + delta_flow; flow = coords1 - coords0
// (core/SMoEStereo.py:236,245-248) -- three tiny elementwise launches per iteration in the
// reference.  All tensors are contiguous (B,2,H,W1); `plane` = H*W1.  c1_out == nullptr: plain lookup.
struct CoordUpd {
  const float* delta;      // GRU update; only channel 0 (x) is applied.  May be nullptr.
  const float* c0;         // coords0, needed iff flow_out != nullptr
  float* c1_out;           // updated coordinates (must not alias the input coordinates)
  float* flow_out;         // c1_out - coords0, or nullptr
  long long plane;
};

// x coordinate of pixel (b, hw); in update mode one designated lane per pixel also writes the
// updated coordinates / flow.
__device__ __forceinline__ float fetch_coord(const float* __restrict__ coords, long long bstride,
                                             const CoordUpd& u, int b, int hw, bool writer) {
  float x = __ldg(coords + (long long)b * bstride + hw);
  if (u.c1_out != nullptr) {
    const long long o = (long long)b * 2 * u.plane + hw;
    if (u.delta != nullptr) x += __ldg(u.delta + o);
    if (writer) {
      const float y = __ldg(coords + (long long)b * bstride + u.plane + hw);
      u.c1_out[o] = x;
      u.c1_out[o + u.plane] = y;
      if (u.flow_out != nullptr) {
        u.flow_out[o] = x - __ldg(u.c0 + o);
        u.flow_out[o + u.plane] = y - __ldg(u.c0 + o + u.plane);
      }
    }
  }
  return x;
}

// Split a level coordinate into (base tap index, fractional weight).  Non-finite or
// unrepresentably large coordinates are sent far out of range so that every tap is masked
// (the reference's int cast of such values is undefined; all-zero is the defined behaviour here).
__device__ __forceinline__ void split_coord(float xi, int radius, int& base, float& dx) {
  float fl = floorf(xi);
  dx = xi - fl;
  if (!(fl > -1.0e8f && fl < 1.0e8f)) {   // also catches NaN
    fl = -1.0e8f;
    dx = 0.f;
  }
  base = static_cast<int>(fl) - radius;
}

// Interpolation in the reference's evaluation order (sampler_kernel.cu:46-59):
//   out = 0; out += v0*(1-dx); out += v1*dx    (tap i=k first, then tap i=k+1)
template <typename OT> __device__ __forceinline__ OT lerp_ref(float v0, float v1, float dx);
template <> __device__ __forceinline__ float lerp_ref<float>(float v0, float v1, float dx) {
  return fmaf(v1, dx, v0 * (1.0f - dx));
}
// scalar_t == c10::Half: every product and sum is computed in fp32 and rounded to half
// (c10/util/Half-inl.h operators), weights are rounded to half first (scalar_t(dx)).
template <> __device__ __forceinline__ __half lerp_ref<__half>(float v0, float v1, float dx) {
  const float w1 = __half2float(__float2half_rn(1.0f - dx));
  const float w0 = __half2float(__float2half_rn(dx));
  const float a = __half2float(__float2half_rn(v0 * w1));
  const float b = __half2float(__float2half_rn(v1 * w0));
  return __float2half_rn(a + b);
}

// Interpolate a chunk: e[j] = lerp(v[j], v[j+1]) with v[N] = nxt (first value of the next chunk).
// Generic form: lerp_ref per element.
template <typename OT, int N> struct LerpChunk {
  __device__ static __forceinline__ void run(const float (&v)[N], float nxt, float dx, float (&e)[N]) {
#pragma unroll
    for (int j = 0; j < N; ++j) e[j] = to_float(lerp_ref<OT>(v[j], (j + 1 < N) ? v[j + 1] : nxt, dx));
  }
};
// fp16 outputs: the reference's c10::Half arithmetic (fp32 op, round to half after every multiply
// and add) equals native IEEE half arithmetic bit for bit -- the fp32 product of two halves is
// exact, and the fp32 sum of two halves is either exact or far from an fp16 rounding tie -- so the
// chunk is done with packed HMUL2 / HADD2 (the _rn forms are never contracted into HFMA2):
// ~3 instructions per output instead of ~9 conversions and fp32 ops.
template <int N> struct LerpChunk<__half, N> {
  static_assert(N % 2 == 0, "packed half path needs an even chunk");
  __device__ static __forceinline__ void run(const float (&v)[N], float nxt, float dx, float (&e)[N]) {
    const __half2 w1 = __float2half2_rn(1.0f - dx);
    const __half2 w0 = __float2half2_rn(dx);
    __half2 a[N / 2], bm[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);      // exact: the values are fp16
      a[i] = __hmul2_rn(h, w1);
      bm[i] = __hmul2_rn(h, w0);
    }
    const __half bn = __hmul_rn(__float2half_rn(nxt), __low2half(w0));
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      const __half2 bs = __halves2half2(__high2half(bm[i]), (i + 1 < N / 2) ? __low2half(bm[i + 1]) : bn);
      const float2 f = __half22float2(__hadd2_rn(a[i], bs));
      e[2 * i] = f.x;
      e[2 * i + 1] = f.y;
    }
  }
};

// The N taps of a window: out[k] = lerp(v[k], v[k+1]), k < N, in LT's arithmetic.  Generic: lerp_ref
// per tap.  LT == __half: the packed native form of LerpChunk above (bit-identical to the c10::Half
// emulation) -- what the thread-per-pixel lookup spends on its 18 taps drops from ~9 to ~3
// instructions per tap.
template <typename LT, int N> struct LerpTaps {
  __device__ static __forceinline__ void run(const float* v, float dx, float (&out)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = to_float(lerp_ref<LT>(v[k], v[k + 1], dx));
  }
};
template <int N> struct LerpTaps<__half, N> {
  __device__ static __forceinline__ void run(const float* v, float dx, float (&out)[N]) {
    constexpr int P = (N + 2) / 2;                                     // half2 registers covering v[0..N]
    const __half2 w1 = __float2half2_rn(1.0f - dx);
    const __half2 w0 = __float2half2_rn(dx);
    __half2 a[P], b[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const __half2 h = __floats2half2_rn(v[2 * i], (2 * i + 1 <= N) ? v[2 * i + 1] : 0.f);   // exact: fp16 values
      a[i] = __hmul2_rn(h, w1);
      b[i] = __hmul2_rn(h, w0);
    }
#pragma unroll
    for (int i = 0; 2 * i < N; ++i) {
      const __half2 bs = __halves2half2(__high2half(b[i]), (i + 1 < P) ? __low2half(b[i + 1]) : __float2half_rn(0.f));
      const float2 f = __half22float2(__hadd2_rn(a[i], bs));
      out[2 * i] = f.x;
      if (2 * i + 1 < N) out[2 * i + 1] = f.y;
    }
  }
};

}  // namespace smoe
