// api.cu -- ABI bookkeeping: version, thread-local error string, pyramid layout arithmetic.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace smoe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

static Tuning read_tuning() {
  Tuning t;
  // SMOE_CORR_PDL: unset -> 2 (only the kernels that measured faster with it), 0 -> never,
  // anything else -> every kernel
  if (const char* e = getenv("SMOE_CORR_PDL"))
    if (e[0]) t.pdl = e[0] == '0' ? 0 : 1;
#ifdef SMOE_CROSSCHECK
  if (const char* e = getenv("SMOE_CORR_LOOKUP_KERNEL"))
    t.fwd_kernel = e[0] == 'p' && e[1] == 'x' ? (e[2] == '2' ? 4 : 1) : (e[0] == 'p' ? 2 : (e[0] == 'v' ? 3 : 0));
  t.px_tpb = env_int("SMOE_CORR_PX_TPB", t.px_tpb);
  t.px_split = env_int("SMOE_CORR_PX_SPLIT", t.px_split);
  t.px_ld256 = env_int("SMOE_CORR_PX_LD256", t.px_ld256);
  t.px_hint = env_int("SMOE_CORR_PX_HINT", t.px_hint);
  t.lookup_tile = env_int("SMOE_CORR_LOOKUP_TILE", t.lookup_tile);
  if (const char* e = getenv("SMOE_CORR_BWD_BATCHED")) t.bwd_staged = e[0] == 's';
  t.bwd_kernel = env_int("SMOE_CORR_BWD_KERNEL", t.bwd_kernel);
  t.bwd_rowmod = env_int("SMOE_CORR_BWD_ROWMOD", t.bwd_rowmod);
  t.build_pair = env_int("SMOE_CORR_BUILD_PAIR", 0);
  t.build_stages = env_int("SMOE_CORR_BUILD_STAGES", 0);
  t.build_debug = env_int("SMOE_CORR_BUILD_DEBUG", 0);
  if (const char* e = getenv("SMOE_CORR_LOOKUP_CONV")) t.conv_lanes = e[0] == 'l';
#else
  (void)env_int;
#endif
  return t;
}

#ifdef SMOE_CROSSCHECK
static Tuning g_tuning = read_tuning();           // the cross-check library may change it at run time
#else
static const Tuning g_tuning = read_tuning();     // dynamic initialisation == library load
#endif
const Tuning& tuning() { return g_tuning; }

int validate_pyramid(const smoe_pyramid_t* p, const char* who) {
  SMOE_REQUIRE(p != nullptr, SMOE_ERR_INVALID_ARG, "%s: pyramid is NULL", who);
  SMOE_REQUIRE(p->num_levels >= 1 && p->num_levels <= SMOE_MAX_LEVELS, SMOE_ERR_INVALID_ARG,
               "%s: num_levels=%d outside [1,%d]", who, p->num_levels, SMOE_MAX_LEVELS);
  SMOE_REQUIRE(dtype_ok(p->dtype), SMOE_ERR_INVALID_ARG, "%s: unknown pyramid dtype %d", who,
               p->dtype);
  SMOE_REQUIRE(p->B >= 0 && p->H >= 0 && p->W1 >= 0, SMOE_ERR_INVALID_ARG,
               "%s: negative pixel dims (%d,%d,%d)", who, p->B, p->H, p->W1);
  const long long rows = (long long)p->B * p->H * p->W1;
  SMOE_REQUIRE(rows < (1ll << 31), SMOE_ERR_UNSUPPORTED, "%s: %lld pixel rows exceed 2^31", who,
               rows);
  for (int i = 0; i < p->num_levels; ++i) {
    SMOE_REQUIRE(p->width[i] >= 0, SMOE_ERR_INVALID_ARG, "%s: level %d width %d < 0", who, i,
                 p->width[i]);
    SMOE_REQUIRE(p->pitch[i] >= p->width[i], SMOE_ERR_INVALID_ARG,
                 "%s: level %d pitch %lld < width %d", who, i, (long long)p->pitch[i], p->width[i]);
    SMOE_REQUIRE(rows == 0 || p->width[i] == 0 || p->level_ptr[i] != nullptr,
                 SMOE_ERR_INVALID_ARG, "%s: level %d pointer is NULL", who, i);
  }
  return SMOE_OK;
}

}  // namespace smoe

extern "C" {

#ifdef SMOE_CROSSCHECK
// Tests / probes: set one Tuning field by name (process-wide; not thread-safe by design).
int smoe_corr_debug_set_tuning(const char* name, int value) {
  using smoe::g_tuning;
#define SMOE_FIELD(f) if (strcmp(name, #f) == 0) { g_tuning.f = value; return 0; }
  SMOE_FIELD(pdl) SMOE_FIELD(fwd_kernel) SMOE_FIELD(px_tpb) SMOE_FIELD(px_split) SMOE_FIELD(px_ld256)
  SMOE_FIELD(px_hint) SMOE_FIELD(lookup_tile) SMOE_FIELD(bwd_staged) SMOE_FIELD(bwd_kernel)
  SMOE_FIELD(bwd_rowmod) SMOE_FIELD(build_pair) SMOE_FIELD(build_stages) SMOE_FIELD(build_debug)
  SMOE_FIELD(conv_lanes)
#undef SMOE_FIELD
  return 1;
}
#endif

int smoe_corr_abi_version(void) { return SMOE_CORR_ABI_VERSION; }

const char* smoe_corr_last_error(void) { return smoe::g_err; }

int smoe_corr_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return 0;
  return major == 10 ? 1 : 0;
}

int smoe_corr_pyramid_layout(int W2, int num_levels, int dtype, int32_t* widths,
                             int64_t* level_offset, int64_t* pitch) {
  using namespace smoe;
  SMOE_REQUIRE(W2 >= 0, SMOE_ERR_INVALID_ARG, "pyramid_layout: W2=%d < 0", W2);
  SMOE_REQUIRE(num_levels >= 1 && num_levels <= SMOE_MAX_LEVELS, SMOE_ERR_INVALID_ARG,
               "pyramid_layout: num_levels=%d outside [1,%d]", num_levels, SMOE_MAX_LEVELS);
  SMOE_REQUIRE(dtype_ok(dtype), SMOE_ERR_INVALID_ARG, "pyramid_layout: unknown dtype %d", dtype);
  SMOE_REQUIRE(widths && level_offset && pitch, SMOE_ERR_INVALID_ARG,
               "pyramid_layout: NULL output pointer");
  const int per16 = 16 / dtype_size(dtype);     // elements per 16 bytes
  const int per128 = 128 / dtype_size(dtype);   // elements per 128-byte line
  int64_t off = 0;
  int w = W2;
  for (int i = 0; i < num_levels; ++i) {
    widths[i] = w;
    level_offset[i] = off;
    off += ((int64_t)w + per16 - 1) / per16 * per16;   // every level starts 16-byte aligned
    w /= 2;                                            // floor-halving, core/corr.py:124
  }
  if (off == 0) off = per128;
  *pitch = (off + per128 - 1) / per128 * per128;       // rows start on 128-byte lines
  return SMOE_OK;
}

}  // extern "C"
