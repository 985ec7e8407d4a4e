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

// SMOE_CORR_PDL: unset -> 2 (programmatic dependent launch only for the kernels that measured
// faster with it: the thread-per-pixel lookup), 0 -> never, anything else -> every kernel.
int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("SMOE_CORR_PDL");
    if (!e || !e[0]) return 2;
    return e[0] == '0' ? 0 : 1;
  }();
  return mode;
}

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
