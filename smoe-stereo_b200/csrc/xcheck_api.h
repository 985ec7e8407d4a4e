/*
 * xcheck_api.h -- entry points that exist ONLY in the cross-check library
 * (tests/_build/libsmoecorr_xcheck.so, built with -DSMOE_CROSSCHECK).  Not part of the product ABI.
 */
#ifndef SMOE_XCHECK_API_H_
#define SMOE_XCHECK_API_H_
#ifdef __cplusplus
extern "C" {
#endif

/*
 * When a device buffer of 148 * 3 * 16 int64 is set on the calling thread, the next
 * smoe_corr_build launches record per-CTA clock64 stamps of the producer / MMA / epilogue roles
 * into it.  NULL disables.
 */
void smoe_corr_debug_set_timeline(void* device_buffer);

/*
 * Force one forward lookup kernel on the calling thread (parity tests compare the kernels bit for
 * bit).  0 = automatic choice by volume size, 1 = thread-per-pixel kernel with 16-byte loads,
 * 2 = level-pair lane-group kernel, 3 = four-window lane-group kernel, 4 = thread-per-pixel kernel
 * with 256-bit loads (fp32 volumes), -1 = back to the SMOE_CORR_LOOKUP_KERNEL environment setting.
 * Shapes a kernel does not cover fall through to the automatic choice.
 */
void smoe_corr_debug_set_lookup_kernel(int which);

/*
 * Set one launch-tuning field (struct Tuning, csrc/common.cuh) by name for the whole process:
 * "bwd_kernel", "px_hint", "build_pair", "bwd_staged", "conv_lanes", ...  Returns 0, or 1 for an
 * unknown name.  Replaces the SMOE_CORR_* environment switches when a test or probe wants to compare
 * variants inside one process.
 */
int smoe_corr_debug_set_tuning(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif
