/*
 * smoe_corr.h -- C ABI of libsmoecorr.so, the B200 (sm_100a) correlation hot path.
 *
 * This is the drop-in boundary for the RAFT-Stereo correlation path of SMoE-Stereo.
 * Every entry point takes raw device pointers, explicit sizes and a CUDA stream;
 * nothing here allocates, synchronises, or depends on torch/ATen/pybind.  All
 * functions return SMOE_OK (0) or a non-zero error code; smoe_corr_last_error()
 * returns a thread-local description of the last failure.  The library keeps no
 * global mutable state and is re-entrant (the reference is driven from one host thread
 * per GPU under nn.DataParallel, train_stereo_sceneflow.py:133).
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   smoe_corr_build            core/corr.py:148-156 (CorrBlock1D.corr: einsum + /sqrt(D))
 *                              + core/corr.py:119-125 / :38-42 (avg_pool2d pyramid)
 *   smoe_corr_lookup           core/corr.py:127-146 (CorrBlock1D.__call__), core/corr.py:44-51
 *   smoe_corr_lookup_to        the same + the output gather of nn.DataParallel, train_stereo_sceneflow.py:133
 *                              (CorrBlockFast1D.__call__), sampler/sampler.cpp:24-32 and
 *                              sampler/sampler_kernel.cu:19-60,107-136 (corr_sampler.forward)
 *   smoe_corr_lookup_update    the loop glue of core/SMoEStereo.py:234-248 + the lookup (opt-in)
 *   smoe_corr_lookup_conv1x1   the lookup + BasicMotionEncoder.convc1 + ReLU, core/update.py:70,77
 *                              (opt-in)
 *   smoe_corr_lookup_bwd       sampler/sampler.cpp:34-45, sampler/sampler_kernel.cu:63-105,138-166
 *                              (corr_sampler.backward)
 *   smoe_corr_fold_pyramid_grad  autograd of core/corr.py:122-125 (avg_pool2d backward chain)
 *   smoe_corr_build_bwd        autograd of core/corr.py:154-156 (einsum backward: d fmap1, d fmap2)
 *
 * Data layout
 * -----------
 * Feature maps are NCHW contiguous, (B,C,H,W), fp32 or fp16 (core/extractor.py:603-609); or, with
 * SMOE_BUILD_CHANNELS_LAST, the same logical tensor stored (B,H,W,C).
 * Coordinates are (B,coords_ch,H,W1) fp32; only channel 0 (x) is read (core/corr.py:129,
 * :47).  Lookup outputs are (B, L*(2r+1), H, W1) contiguous, channel = level*(2r+1)+tap
 * (core/corr.py:143-146).
 * The correlation pyramid is described by smoe_pyramid_t: per level a base pointer, the
 * level width W2_i = floor(W2 / 2^i) and the pitch (elements) between consecutive pixel rows.
 * Pixel row index = (b*H + y)*W1 + x.  Two layouts are in use:
 *   - "fused"   (owned by CorrBlock1D/CorrBlockFast1D): one buffer, all levels of a pixel in one
 *               row, level_ptr[i] = base + level_offset[i], common pitch (smoe_corr_pyramid_layout);
 *   - "planar"  (corr_sampler.forward/backward API): a single dense (B,H,W1,W2) level, pitch = W2.
 */
#ifndef SMOE_CORR_H_
#define SMOE_CORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMOE_CORR_ABI_VERSION 2
#define SMOE_MAX_LEVELS 4

/* error codes */
#define SMOE_OK 0
#define SMOE_ERR_INVALID_ARG 1   /* bad pointer / size / dtype combination            */
#define SMOE_ERR_UNSUPPORTED 2   /* valid request this build cannot serve             */
#define SMOE_ERR_CUDA 3          /* a CUDA runtime/driver call failed (see last_error) */

/* element types */
#define SMOE_DTYPE_F32 0
#define SMOE_DTYPE_F16 1

/* smoe_corr_build flags */
#define SMOE_BUILD_TF32_TRUNCATE 1 /* fp32 inputs: load with a FLOAT32 tensor map so that tcgen05
                                      kind::tf32 truncates the low 13 mantissa bits itself; default
                                      (0) is a TFLOAT32 tensor map (TMA converts on the way in)  */

#define SMOE_BUILD_CHANNELS_LAST 2 /* feature maps are channels-last: (B,H,W,C) in memory (what a
                                      torch NCHW-shaped tensor in memory_format=channels_last is, and
                                      what a channels-last producer convolution emits; SURVEY.md 8f
                                      row f4).  Both operands are then K-major for the tensor cores.
                                      Needs C*sizeof(T) % 16 == 0; no constraint on W1/W2 and no
                                      workspace.                                                   */

#define SMOE_BUILD_SKIP_ODD_LEVELS 4 /* do not write levels 1 and 3.  The radius-4 lookup kernels read
                                      levels 0 and 2 only (level i+1 is the pair average of level i
                                      and is recomputed on the fly, bit-exactly), so a third of the
                                      pyramid bytes never has to reach memory.  The caller then sets
                                      levels_missing = 0b1010 (masked to num_levels) in the pyramids
                                      it passes to smoe_corr_lookup*.                              */

/* smoe_corr_lookup_bwd modes */
#define SMOE_BWD_OVERWRITE 0     /* write every element of the gradient (zeros outside the taps):
                                    == zeros_like + scatter of sampler_kernel.cu:148-163         */
#define SMOE_BWD_ACCUMULATE 1    /* read-modify-write the touched taps of a persistent gradient  */

typedef void* smoe_stream_t;     /* cudaStream_t */

typedef struct smoe_pyramid {
  void* level_ptr[SMOE_MAX_LEVELS];     /* device pointers                                   */
  int32_t width[SMOE_MAX_LEVELS];       /* W2_i                                              */
  int64_t pitch[SMOE_MAX_LEVELS];       /* elements between consecutive pixel rows           */
  int32_t num_levels;                   /* 1..SMOE_MAX_LEVELS                                */
  int32_t dtype;                        /* SMOE_DTYPE_*                                      */
  int32_t B, H, W1;                     /* pixel rows = B*H*W1                               */
  int32_t levels_missing;               /* bit i set: level i holds no data (it was skipped by
                                           smoe_corr_build(SMOE_BUILD_SKIP_ODD_LEVELS)); entry
                                           points that would read it return SMOE_ERR_INVALID_ARG.
                                           0 = every level is present                        */
} smoe_pyramid_t;

/* ABI version of the loaded library (== SMOE_CORR_ABI_VERSION). */
int smoe_corr_abi_version(void);

/* Thread-local, NUL-terminated description of the last error on the calling thread. */
const char* smoe_corr_last_error(void);

/* 1 if the current device can run the kernels (compute capability 10.x), else 0. */
int smoe_corr_device_supported(void);

/*
 * Fused pyramid layout for a volume of width W2: widths[i] = floor(W2/2^i)
 * (core/corr.py:124 floor-halving), level_offset[i] (elements, 16-byte aligned) and the row
 * pitch (elements, 128-byte aligned).  Pure host arithmetic.
 */
int smoe_corr_pyramid_layout(int W2, int num_levels, int dtype, int32_t* widths,
                             int64_t* level_offset, int64_t* pitch);

/* Bytes of device scratch smoe_corr_build needs for these shapes (0 when W1 and W2 satisfy the
 * TMA 16-byte pitch rule; otherwise room for pitch-padded copies of the feature maps). */
int64_t smoe_corr_build_workspace_bytes(int B, int C, int H, int W1, int W2, int in_dtype);

/*
 * Correlation volume + pyramid:
 *   level0[(b,h,w1), w2] = (sum_c fmap1[b,c,h,w1] * fmap2[b,c,h,w2]) / denom     (denom = sqrt(C))
 *   level(i+1)[., j]     = (level i[., 2j] + level i[., 2j+1]) / 2,  j < floor(W2_i/2)
 * computed on tcgen05 tensor cores (kind::tf32 for fp32 inputs, kind::f16 for fp16 inputs,
 * fp32 accumulation in TMEM), operands staged by TMA, scale and pooling fused in the epilogue.
 * pyr->dtype is the output type: fp32, or fp16.  fp16 inputs + fp16 pyramid is the reference's
 * reg_cuda path under autocast (each fp16 level is pooled from the rounded previous level, as
 * F.avg_pool2d on a half tensor does).  fp32 inputs + fp16 pyramid is an opt-in storage format
 * (one rounding of the fp32 result, ~5e-4 relative): half the pyramid bytes, so e.g. the B=8
 * training volume fits in L2.  The division is a true fp32 division as in
 * core/corr.py:156 (for fp16 outputs it is applied to the fp16-rounded sum, as the reference's
 * fp16 einsum output is).
 */
int smoe_corr_build(const void* fmap1, const void* fmap2, int in_dtype, int B, int C, int H,
                    int W1, int W2, const smoe_pyramid_t* pyr, float denom, void* workspace,
                    int64_t workspace_bytes, unsigned flags, smoe_stream_t stream);

/*
 * Lookup: for every pixel and level i, with xi = coords[b,0,y,x] * coord_scale / 2^i,
 *   out[b, i*(2r+1)+k, y, x] = (1-dx)*L_i[row, fl-r+k] + dx*L_i[row, fl-r+k+1],
 *   fl = floor(xi), dx = xi - fl, taps outside [0, W2_i) contribute 0.
 * coords_batch_stride = elements between batches of coords (coords_ch*H*W1).
 * With more than one level the levels must form a pyramid as smoe_corr_build writes it (level
 * i+1 == pair average of level i): for volumes larger than L2 the kernel reads one level-0 and
 * one level-2 segment per pixel and recomputes the level-1 / level-3 taps from them with the
 * same fp32 arithmetic (bit-identical).  Single-level calls (corr_sampler.forward) read exactly
 * the level they are given.
 * out_dtype fp32, or fp16 for an fp16 pyramid (arithmetic then follows the reference's
 * half-precision accumulation order, sampler_kernel.cu:46-59).
 */
int smoe_corr_lookup(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                     float coord_scale, int radius, void* out, int out_dtype,
                     smoe_stream_t stream);

/*
 * Lookup with the GRU loop's coordinate-update glue fused into the kernel prologue (opt-in
 * extension).  Replaces, per iteration, the reference's elementwise launches
 *   delta_flow[:,1] = 0; coords1 = coords1 + delta_flow        (core/SMoEStereo.py:245-248)
 *   corr = corr_fn(coords1); flow = coords1 - coords0           (core/SMoEStereo.py:235-236)
 * by ONE launch:  coords1_out = coords1 + (delta_flow.x, 0);  flow_out = coords1_out - coords0;
 * out = lookup(coords1_out.x).  All coordinate tensors are contiguous (B,2,H,W1) fp32.
 * delta_flow may be NULL (no update), flow_out may be NULL; coords1_out must not alias coords1.
 */
int smoe_corr_lookup_update(const smoe_pyramid_t* pyr, const float* coords1, const float* delta_flow,
                            const float* coords0, float* coords1_out, float* flow_out, int radius,
                            void* out, int out_dtype, smoe_stream_t stream);

/*
 * Lookup written through a caller-described view of a larger output, optionally a multicast address:
 * the fused form of "lookup, then all-gather the result" for a path that is sharded by batch or by
 * image rows over the GPUs of one NVSwitch domain (the reference's counterpart is the
 * gather nn.DataParallel performs on the outputs, train_stereo_sceneflow.py:133).
 *   element (b, ch, y, x) of this rank's lookup goes to  out + b*out_batch_stride + ch*out_channel_stride
 *   + y*W1 + x  (strides in elements; the caller has already offset `out` to its shard's place)
 *   SMOE_OUT_MULTIMEM: `out` is a multicast (NVLS) address mapped over one buffer per GPU; every value
 *   is stored once with multimem.st and the switch replicates it into all of them.  The caller owns
 *   the cross-GPU synchronisation that makes the result visible (a barrier after the call).
 * Covers radius 4, 3 or 4 levels, fp32 outputs, pyramids laid out by smoe_corr_pyramid_layout; other
 * cases return SMOE_ERR_UNSUPPORTED (use smoe_corr_lookup and a collective).  Values are bit-identical
 * to smoe_corr_lookup.
 */
#define SMOE_OUT_MULTIMEM 1u
#define SMOE_OUT_SCALAR 2u     /* debugging / A-B: 4-byte stores even where 16-byte stores are possible */
int smoe_corr_lookup_to(const smoe_pyramid_t* pyr, const float* coords, int64_t coords_batch_stride,
                        float coord_scale, int radius, void* out, int64_t out_batch_stride,
                        int64_t out_channel_stride, int out_dtype, unsigned out_flags,
                        smoe_stream_t stream);

/*
 * Lookup fused with the 1x1 convolution that consumes it (opt-in extension):
 *   out[b,o,y,x] = act( bias[o] + sum_j weight[o,j] * lookup(pyr, coords)[b,j,y,x] )
 * which is `F.relu(self.convc1(corr))` of the reference's motion encoder (core/update.py:70,77,
 * Conv2d(L*(2r+1), 64, kernel 1)) applied to what smoe_corr_lookup would have returned; the
 * L*(2r+1)-channel lookup result never reaches global memory.  Inference only (no backward).
 *   lookup_dtype  arithmetic/rounding of the lookup stage = the out_dtype smoe_corr_lookup would be
 *                 called with (fp32 for CorrBlock1D, the volume dtype for CorrBlockFast1D)
 *   weight        (cout, L*(2r+1)) fp32 contiguous (Conv2d.weight viewed 2-D); bias (cout) fp32 or NULL
 *   relu          non-zero: act = max(.,0); zero: identity
 *   out           (B, cout, H, W1) contiguous, out_dtype
 * The products run on the tensor cores in tf32 (lookup values and weights rounded to nearest,
 * fp32 accumulate): ~3e-4 relative to an fp32 convolution, exact to fp16-rounding when the
 * reference runs this layer under autocast.  Supported: radius 4, cout a multiple of 16 up to 256,
 * 16-byte aligned levels (smoe_corr_pyramid_layout); otherwise SMOE_ERR_UNSUPPORTED -- run
 * smoe_corr_lookup and the convolution separately.
 */
int smoe_corr_lookup_conv1x1(const smoe_pyramid_t* pyr, const float* coords,
                             int64_t coords_batch_stride, float coord_scale, int radius,
                             int lookup_dtype, const float* weight, const float* bias, int cout,
                             int relu, void* out, int out_dtype, smoe_stream_t stream);

/*
 * Lookup backward (gradient w.r.t. the pyramid levels; coords get no gradient, core/corr.py:29).
 *   grad_pyr level i [row, fl-r+t] (+)= grad_out[i*(2r+1)+t-1]*dx + grad_out[i*(2r+1)+t]*(1-dx)
 * grad_out: (B, L*(2r+1), H, W1), dtype grad_out_dtype.  grad_pyr->dtype: fp32 (or fp16 in
 * OVERWRITE mode for the fp16 corr_sampler.backward API).
 */
int smoe_corr_lookup_bwd(const smoe_pyramid_t* grad_pyr, const float* coords,
                         int64_t coords_batch_stride, float coord_scale, int radius,
                         const void* grad_out, int grad_out_dtype, int mode,
                         smoe_stream_t stream);

/*
 * Lookup backward for a whole step at once.  Autograd runs every lookup backward of a step before
 * the build backward, so the num_iters lookups can be back-propagated together: per 32-pixel tile
 * the gradient rows of all `num_levels` levels live in shared memory, the (coords[t], grad_out[t])
 * tiles are streamed through, the levels are folded (avg_pool2d backward chain) and ONLY the folded
 * level-0 gradient is written, once -- instead of num_iters read-modify-write passes over a
 * gradient pyramid, a memset and a fold pass.  g0 describes that level-0 output (fp32; level_ptr[0],
 * width[0] = W2, pitch[0]); accumulate != 0 adds to its current contents.  Result ==
 * smoe_corr_lookup_bwd(ACCUMULATE) x num_iters followed by smoe_corr_fold_pyramid_grad.
 * Host arrays of num_iters device pointers / strides.  Radius 4 only, and W2 small enough for the
 * gradient rows of 32 pixels to fit shared memory (W2 <= ~900 for 4 levels); SMOE_ERR_UNSUPPORTED
 * otherwise: use the per-call path.  No alignment rule on the gradient planes or coordinates.
 */
int smoe_corr_lookup_bwd_batched(const smoe_pyramid_t* g0, int num_levels, int num_iters,
                                 const float* const* coords, const int64_t* coords_batch_stride,
                                 const void* const* grad_out, int grad_out_dtype, int radius,
                                 int accumulate, smoe_stream_t stream);

/*
 * Fold the per-level gradients onto level 0 in place (avg_pool2d backward chain):
 *   g_{i-1}[., 2j] += g_i[., j]/2 ; g_{i-1}[., 2j+1] += g_i[., j]/2, for i = L-1 .. 1.
 */
int smoe_corr_fold_pyramid_grad(const smoe_pyramid_t* grad_pyr, smoe_stream_t stream);

/*
 * Fill in the levels a build with SMOE_BUILD_SKIP_ODD_LEVELS left unwritten: for every bit i set in
 * pyr->levels_missing (i >= 1, level i-1 present or filled earlier in the same call)
 *   level i[., j] = (level i-1[., 2j] + level i-1[., 2j+1]) * 0.5      (fp32 sum, one rounding)
 * -- the arithmetic of the build epilogue / F.avg_pool2d (core/corr.py:124), so the result is
 * bit-identical to a full build.  Only consumers other than the radius-4 lookup need it (the
 * reference-shaped `corr_pyramid` views, core/corr.py:41).  The caller clears levels_missing.
 */
int smoe_corr_fill_levels(const smoe_pyramid_t* pyr, smoe_stream_t stream);

/*
 * Volume-build backward on tcgen05 tensor cores (kind::tf32): with G = grad_pyr level 0 after
 * smoe_corr_fold_pyramid_grad (fp32),
 *   dfmap1[b,c,h,w1] = (sum_w2 G[(b,h,w1),w2] * fmap2[b,c,h,w2]) / denom
 *   dfmap2[b,c,h,w2] = (sum_w1 G[(b,h,w1),w2] * fmap1[b,c,h,w1]) / denom
 * i.e. the autograd of core/corr.py:154-156 (einsum, then / sqrt(D)).  fp32 NCHW feature maps and
 * outputs.  row_pitch1 / row_pitch2 = elements between consecutive image rows of fmap1+dfmap1 /
 * fmap2+dfmap2 (>= W, a multiple of 4: TMA wants 16-byte strides; pass W itself for contiguous
 * tensors whose W is a multiple of 4, a padded pitch otherwise).  Strides: (C*H*pitch, H*pitch, pitch, 1).
 */
int smoe_corr_build_bwd(const void* fmap1, const void* fmap2, int in_dtype, int B, int C, int H,
                        int W1, int W2, int64_t row_pitch1, int64_t row_pitch2,
                        const smoe_pyramid_t* grad_pyr, float denom, void* dfmap1, void* dfmap2,
                        smoe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMOE_CORR_H_ */
