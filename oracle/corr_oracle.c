/*
 * corr_oracle.c -- CPU restatement of the SMoE-Stereo / RAFT-Stereo 1-D correlation path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the CPU baseline
 * ("port") for bench.py.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (smoe-stereo_b200/) never calls into it.
 *
 * Parity pinning: the reference ships no tests, golden vectors or fixtures for
 * this path, so this restatement is pinned against outputs
 * of the reference's own Python implementation executed in the build container
 * (tests/golden/make_golden.py imports /root/reference/core/corr.py and stores
 * its outputs in tests/golden/ as .npz files; tests/test_oracle_golden.py checks this
 * file against them).
 *
 * Each function cites the reference file:line whose arithmetic it restates.
 * Everything is plain fp32 C (accumulation in fp32 unless noted), OpenMP over
 * independent (batch,row) units.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* floor(w / 2) chain used by the pyramid: F.avg_pool2d(corr,[1,2],stride=[1,2])
 * drops the last column of an odd-width level (core/corr.py:122-125, :40-42). */
ORACLE_API int oracle_level_width(int w2, int level) {
  for (int i = 0; i < level; ++i) w2 /= 2;
  return w2;
}

/*
 * All-pairs 1-D correlation volume.
 *   corr[b,h,w1,w2] = (1/sqrt(C)) * sum_c f1[b,c,h,w1] * f2[b,c,h,w2]
 * Reference: core/corr.py:148-156 (CorrBlock1D.corr) and :53-61 (CorrBlockFast1D.corr):
 *   einsum('aijk,aijh->ajkh') then division by sqrt(D).
 * f1: (B,C,H,W1) contiguous, f2: (B,C,H,W2) contiguous, out: (B,H,W1,W2) contiguous.
 * The division is a true fp32 division, as in the reference (:156).
 */
ORACLE_API void oracle_corr_volume(const float* f1, const float* f2, float* out,
                                   int B, int C, int H, int W1, int W2) {
  const float denom = sqrtf((float)C);
  const long rows = (long)B * H;
#pragma omp parallel for schedule(dynamic, 1)
  for (long bh = 0; bh < rows; ++bh) {
    const int b = (int)(bh / H), h = (int)(bh % H);
    float* o = out + (size_t)bh * W1 * W2;
    memset(o, 0, sizeof(float) * (size_t)W1 * W2);
    /* register-block 4 rows of the output; the inner loop over w2 vectorises */
    for (int w1 = 0; w1 < W1; w1 += 4) {
      const int nr = (W1 - w1 < 4) ? (W1 - w1) : 4;
      for (int c = 0; c < C; ++c) {
        const float* a = f1 + (((size_t)b * C + c) * H + h) * W1 + w1;
        const float* bb = f2 + (((size_t)b * C + c) * H + h) * W2;
        if (nr == 4) {
          const float a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3];
          float* o0 = o + (size_t)(w1 + 0) * W2;
          float* o1 = o + (size_t)(w1 + 1) * W2;
          float* o2 = o + (size_t)(w1 + 2) * W2;
          float* o3 = o + (size_t)(w1 + 3) * W2;
          for (int w2 = 0; w2 < W2; ++w2) {
            const float v = bb[w2];
            o0[w2] += a0 * v;
            o1[w2] += a1 * v;
            o2[w2] += a2 * v;
            o3[w2] += a3 * v;
          }
        } else {
          for (int r = 0; r < nr; ++r) {
            const float ar = a[r];
            float* orow = o + (size_t)(w1 + r) * W2;
            for (int w2 = 0; w2 < W2; ++w2) orow[w2] += ar * bb[w2];
          }
        }
      }
    }
    for (size_t i = 0; i < (size_t)W1 * W2; ++i) o[i] = o[i] / denom;
  }
}

/*
 * One pyramid step: out[r, j] = (in[r, 2j] + in[r, 2j+1]) / 2 for j < floor(w/2).
 * Reference: core/corr.py:122-125 / :40-42, F.avg_pool2d(corr, [1,2], stride=[1,2]).
 * in: (rows, w) contiguous, out: (rows, w/2) contiguous.
 */
ORACLE_API void oracle_pool_level(const float* in, float* out, long rows, int w) {
  const int wo = w / 2;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    const float* i_ = in + (size_t)r * w;
    float* o_ = out + (size_t)r * wo;
    for (int j = 0; j < wo; ++j) o_[j] = (i_[2 * j] + i_[2 * j + 1]) / 2.0f;
  }
}

/*
 * Single-level lookup ("corr_sampler.forward").
 * Reference: sampler/sampler_kernel.cu:19-60.
 *   x0 = coords[n,0,y,x]; dx = x0 - floor(x0);
 *   for i in 0..2r+1: x1 = floor(x0) - r + i; if 0 <= x1 < W2:
 *       s = volume[n,y,x,x1]; out[i-1] += s*dx (i>0); out[i] += s*(1-dx) (i<2r+1)
 * volume: (B,H,W1,W2), coords: (B,coords_ch,H,W1) (only channel 0 is read),
 * out: (B,2r+1,H,W1).  Coordinates whose floor is not representable contribute 0.
 */
ORACLE_API void oracle_lookup_level(const float* volume, const float* coords, int coords_ch,
                                    float* out, int B, int H, int W1, int W2, int radius,
                                    float coord_scale) {
  const int rd = 2 * radius + 1;
  const long rows = (long)B * H;
#pragma omp parallel for schedule(static)
  for (long bh = 0; bh < rows; ++bh) {
    const int b = (int)(bh / H), y = (int)(bh % H);
    for (int x = 0; x < W1; ++x) {
      const float x0 = coords[(((size_t)b * coords_ch + 0) * H + y) * W1 + x] * coord_scale;
      const float fl = floorf(x0);
      const float dx = x0 - fl;
      float acc[64];
      for (int k = 0; k < rd; ++k) acc[k] = 0.0f;
      if (fl > -1.0e8f && fl < 1.0e8f) {
        const int base = (int)fl - radius;
        const float* row = volume + ((size_t)bh * W1 + x) * W2;
        for (int i = 0; i < rd + 1; ++i) {
          const int x1 = base + i;
          if (x1 >= 0 && x1 < W2) {
            const float s = row[x1];
            if (i > 0) acc[i - 1] += s * dx;
            if (i < rd) acc[i] += s * (1.0f - dx);
          }
        }
      }
      for (int k = 0; k < rd; ++k)
        out[(((size_t)b * rd + k) * H + y) * W1 + x] = acc[k];
    }
  }
}

/*
 * Single-level lookup backward ("corr_sampler.backward").
 * Reference: sampler/sampler_kernel.cu:63-105.
 *   g = grad[i-1]*dx (i>0) + grad[i]*(1-dx) (i<2r+1); volume_grad[n,y,x,x1] += g
 * If accumulate == 0 the gradient is zero-filled first (reference :148 zeros_like).
 */
ORACLE_API void oracle_lookup_level_bwd(const float* coords, int coords_ch, const float* grad_out,
                                        float* volume_grad, int B, int H, int W1, int W2,
                                        int radius, float coord_scale, int accumulate) {
  const int rd = 2 * radius + 1;
  const long rows = (long)B * H;
#pragma omp parallel for schedule(static)
  for (long bh = 0; bh < rows; ++bh) {
    const int b = (int)(bh / H), y = (int)(bh % H);
    for (int x = 0; x < W1; ++x) {
      float* row = volume_grad + ((size_t)bh * W1 + x) * W2;
      if (!accumulate) memset(row, 0, sizeof(float) * (size_t)W2);
      const float x0 = coords[(((size_t)b * coords_ch + 0) * H + y) * W1 + x] * coord_scale;
      const float fl = floorf(x0);
      const float dx = x0 - fl;
      if (!(fl > -1.0e8f && fl < 1.0e8f)) continue;
      const int base = (int)fl - radius;
      for (int i = 0; i < rd + 1; ++i) {
        const int x1 = base + i;
        if (x1 >= 0 && x1 < W2) {
          float g = 0.0f;
          if (i > 0) g += grad_out[(((size_t)b * rd + (i - 1)) * H + y) * W1 + x] * dx;
          if (i < rd) g += grad_out[(((size_t)b * rd + i) * H + y) * W1 + x] * (1.0f - dx);
          row[x1] += g;
        }
      }
    }
  }
}

/*
 * Gradient of the pooling chain folded back onto level 0:
 *   g0[j] += 0.5 * g1[j/2]  (only for j < 2*floor(w0/2)), recursively.
 * Reference: autograd of core/corr.py:122-125 (avg_pool2d backward).
 * g_levels[i]: (rows, w_i) contiguous; on return g_levels[0] holds the total.
 */
ORACLE_API void oracle_fold_pyramid_grad(float* g0, const float* g1, long rows, int w0) {
  const int w1 = w0 / 2;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    float* a = g0 + (size_t)r * w0;
    const float* bsrc = g1 + (size_t)r * w1;
    for (int j = 0; j < w1; ++j) {
      const float v = 0.5f * bsrc[j];
      a[2 * j] += v;
      a[2 * j + 1] += v;
    }
  }
}

/*
 * Volume-build backward: given G = dL/dcorr (B,H,W1,W2),
 *   df1[b,c,h,w1] = (1/sqrt(C)) * sum_w2 G[b,h,w1,w2] * f2[b,c,h,w2]
 *   df2[b,c,h,w2] = (1/sqrt(C)) * sum_w1 G[b,h,w1,w2] * f1[b,c,h,w1]
 * Reference: autograd of core/corr.py:154-156 (einsum then divide).
 */
ORACLE_API void oracle_corr_volume_bwd(const float* f1, const float* f2, const float* G,
                                       float* df1, float* df2, int B, int C, int H, int W1,
                                       int W2) {
  const float denom = sqrtf((float)C);
  const long units = (long)B * C * H;
#pragma omp parallel for schedule(static)
  for (long u = 0; u < units; ++u) {
    const int h = (int)(u % H);
    const int b = (int)(u / ((long)H * C));
    const float* g = G + ((size_t)b * H + h) * W1 * W2;
    const float* a = f1 + (size_t)u * W1;
    const float* bb = f2 + (size_t)u * W2;
    float* da = df1 + (size_t)u * W1;
    float* db = df2 + (size_t)u * W2;
    for (int w2 = 0; w2 < W2; ++w2) db[w2] = 0.0f;
    for (int w1 = 0; w1 < W1; ++w1) {
      const float* grow = g + (size_t)w1 * W2;
      const float av = a[w1];
      float acc = 0.0f;
      for (int w2 = 0; w2 < W2; ++w2) {
        acc += grow[w2] * bb[w2];
        db[w2] += grow[w2] * av;
      }
      da[w1] = acc / denom;
    }
    for (int w2 = 0; w2 < W2; ++w2) db[w2] = db[w2] / denom;
  }
}

/*
 * The whole hot path for one batch, as the reference runs it on CPU
 * (core/corr.py:110-146): build volume, pool num_levels-1 times, then `iters`
 * lookups over all levels; output channel = level*(2r+1)+tap (core/corr.py:143-146).
 * coords: (iters, B, coords_ch, H, W1); out: (iters, B, L*(2r+1), H, W1);
 * pyramid scratch: caller-provided, sum_i rows*W2_i floats.
 * Used by bench.py as the timed CPU baseline.
 */
ORACLE_API void oracle_corr_path(const float* f1, const float* f2, const float* coords,
                                 int coords_ch, float* pyramid, float* level_out, float* out,
                                 int B, int C, int H, int W1, int W2, int num_levels, int radius,
                                 int iters) {
  const long rows = (long)B * H * W1;
  const int rd = 2 * radius + 1;
  float* lv[8] = {0};
  int lw[8] = {0};
  if (num_levels < 1 || num_levels > 8) return;
  size_t off = 0;
  for (int i = 0; i < num_levels; ++i) {
    lv[i] = pyramid + off;
    lw[i] = oracle_level_width(W2, i);
    off += (size_t)rows * lw[i];
  }
  oracle_corr_volume(f1, f2, lv[0], B, C, H, W1, W2);
  for (int i = 1; i < num_levels; ++i) oracle_pool_level(lv[i - 1], lv[i], rows, lw[i - 1]);
  const size_t coords_sz = (size_t)B * coords_ch * H * W1;
  const size_t out_sz = (size_t)B * num_levels * rd * H * W1;
  const size_t plane = (size_t)H * W1;
  for (int t = 0; t < iters; ++t) {
    float scale = 1.0f;
    for (int i = 0; i < num_levels; ++i) {
      oracle_lookup_level(lv[i], coords + t * coords_sz, coords_ch, level_out, B, H, W1, lw[i],
                          radius, scale);
      for (int b = 0; b < B; ++b)
        memcpy(out + t * out_sz + ((size_t)b * num_levels * rd + (size_t)i * rd) * plane,
               level_out + (size_t)b * rd * plane, sizeof(float) * rd * plane);
      scale *= 0.5f;
    }
  }
}
