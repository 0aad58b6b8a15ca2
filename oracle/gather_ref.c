/*
 * CPU ORACLE (C restatement) — test infrastructure only, never the product path.
 *
 * Scalar C restatement of what ONE (sample, key) of the reference chain computes, expressed
 * on the same canonical `adell_item` (include/adell_b200.h) the CUDA path consumes but with
 * HOST pointers.  Every floating-point operation is written out in the order the reference
 * stack executes it on the CPU:
 *   - grid product:   MONAI AffineGrid  `affine @ grid.view(4,-1)`  -> torch CPU mm -> MKL sgemm,
 *                     = fp32 FMA chain in k order (pinned by tests/test_oracle_c_restatement.py
 *                     against the literal torch product in oracle/monai_restated.py)
 *   - normalisation:  MONAI Resample `grid_t[..., i] *= 2.0 / max(2, dim)`              (†)
 *   - un-normalise, padding, nearest/trilinear: ATen GridSampler.cpp grid_sampler_3d_cpu_impl,
 *                     grid_sampler_unnormalize / reflect_coordinates / clip_coordinates
 *                     (align_corners=False) — pinned against F.grid_sample by the same test.
 * Wiring citations (reference): /root/reference/adell_mri/transform_factory/augmentations.py:98-176,
 * 255-301,427-515; transforms.py:143-204,430-499,772-820; utils/utils.py:308-377.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; no -ffast-math).
 * PARITY STATUS: unpinned by reference golden vectors (none exist); pinned
 * against torch's own CPU kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/adell_b200.h"

static float load_src(const void* base, int64_t idx, int dtype) {
  if (dtype == ADELL_F32) return ((const float*)base)[idx];
  if (dtype == ADELL_I16) return (float)((const int16_t*)base)[idx];
  return (float)((const uint8_t*)base)[idx];
}

static float pad_coord(float u, int pad, int S) {
  const float Sf = (float)S, Sm1 = (float)(S - 1);
  if (pad == ADELL_PAD_BORDER) {
    float m = u > 0.0f ? u : 0.0f; /* std::max(in, 0) */
    return Sm1 < m ? Sm1 : m;      /* std::min(size-1, .) */
  }
  if (pad == ADELL_PAD_REFLECTION) {
    /* reflect_coordinates(in, -1, 2*size-1) */
    const float mn = -0.5f, span = Sf;
    float in = fabsf(u - mn);
    float extra = fmodf(in, span);
    int flips = (int)floorf(in / span);
    float r = (flips % 2 == 0) ? (extra + mn) : (span - extra + mn);
    float m = r > 0.0f ? r : 0.0f;
    return Sm1 < m ? Sm1 : m;
  }
  return u;
}

static int tap_in(const adell_item* it, int t0, int t1, int t2) {
  const int t[3] = {t0, t1, t2};
  for (int a = 0; a < 3; ++a) {
    int lo = it->src_vlo[a] > 0 ? it->src_vlo[a] : 0;
    int hi = it->src_vhi[a] < it->src_shape[a] ? it->src_vhi[a] : it->src_shape[a];
    if (t[a] < lo || t[a] >= hi) return 0;
  }
  return 1;
}

static float premap(const adell_item* it, float v, float s, float o) {
  v = fmaf(v, s, o);
  if (it->flags & ADELL_F_CLIP) {
    v = v > it->clip_lo ? v : it->clip_lo;
    v = v < it->clip_hi ? v : it->clip_hi;
  }
  return v;
}

static float tap(const adell_item* it, int t0, int t1, int t2) {
  int64_t idx = t0 * it->src_stride[0] + t1 * it->src_stride[1] + t2 * it->src_stride[2];
  return load_src(it->src, idx, it->src_dtype);
}

static float voxel(const adell_item* it, int g0, int g1, int g2, float pre_s, float pre_o) {
  if (it->flags & ADELL_F_IDENTITY) {
    if (!tap_in(it, g0, g1, g2)) return 0.0f;
    return premap(it, tap(it, g0, g1, g2), pre_s, pre_o);
  }
  const int g[3] = {g0, g1, g2};
  float c[3], u[3];
  for (int a = 0; a < 3; ++a) c[a] = (float)g[a] - (float)(it->grid_shape[a] - 1) * 0.5f;
  for (int a = 0; a < 3; ++a) {
    const float* A = it->A + 4 * a;
    float x = A[0] * c[0];
    x = fmaf(A[1], c[1], x);
    x = fmaf(A[2], c[2], x);
    x = fmaf(A[3], 1.0f, x);
    float n = x * it->nrm[a];
    float uu = ((n + 1.0f) * (float)it->src_shape[a] - 1.0f) / 2.0f;
    u[a] = pad_coord(uu, it->padding, it->src_shape[a]);
  }
  if (it->interp == ADELL_NEAREST) {
    int t0 = (int)nearbyintf(u[0]), t1 = (int)nearbyintf(u[1]), t2 = (int)nearbyintf(u[2]);
    if (!tap_in(it, t0, t1, t2)) return 0.0f;
    return premap(it, tap(it, t0, t1, t2), pre_s, pre_o);
  }
  float f[3];
  int i[3];
  float w[3][2];
  for (int a = 0; a < 3; ++a) {
    f[a] = floorf(u[a]);
    i[a] = (int)f[a];
    w[a][0] = (f[a] + 1.0f) - u[a];
    w[a][1] = u[a] - f[a];
  }
  const int pertap = (it->flags & (ADELL_F_STRICT | ADELL_F_CLIP)) != 0;
  float acc = 0.0f, wsum = 0.0f;
  for (int b0 = 0; b0 < 2; ++b0)
    for (int b1 = 0; b1 < 2; ++b1)
      for (int b2 = 0; b2 < 2; ++b2) {
        int t0 = i[0] + b0, t1 = i[1] + b1, t2 = i[2] + b2;
        float wt = (w[2][b2] * w[1][b1]) * w[0][b0];
        if (!tap_in(it, t0, t1, t2)) continue;
        float v = tap(it, t0, t1, t2);
        if (pertap) {
          v = premap(it, v, pre_s, pre_o);
          acc = acc + v * wt;
        } else {
          acc = fmaf(v, wt, acc);
          wsum += wt;
        }
      }
  if (!pertap) acc = fmaf(pre_s, acc, pre_o * wsum);
  return acc;
}

int adell_ref_gather(const adell_item* items, int n_items) {
  for (int n = 0; n < n_items; ++n) {
    adell_item shifted;
    const adell_item* it = items + n;
    if (it->flags & ADELL_F_WIN_DEV) {
      /* window of a parent volume whose start is read at run time (here: from host memory): the item
         describes the window at start (0,0,0) */
      const int es = it->src_dtype == ADELL_F32 ? 4 : (it->src_dtype == ADELL_I16 ? 2 : 1);
      int64_t shift = 0;
      memcpy(&shifted, it, sizeof(shifted));
      for (int a = 0; a < 3; ++a) {
        const int64_t st = it->src_stride[a];
        shift += (int64_t)it->win_dev[a] * (st < 0 ? -st : st);
        shifted.src_vhi[a] = it->src_shape[a]; /* src_vhi carried the parent's extents */
      }
      shifted.src = (const char*)it->src + shift * es;
      it = &shifted;
    }
    float pre_s = it->pre_scale, pre_o = it->pre_offset;
    if (it->flags & ADELL_F_PRE_DEV) {
      pre_s = it->pre_dev[0];
      pre_o = it->pre_dev[1];
    }
    if (it->flags & ADELL_F_PHILOX) return ADELL_ERR_UNSUPPORTED; /* no reference stream for it */
    const int strict = (it->flags & ADELL_F_STRICT) != 0;
    for (int o0 = 0; o0 < it->out_shape[0]; ++o0)
      for (int o1 = 0; o1 < it->out_shape[1]; ++o1)
        for (int o2 = 0; o2 < it->out_shape[2]; ++o2) {
          const int o[3] = {o0, o1, o2};
          int g[3], ok = 1;
          for (int a = 0; a < 3; ++a) {
            g[a] = it->grid_off[a] + it->grid_sign[a] * o[a];
            if (o[a] < it->out_vlo[a] || o[a] >= it->out_vhi[a]) ok = 0;
          }
          float val = ok ? voxel(it, g[0], g[1], g[2], pre_s, pre_o) : 0.0f;
          if (strict) {
            if (it->post_scale != 1.0f) val = val * it->post_scale;
            if (it->post_offset != 0.0f) val = val + it->post_offset;
          } else {
            val = fmaf(val, it->post_scale, it->post_offset);
          }
          int64_t olin = ((int64_t)o0 * it->out_shape[1] + o1) * it->out_shape[2] + o2;
          if (it->noise) val = val + it->noise[olin];
          it->dst[o0 * it->dst_stride[0] + o1 * it->dst_stride[1] + o2 * it->dst_stride[2]] = val;
        }
  }
  return ADELL_OK;
}

/* ---- statistics ---------------------------------------------------------------------- */
uint32_t adell_ref_key_f32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

void adell_ref_minmax(const float* x, int64_t n, float* out2) {
  float mn = x[0], mx = x[0];
  for (int64_t i = 1; i < n; ++i) {
    if (x[i] < mn) mn = x[i];
    if (x[i] > mx) mx = x[i];
  }
  out2[0] = mn;
  out2[1] = mx;
}

/* y = ((x*m0 - a)/d)*m1*m2 + b, each op rounded to fp32 */
void adell_ref_intensity_map(const float* x, int64_t n, const float* c, int clip, float lo, float hi, float* y) {
  for (int64_t i = 0; i < n; ++i) {
    float v = x[i] * c[0];
    v = v - c[1];
    v = v / c[2];
    v = v * c[3];
    v = v * c[4];
    v = v + c[5];
    if (clip) {
      v = v > lo ? v : lo;
      v = v < hi ? v : hi;
    }
    y[i] = v;
  }
}

/* numpy _lerp on two float32 order statistics with a float64 weight, cast to fp32 */
float adell_ref_percentile_lerp(float a, float b, double t) {
  float diff = b - a;
  double r = (double)a + (double)diff * t;
  if (t >= 0.5) r = (double)b - (double)diff * (1.0 - t);
  return (float)r;
}

static int cmp_f32(const void* p, const void* q) {
  float a = *(const float*)p, b = *(const float*)q;
  return (a > b) - (a < b);
}

/* k-th and (k+1)-th order statistics by full sort (small inputs only) */
int adell_ref_order_stats(const float* x, int64_t n, int64_t k_lo, int64_t k_hi, float* out2) {
  float* tmp = (float*)malloc((size_t)n * sizeof(float));
  if (!tmp) return ADELL_ERR_BAD_ARG;
  memcpy(tmp, x, (size_t)n * sizeof(float));
  qsort(tmp, (size_t)n, sizeof(float), cmp_f32);
  out2[0] = tmp[k_lo];
  out2[1] = tmp[k_hi];
  free(tmp);
  return ADELL_OK;
}
