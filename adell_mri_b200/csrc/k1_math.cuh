// K1 arithmetic shared by the direct-gather and the TMA-staged paths.
//
// Everything here replays, operation for operation in fp32, what the reference chain
// computes on the CPU (MONAI AffineGrid -> Resample -> ATen grid_sampler_3d, see
// oracle/monai_restated.py and oracle/gather_ref.c for the restatement and its citations):
//   x_a = fma(A[a][3],1, fma(A[a][2],c2, fma(A[a][1],c1, A[a][0]*c0)))   (MKL sgemm k-order)
//   n_a = x_a * (float)(2/max(2,S_a))                                      (Resample norm_coords)
//   u_a = ((n_a + 1) * S_a - 1) / 2                                        (grid_sampler_unnormalize)
// No -use_fast_math, and explicit _rn intrinsics so ptxas cannot contract mul+add pairs.
#pragma once
#include "common.cuh"

struct K1Ctx {
  adell_item it;
  float cg[3];    // (G-1)/2, exact in fp32
  float Sf[3];    // (float)S
  float Sm1[3];   // (float)(S-1)
  int tlo[3];     // max(0, src_vlo)
  int thi[3];     // min(S, src_vhi)
  float pre_s, pre_o;
  float tie;      // nearest fast path: |fast coordinate - rint| above this => bit-faithful replay (0.5 - window)
};

__device__ __forceinline__ void k1_ctx_finish(K1Ctx& c) {
  // called by one thread after the raw item has been copied into c.it (the vector-copy
  // eligibility that used to be decided here is now adell_item.kind, set by adell_aug_prepare)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c.cg[a] = static_cast<float>(c.it.grid_shape[a] - 1) * 0.5f;
    c.Sf[a] = static_cast<float>(c.it.src_shape[a]);
    c.Sm1[a] = static_cast<float>(c.it.src_shape[a] - 1);
    c.tlo[a] = max(0, c.it.src_vlo[a]);
    c.thi[a] = min(c.it.src_shape[a], c.it.src_vhi[a]);
  }
  // The fast (tile-local fp32) coordinate and the reference's own fp32 chain each stay within
  // ~3e-7 * extent of the real coordinate (a handful of roundings at magnitude <= extent), so they
  // disagree by <= ~6e-7 * extent; the tie window is 2e-6 * extent, at least 2e-4 voxel.
  {
    int ext = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) ext = max(ext, max(c.it.src_shape[a], c.it.grid_shape[a]));
    c.tie = 0.5f - fmaxf(2.0e-4f, 2.0e-6f * static_cast<float>(ext));
  }
  if (c.it.flags & ADELL_F_WIN_DEV) {
    // device-side crop window (RandCropByPosNegLabeld centres chosen by adell_posneg_starts): the item was composed
    // for a window at parent index (0,0,0); shift the source and the tensor-map coordinates by the real start
    const int es = c.it.src_dtype == ADELL_F32 ? 4 : (c.it.src_dtype == ADELL_I16 ? 2 : 1);
    int64_t shift = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int w = c.it.win_dev[a];
      const int64_t st = c.it.src_stride[a];
      shift += static_cast<int64_t>(w) * (st < 0 ? -st : st);
      c.it.tmap_off[a] += w;
    }
    c.it.src = static_cast<const char*>(c.it.src) + shift * es;
  }
  if (c.it.flags & ADELL_F_PRE_DEV) {
    c.pre_s = c.it.pre_dev[0];
    c.pre_o = c.it.pre_dev[1];
  } else {
    c.pre_s = c.it.pre_scale;
    c.pre_o = c.it.pre_offset;
  }
}

// ATen compute_coordinates (GridSampler.h) for align_corners=False.
template <int PAD>
__device__ __forceinline__ float k1_pad_coord(float u, float Sf, float Sm1) {
  if (PAD == ADELL_PAD_BORDER) {
    return fminf(Sm1, fmaxf(u, 0.0f));
  } else if (PAD == ADELL_PAD_REFLECTION) {
    // reflect_coordinates(in, twice_low=-1, twice_high=2*size-1): min=-0.5, span=size
    float in = fabsf(__fadd_rn(u, 0.5f));
    float r;
    if (in < Sf) {  // flips == 0 and fmod(in, span) == in
      r = __fadd_rn(in, -0.5f);
    } else {
      float extra = fmodf(in, Sf);
      int flips = static_cast<int>(floorf(__fdiv_rn(in, Sf)));
      r = (flips & 1) ? __fadd_rn(__fsub_rn(Sf, extra), -0.5f) : __fadd_rn(extra, -0.5f);
    }
    return fminf(Sm1, fmaxf(r, 0.0f));
  }
  return u;
}

// Bit-faithful source coordinate of grid index (g0,g1,g2) along axis a (before padding).
__device__ __forceinline__ float k1_coord_exact(const K1Ctx& c, int a, float c0, float c1, float c2) {
  const float* A = c.it.A + 4 * a;
  float x = __fmul_rn(A[0], c0);
  x = __fmaf_rn(A[1], c1, x);
  x = __fmaf_rn(A[2], c2, x);
  x = __fmaf_rn(A[3], 1.0f, x);
  float n = __fmul_rn(x, c.it.nrm[a]);
  return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(n, 1.0f), c.Sf[a]), -1.0f), 0.5f);
}

__device__ __forceinline__ float k1_premap(float v, float s, float o, bool clip, float lo, float hi) {
  v = __fmaf_rn(v, s, o);
  if (clip) v = fminf(hi, fmaxf(v, lo));
  return v;
}
