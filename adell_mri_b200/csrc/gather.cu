// K1 — fused gather: affine ∘ flip ∘ crop ∘ pad ∘ intensity ∘ noise ∘ collate in one pass.
//
// Replaces, per (sample, key), the reference's chain of MONAI dictionary transforms
// (RandAffined -> RandFlipd -> CenterSpatialCropd/SpatialPadd/RandSpatialCropd ->
//  RandGaussianNoised/RandScale/ShiftIntensityd -> ConcatItemsd -> safe_collate;
//  /root/reference/adell_mri/transform_factory/augmentations.py:98-176,255-301,427-515,
//  /root/reference/adell_mri/utils/utils.py:308-377), each of which materialises a full
//  copy of the volume on the CPU.  Here every output voxel is produced once and written
//  straight into the collated [B,C,H,W,D] batch.
//
// One CTA = one 16x16x16 output tile of one item.  Three code paths, chosen per tile
// (block-uniform):
//   STAGED  the tile's source footprint (a box whose extents depend only on the item's matrix)
//           is fetched by ONE TMA tensor copy (cp.async.bulk.tensor.3d, zero fill outside the
//           volume = "zeros" padding for free) into shared memory; border/reflection halos are
//           completed in shared memory; taps are then LDS with compile-time-free bank spread.
//           Trilinear uses tile-local incremental coordinates + nested lerps (<=1e-4 contract);
//           nearest uses the same coordinates and re-evaluates the bit-faithful MONAI/ATen
//           chain only when a coordinate is within 1e-3 of a rounding tie => masks stay bit-exact.
//           ADELL_F_STRICT / ADELL_F_CLIP items take the bit-faithful chain for every voxel.
//   COPY    identity items (no resample fired): 128-bit vectorised flip/crop copy.
//   DIRECT  generic fallback: taps fetched with read-only global loads (int16/uint8 sources,
//           unaligned or oversized footprints, multiply-reflected coordinates, pad bands).
#include <cuda.h>
#include <stdlib.h>

#include "k1_math.cuh"

namespace {

constexpr int K1_CWARPS = 16;                  // consumer warps per CTA
constexpr int K1_CTHREADS = 32 * K1_CWARPS;    // consumer threads: one 16x16x16 tile = 8 voxels each
constexpr int K1_PWARPS = 4;                   // producer warps: tile k is prepared and issued by warp k % 4
constexpr int K1_THREADS = K1_CTHREADS + 32 * K1_PWARPS;
constexpr int K1_T = 16;                       // tile edge
constexpr int K1_MAX_STAGES = 4;
constexpr int K1_SMEM_BUDGET = 224 * 1024;     // dynamic shared memory per persistent CTA (1 CTA per SM)
constexpr int K1_MAX_BOX_BYTES = 100 * 1024;   // staged footprint limit (>= 2 CTAs per SM)
constexpr double K1_EPS = 1e-3;                // coordinate slack of the fast path / tie window

enum { MODE_DIRECT = 0, MODE_STAGED = 1, MODE_ZERO = 2, MODE_COPY = 3 };

struct K1Tile {
  int mode;
  int item;         // index of the tile's item (descriptor address for the TMA issue)
  int o0[3];        // tile origin (output index space)
  int box[3];       // staged box extents, axes 0,1,2
  int mconst[3];    // box-local memory index = msign*t + mconst
  int msign[3];
  int lo_t[3], hi_t[3];
  float V0[3];      // fast path: coordinate of the tile-origin voxel — box-local memory order for
  float Dm[3][3];   // plain axes, absolute source index for axes in rmask; and its derivative
  int rmask;        // axes whose coordinates leave [0,S): border / reflection applied per voxel
  float rA[3], rB[3];  // box-local index = rA*u' + rB for the axes in rmask
  int all_valid;    // every tap of the tile lies inside the valid source region
};

// ------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// The descriptor lives in global memory and was written by a host copy: the tensormap proxy
// must acquire it before the TMA unit reads it (CUDA programming guide, "tensor map in global
// memory").
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// ------------------------------------------------------------------------- tiling policy
__host__ __device__ inline void k1_tile_counts(const int32_t* O, int& n0, int& n1, int& n2) {
  n0 = (O[0] + K1_T - 1) / K1_T;
  n1 = (O[1] + K1_T - 1) / K1_T;
  n2 = (O[2] + K1_T - 1) / K1_T;
}

// ------------------------------------------------------------------------- tap fetchers
struct GlobalTaps {
  template <int DT>
  static __device__ __forceinline__ float get(const K1Ctx& c, const K1Tile&, const float*, int t0, int t1, int t2) {
    int64_t idx = t0 * c.it.src_stride[0] + t1 * c.it.src_stride[1] + t2 * c.it.src_stride[2];
    if (DT == ADELL_F32) return adell_load_src_t<ADELL_F32>(c.it.src, idx);
    return adell_load_src(c.it.src, idx, c.it.src_dtype);
  }
};
struct SmemTaps {
  template <int DT>
  static __device__ __forceinline__ float get(const K1Ctx&, const K1Tile& t, const float* box, int t0, int t1, int t2) {
    int m0 = t.msign[0] * t0 + t.mconst[0], m1 = t.msign[1] * t1 + t.mconst[1], m2 = t.msign[2] * t2 + t.mconst[2];
    return box[(m0 * t.box[1] + m1) * t.box[2] + m2];
  }
};

__device__ __forceinline__ bool k1_in(const K1Ctx& c, int t0, int t1, int t2) {
  return (t0 >= c.tlo[0]) & (t0 < c.thi[0]) & (t1 >= c.tlo[1]) & (t1 < c.thi[1]) & (t2 >= c.tlo[2]) &
         (t2 < c.thi[2]);
}

__device__ __forceinline__ float k1_pad_rt(float u, int pad, float Sf, float Sm1) {
  if (pad == ADELL_PAD_BORDER) return k1_pad_coord<ADELL_PAD_BORDER>(u, Sf, Sm1);
  if (pad == ADELL_PAD_REFLECTION) return k1_pad_coord<ADELL_PAD_REFLECTION>(u, Sf, Sm1);
  return u;
}

// Bit-faithful (MONAI/ATen operation order) value of one output voxel.  PERTAP: pre map (and
// clip) on every tap, ATen-order mul/add accumulation (ADELL_F_STRICT / ADELL_F_CLIP); otherwise
// fma accumulation of sum(w*v), sum(w_valid) and one pre map at the end.
template <class Taps, int DT, bool PERTAP>
__device__ __forceinline__ float k1_exact_voxel(const K1Ctx& c, const K1Tile& tl, const float* box, int g0, int g1, int g2) {
  const float c0 = static_cast<float>(g0) - c.cg[0];
  const float c1 = static_cast<float>(g1) - c.cg[1];
  const float c2 = static_cast<float>(g2) - c.cg[2];
  const int pad = c.it.padding;
  float u0 = k1_pad_rt(k1_coord_exact(c, 0, c0, c1, c2), pad, c.Sf[0], c.Sm1[0]);
  float u1 = k1_pad_rt(k1_coord_exact(c, 1, c0, c1, c2), pad, c.Sf[1], c.Sm1[1]);
  float u2 = k1_pad_rt(k1_coord_exact(c, 2, c0, c1, c2), pad, c.Sf[2], c.Sm1[2]);
  const bool clip = (c.it.flags & ADELL_F_CLIP) != 0;

  if (c.it.interp == ADELL_NEAREST) {
    int t0 = __float2int_rn(u0), t1 = __float2int_rn(u1), t2 = __float2int_rn(u2);
    if (!k1_in(c, t0, t1, t2)) return 0.0f;
    return k1_premap(Taps::template get<DT>(c, tl, box, t0, t1, t2), c.pre_s, c.pre_o, clip, c.it.clip_lo, c.it.clip_hi);
  }
  float f0 = floorf(u0), f1 = floorf(u1), f2 = floorf(u2);
  int i0 = static_cast<int>(f0), i1 = static_cast<int>(f1), i2 = static_cast<int>(f2);
  // ATen: (ix_tnw + 1) - ix  and  ix - ix_tnw, integers converted to float
  float w0[2] = {__fsub_rn(__fadd_rn(f0, 1.0f), u0), __fsub_rn(u0, f0)};
  float w1[2] = {__fsub_rn(__fadd_rn(f1, 1.0f), u1), __fsub_rn(u1, f1)};
  float w2[2] = {__fsub_rn(__fadd_rn(f2, 1.0f), u2), __fsub_rn(u2, f2)};
  float acc = 0.0f, wsum = 0.0f;
#pragma unroll
  for (int b0 = 0; b0 < 2; ++b0) {
#pragma unroll
    for (int b1 = 0; b1 < 2; ++b1) {
#pragma unroll
      for (int b2 = 0; b2 < 2; ++b2) {
        int t0 = i0 + b0, t1 = i1 + b1, t2 = i2 + b2;
        // weight = (wx * wy) * wz with x = axis 2, y = axis 1, z = axis 0 (ATen naming)
        float w = __fmul_rn(__fmul_rn(w2[b2], w1[b1]), w0[b0]);
        if (k1_in(c, t0, t1, t2)) {
          float v = Taps::template get<DT>(c, tl, box, t0, t1, t2);
          if (PERTAP) {
            v = k1_premap(v, c.pre_s, c.pre_o, clip, c.it.clip_lo, c.it.clip_hi);
            acc = __fadd_rn(acc, __fmul_rn(v, w));
          } else {
            acc = fmaf(v, w, acc);
            wsum += w;
          }
        }
      }
    }
  }
  if (!PERTAP) acc = fmaf(c.pre_s, acc, c.pre_o * wsum);
  return acc;
}

template <int DT>
__device__ __forceinline__ float k1_identity_voxel(const K1Ctx& c, int g0, int g1, int g2) {
  if (!k1_in(c, g0, g1, g2)) return 0.0f;
  const bool clip = (c.it.flags & ADELL_F_CLIP) != 0;
  K1Tile dummy;
  return k1_premap(GlobalTaps::get<DT>(c, dummy, nullptr, g0, g1, g2), c.pre_s, c.pre_o, clip, c.it.clip_lo, c.it.clip_hi);
}

// post intensity map, noise, store
__device__ __forceinline__ void k1_finish(const adell_item& it, float val, int o0, int o1, int o2, bool strict) {
  if (strict) {
    if (it.post_scale != 1.0f) val = __fmul_rn(val, it.post_scale);
    if (it.post_offset != 0.0f) val = __fadd_rn(val, it.post_offset);
  } else {
    val = fmaf(val, it.post_scale, it.post_offset);
  }
  if (it.noise != nullptr || (it.flags & ADELL_F_PHILOX)) {
    const int64_t olin = (static_cast<int64_t>(o0) * it.out_shape[1] + o1) * it.out_shape[2] + o2;
    if (it.noise != nullptr) val = __fadd_rn(val, __ldg(it.noise + olin));
    if (it.flags & ADELL_F_PHILOX)
      val = fmaf(it.noise_std, adell_philox_normal(it.philox_seed, it.philox_offset + olin), val);
  }
  it.dst[o0 * it.dst_stride[0] + o1 * it.dst_stride[1] + o2 * it.dst_stride[2]] = val;
}

// Thread -> voxel mapping shared by every path: lane = (dj parity, dk), consumer warp = i plane.
template <class F>
__device__ __forceinline__ void k1_for_each_voxel(const K1Tile& tl, const adell_item& it, F&& body) {
  const int dk = threadIdx.x & 15, jj = (threadIdx.x >> 4) & 1, di = threadIdx.x >> 5;
  const int o2 = tl.o0[2] + dk, o0 = tl.o0[0] + di;
  if (o2 >= it.out_shape[2] || o0 >= it.out_shape[0]) return;
#pragma unroll 2
  for (int s = 0; s < 8; ++s) {
    const int dj = 2 * s + jj, o1 = tl.o0[1] + dj;
    if (o1 >= it.out_shape[1]) break;
    body(di, dj, dk, o0, o1, o2);
  }
}

template <class Taps, int DT, bool PERTAP, bool IDENT>
__device__ __forceinline__ void k1_tile_exact(const K1Ctx& c, const K1Tile& tl, const float* box) {
  const adell_item& it = c.it;
  const bool strict = (it.flags & ADELL_F_STRICT) != 0;
  k1_for_each_voxel(tl, it, [&](int di, int dj, int dk, int o0, int o1, int o2) {
    float val = 0.0f;
    const bool ov = (o0 >= it.out_vlo[0]) & (o0 < it.out_vhi[0]) & (o1 >= it.out_vlo[1]) & (o1 < it.out_vhi[1]) &
                    (o2 >= it.out_vlo[2]) & (o2 < it.out_vhi[2]);
    if (ov) {
      const int g0 = it.grid_off[0] + it.grid_sign[0] * o0;
      const int g1 = it.grid_off[1] + it.grid_sign[1] * o1;
      const int g2 = it.grid_off[2] + it.grid_sign[2] * o2;
      val = IDENT ? k1_identity_voxel<DT>(c, g0, g1, g2) : k1_exact_voxel<Taps, DT, PERTAP>(c, tl, box, g0, g1, g2);
    }
    k1_finish(it, val, o0, o1, o2, strict);
  });
}

// Out of line on purpose: the exact / direct variants are cold next to the staged fast loops, and
// inlining all of them made the kernel ~400 KB of SASS (instruction-cache misses dominated).
template <class Taps, bool IDENT>
__device__ __noinline__ void k1_tile_exact_dispatch(const K1Ctx& c, const K1Tile& tl, const float* box) {
  const bool pertap = (c.it.flags & (ADELL_F_STRICT | ADELL_F_CLIP)) != 0;
  if (c.it.src_dtype == ADELL_F32) {
    if (pertap) k1_tile_exact<Taps, ADELL_F32, true, IDENT>(c, tl, box);
    else k1_tile_exact<Taps, ADELL_F32, false, IDENT>(c, tl, box);
  } else {
    if (pertap) k1_tile_exact<Taps, -1, true, IDENT>(c, tl, box);
    else k1_tile_exact<Taps, -1, false, IDENT>(c, tl, box);
  }
}

// ------------------------------------------------------------------------- staged fast paths
// Tile-local incremental coordinates: v_a = V0_a + Dm_a0*di + Dm_a1*dj + Dm_a2*dk is the
// box-local (memory order) source coordinate; floor/frac/lerp directly on it.  Everything the
// inner loop needs is copied into registers first: the output stores go through a generic
// pointer, so the compiler would otherwise reload every shared-memory field per voxel.
struct __align__(16) K1Fast {
  float V0[3], D0[3], D1[3], D2[3];
  float Sf[3], Sm1[3], rA[3], rB[3];
  float p0f, p1f, gain, bias, post_o, noise_std;
  int p0, p1;
  int n0, n1, n2;        // voxels of the tile along each axis
  int vlo[3], vhi[3];    // valid output range, tile-local
  int rmask, pad;        // per-voxel border / reflection handling (axes in rmask)
  int philox, padded;
  float* dst;            // tile origin in the destination
  const float* noise;    // tile origin in the noise tensor (or null)
  int64_t ds0, ds1, ds2;
  int64_t ns0, ns1;      // noise strides (contiguous [O0,O1,O2])
  uint64_t olin0;        // linear output index of the tile origin (Philox counter)
  uint64_t philox_seed, philox_offset;
};

// ATen compute_coordinates on the fast coordinate (same formulas as k1_pad_coord, plain fp32).
__device__ __noinline__ float k1_fast_reflect_far(float x, float Sf) {
  const float n = floorf(x / Sf);
  x = fmaf(-n, Sf, x);
  if (static_cast<int>(n) & 1) x = Sf - x;
  return x;
}
__device__ __forceinline__ float k1_fast_pad(float u, int pad, float Sf, float Sm1) {
  if (pad == ADELL_PAD_REFLECTION) {
    float x = fabsf(u + 0.5f);
    if (x >= Sf) x = k1_fast_reflect_far(x, Sf);
    u = x - 0.5f;
  }
  return fminf(Sm1, fmaxf(u, 0.0f));
}

// bit-faithful replay for one nearest voxel (tie window of the fast path); cold
__device__ __noinline__ float k1_exact_nearest_smem(const K1Ctx& c, const K1Tile& tl, const float* box, int di, int dj, int dk) {
  const adell_item& it = c.it;
  const int g0 = it.grid_off[0] + it.grid_sign[0] * (tl.o0[0] + di);
  const int g1 = it.grid_off[1] + it.grid_sign[1] * (tl.o0[1] + dj);
  const int g2 = it.grid_off[2] + it.grid_sign[2] * (tl.o0[2] + dk);
  return fmaf(k1_exact_voxel<SmemTaps, ADELL_F32, false>(c, tl, box, g0, g1, g2), it.post_scale, it.post_offset);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Filled by the producer lane once per staged tile; the consumers copy it into registers with
// 128-bit shared loads (the output stores go through a generic pointer, so anything left in
// shared memory would be reloaded per voxel).
__device__ __forceinline__ void k1_fast_fill(const K1Ctx& c, const K1Tile& tl, K1Fast& f) {
  const adell_item& it = c.it;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    f.V0[a] = tl.V0[a]; f.D0[a] = tl.Dm[a][0]; f.D1[a] = tl.Dm[a][1]; f.D2[a] = tl.Dm[a][2];
    f.Sf[a] = c.Sf[a]; f.Sm1[a] = c.Sm1[a]; f.rA[a] = tl.rA[a]; f.rB[a] = tl.rB[a];
    f.vlo[a] = it.out_vlo[a] - tl.o0[a];
    f.vhi[a] = it.out_vhi[a] - tl.o0[a];
  }
  f.p1 = tl.box[2]; f.p0 = tl.box[1] * tl.box[2];
  f.p1f = static_cast<float>(f.p1); f.p0f = static_cast<float>(f.p0);
  f.gain = c.pre_s * it.post_scale;
  f.bias = fmaf(c.pre_o, it.post_scale, it.post_offset);
  f.post_o = it.post_offset;
  f.noise_std = it.noise_std;
  f.n0 = min(K1_T, it.out_shape[0] - tl.o0[0]);
  f.n1 = min(K1_T, it.out_shape[1] - tl.o0[1]);
  f.n2 = min(K1_T, it.out_shape[2] - tl.o0[2]);
  bool padded = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) padded = padded || f.vlo[a] > 0 || f.vhi[a] < (a == 0 ? f.n0 : a == 1 ? f.n1 : f.n2);
  f.padded = padded ? 1 : 0;
  f.ds0 = it.dst_stride[0]; f.ds1 = it.dst_stride[1]; f.ds2 = it.dst_stride[2];
  f.dst = it.dst + tl.o0[0] * it.dst_stride[0] + tl.o0[1] * it.dst_stride[1] + tl.o0[2] * it.dst_stride[2];
  f.ns1 = it.out_shape[2]; f.ns0 = static_cast<int64_t>(it.out_shape[1]) * it.out_shape[2];
  f.olin0 = (static_cast<uint64_t>(tl.o0[0]) * it.out_shape[1] + tl.o0[1]) * it.out_shape[2] + tl.o0[2];
  f.noise = it.noise ? it.noise + f.olin0 : nullptr;
  f.philox = (it.flags & ADELL_F_PHILOX) ? 1 : 0;
  f.philox_seed = it.philox_seed; f.philox_offset = it.philox_offset;
  f.rmask = tl.rmask; f.pad = it.padding;
}

// Hot register image of a staged tile (everything else stays in the shared-memory K1Fast and is
// only touched on the rare padded / noise paths).
struct K1Hot {
  float D1[3], P[3];
  float Sf[3], Sm1[3], rA[3], rB[3];
  float p0f, p1f, gain, bias;
  int rmask, pad, n1;
  int64_t ds1;
  float* drow;
  bool cold;  // padded output region or noise: take the slow store
};

__device__ __forceinline__ K1Hot k1_hot_load(const K1Fast& f, int di, int dk) {
  K1Hot h;
  const float fk = static_cast<float>(dk), fi = static_cast<float>(di);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    h.D1[a] = f.D1[a];
    h.P[a] = fmaf(f.D0[a], fi, fmaf(f.D2[a], fk, f.V0[a]));
    h.Sf[a] = f.Sf[a]; h.Sm1[a] = f.Sm1[a]; h.rA[a] = f.rA[a]; h.rB[a] = f.rB[a];
  }
  h.p0f = f.p0f; h.p1f = f.p1f; h.gain = f.gain; h.bias = f.bias;
  h.rmask = f.rmask; h.pad = f.pad; h.n1 = f.n1;
  h.ds1 = f.ds1;
  h.drow = f.dst + di * f.ds0 + dk * f.ds2;
  h.cold = f.padded || f.noise != nullptr || f.philox;
  return h;
}

__device__ __forceinline__ void k1_fast_coords(const K1Hot& h, int dj, float& v0, float& v1, float& v2) {
  const float fj = static_cast<float>(dj);
  v0 = fmaf(h.D1[0], fj, h.P[0]); v1 = fmaf(h.D1[1], fj, h.P[1]); v2 = fmaf(h.D1[2], fj, h.P[2]);
  if (h.rmask) {  // block-uniform: some axis leaves [0,S) inside this tile
    if (h.rmask & 1) v0 = fmaf(h.rA[0], k1_fast_pad(v0, h.pad, h.Sf[0], h.Sm1[0]), h.rB[0]);
    if (h.rmask & 2) v1 = fmaf(h.rA[1], k1_fast_pad(v1, h.pad, h.Sf[1], h.Sm1[1]), h.rB[1]);
    if (h.rmask & 4) v2 = fmaf(h.rA[2], k1_fast_pad(v2, h.pad, h.Sf[2], h.Sm1[2]), h.rB[2]);
  }
}

// rare: output pad band (SpatialPadd after the resample), injected or Philox noise
__device__ __noinline__ void k1_cold_store(const K1Fast& f, float* p, int di, int dj, int dk, float val) {
  if (f.padded) {
    const bool ov = (di >= f.vlo[0]) & (di < f.vhi[0]) & (dj >= f.vlo[1]) & (dj < f.vhi[1]) & (dk >= f.vlo[2]) & (dk < f.vhi[2]);
    if (!ov) val = f.post_o;
  }
  const int64_t rel = di * f.ns0 + dj * f.ns1 + dk;
  if (f.noise != nullptr) val = __fadd_rn(val, __ldg(f.noise + rel));
  if (f.philox) val = fmaf(f.noise_std, adell_philox_normal(f.philox_seed, f.philox_offset + f.olin0 + rel), val);
  *p = val;
}

__device__ __forceinline__ void k1_fast_store(const K1Fast& f, const K1Hot& h, int di, int dj, int dk, float val) {
  float* p = h.drow + dj * h.ds1;
  if (h.cold) k1_cold_store(f, p, di, dj, dk, val);
  else *p = val;
}

__device__ __forceinline__ void k1_tile_staged_nearest(const K1Ctx& c, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box) {
  const int dk = threadIdx.x & 15, jj = (threadIdx.x >> 4) & 1, di = threadIdx.x >> 5;
  if (dk >= f.n2 || di >= f.n0) return;
  const K1Hot h = k1_hot_load(f, di, dk);
  const float tie = 0.5f - static_cast<float>(K1_EPS);
#pragma unroll 2
  for (int s = 0; s < 8; ++s) {
    const int dj = 2 * s + jj;
    if (dj >= h.n1) break;
    float v0, v1, v2;
    k1_fast_coords(h, dj, v0, v1, v2);
    const float n0 = rintf(v0), n1 = rintf(v1), n2 = rintf(v2);
    float val;
    if (fabsf(v0 - n0) > tie || fabsf(v1 - n1) > tie || fabsf(v2 - n2) > tie)
      val = k1_exact_nearest_smem(c, tl, box, di, dj, dk);  // within 1e-3 of a rounding tie
    else
      val = fmaf(box[__float2int_rn(fmaf(n0, h.p0f, fmaf(n1, h.p1f, n2)))], h.gain, h.bias);
    k1_fast_store(f, h, di, dj, dk, val);
  }
}

struct K1Vox {
  float r0, r1, r2;
  uint32_t a;  // shared-memory byte address of tap (0,0,0)
};

__device__ __forceinline__ K1Vox k1_fast_vox(const K1Hot& h, int dj, uint32_t box_addr) {
  float v0, v1, v2;
  k1_fast_coords(h, dj, v0, v1, v2);
  const float f0 = floorf(v0), f1 = floorf(v1), f2 = floorf(v2);
  K1Vox x;
  x.r0 = v0 - f0; x.r1 = v1 - f1; x.r2 = v2 - f2;
  x.a = box_addr + 4u * static_cast<uint32_t>(__float2int_rn(fmaf(f0, h.p0f, fmaf(f1, h.p1f, f2))));
  return x;
}

__device__ __forceinline__ float k1_lerp8(const K1Vox& x, const float* t) {
  const float x00 = fmaf(x.r2, t[1] - t[0], t[0]), x01 = fmaf(x.r2, t[3] - t[2], t[2]);
  const float x10 = fmaf(x.r2, t[5] - t[4], t[4]), x11 = fmaf(x.r2, t[7] - t[6], t[6]);
  const float y0 = fmaf(x.r1, x01 - x00, x00), y1 = fmaf(x.r1, x11 - x10, x10);
  return fmaf(x.r0, y1 - y0, y0);
}

// Two voxels per iteration, all sixteen shared-memory taps issued before any arithmetic that
// depends on them (the loads are volatile asm so ptxas keeps them batched).
__device__ __forceinline__ void k1_tile_staged_trilinear(const K1Ctx&, const K1Tile&, const K1Fast& f, const float* __restrict__ box) {
  const int dk = threadIdx.x & 15, jj = (threadIdx.x >> 4) & 1, di = threadIdx.x >> 5;
  if (dk >= f.n2 || di >= f.n0) return;
  const K1Hot h = k1_hot_load(f, di, dk);
  const uint32_t box_addr = smem_u32(box);
  const uint32_t o1 = 4u * f.p1, o0 = 4u * f.p0;
#pragma unroll 1
  for (int s = 0; s < 8; s += 2) {
    const int dja = 2 * s + jj, djb = dja + 2;
    if (dja >= h.n1) break;
    const bool hasb = djb < h.n1;
    const K1Vox xa = k1_fast_vox(h, dja, box_addr);
    const K1Vox xb = k1_fast_vox(h, hasb ? djb : dja, box_addr);
    float ta[8], tb[8];
    ta[0] = lds_f32(xa.a); ta[1] = lds_f32(xa.a + 4); ta[2] = lds_f32(xa.a + o1); ta[3] = lds_f32(xa.a + o1 + 4);
    tb[0] = lds_f32(xb.a); tb[1] = lds_f32(xb.a + 4); tb[2] = lds_f32(xb.a + o1); tb[3] = lds_f32(xb.a + o1 + 4);
    ta[4] = lds_f32(xa.a + o0); ta[5] = lds_f32(xa.a + o0 + 4); ta[6] = lds_f32(xa.a + o0 + o1); ta[7] = lds_f32(xa.a + o0 + o1 + 4);
    tb[4] = lds_f32(xb.a + o0); tb[5] = lds_f32(xb.a + o0 + 4); tb[6] = lds_f32(xb.a + o0 + o1); tb[7] = lds_f32(xb.a + o0 + o1 + 4);
    const float va = fmaf(k1_lerp8(xa, ta), h.gain, h.bias);
    const float vb = fmaf(k1_lerp8(xb, tb), h.gain, h.bias);
    k1_fast_store(f, h, di, dja, dk, va);
    if (hasb) k1_fast_store(f, h, di, djb, dk, vb);
  }
}

// 128-bit vectorised identity copy (flip / crop only, everything valid and aligned).  Each
// consumer thread moves two float4: rows (di, dj) and (di + 8, dj), quad k4 of the 16-wide tile.
__device__ __forceinline__ void k1_tile_copy_vec(const K1Ctx& c, const K1Tile& tl) {
  const adell_item& it = c.it;
  const int k4 = threadIdx.x & 3, dj = (threadIdx.x >> 2) & 15, di = threadIdx.x >> 6;
  const int o0 = tl.o0[0] + di, o1 = tl.o0[1] + dj, o2 = tl.o0[2] + 4 * k4;
  const int O0 = it.out_shape[0];
  if (o1 >= it.out_shape[1] || o2 >= it.out_shape[2] || o0 >= O0) return;
  const bool rev = it.grid_sign[2] * it.src_stride[2] < 0;
  const bool clip = (it.flags & ADELL_F_CLIP) != 0;
  const float pre_s = c.pre_s, pre_o = c.pre_o, post_s = it.post_scale, post_o = it.post_offset;
  const float clo = it.clip_lo, chi = it.clip_hi;
  const float gain = pre_s * post_s, bias = fmaf(pre_o, post_s, post_o);
  const bool plain = !clip && gain == 1.0f && bias == 0.0f;
  const int g0 = it.grid_off[0] + it.grid_sign[0] * o0;
  const int g1 = it.grid_off[1] + it.grid_sign[1] * o1;
  const int g2 = it.grid_off[2] + it.grid_sign[2] * (rev ? o2 + 3 : o2);  // lowest address of the quad
  const int64_t sstep = 8 * it.grid_sign[0] * it.src_stride[0], dstep = 8 * it.dst_stride[0];
  const float4* sp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(it.src) + g0 * it.src_stride[0] +
                                                      g1 * it.src_stride[1] + g2 * it.src_stride[2]);
  float4* dp = reinterpret_cast<float4*>(it.dst + o0 * it.dst_stride[0] + o1 * it.dst_stride[1] + o2);
  const bool two = o0 + 8 < O0 && di + 8 < K1_T;
  float4 qa = __ldg(sp);
  float4 qb = qa;
  if (two) qb = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(sp) + sstep));
  auto fix = [&](float4 q) {
    if (rev) { float t = q.x; q.x = q.w; q.w = t; t = q.y; q.y = q.z; q.z = t; }
    if (plain) return q;
    if (clip) {
      q.x = fmaf(k1_premap(q.x, pre_s, pre_o, true, clo, chi), post_s, post_o);
      q.y = fmaf(k1_premap(q.y, pre_s, pre_o, true, clo, chi), post_s, post_o);
      q.z = fmaf(k1_premap(q.z, pre_s, pre_o, true, clo, chi), post_s, post_o);
      q.w = fmaf(k1_premap(q.w, pre_s, pre_o, true, clo, chi), post_s, post_o);
    } else {
      q.x = fmaf(q.x, gain, bias); q.y = fmaf(q.y, gain, bias); q.z = fmaf(q.z, gain, bias); q.w = fmaf(q.w, gain, bias);
    }
    return q;
  };
  *dp = fix(qa);
  if (two) *reinterpret_cast<float4*>(reinterpret_cast<float*>(dp) + dstep) = fix(qb);
}

// ------------------------------------------------------------------------- per-tile set-up
__device__ void k1_tile_setup(const K1Ctx& c, K1Tile& tl, int b0, int b1, int b2) {
  const adell_item& it = c.it;
  tl.o0[0] = b0 * K1_T; tl.o0[1] = b1 * K1_T; tl.o0[2] = b2 * K1_T;
  tl.mode = MODE_DIRECT;
  tl.rmask = 0;
  tl.all_valid = 0;
  if (it.flags & ADELL_F_IDENTITY) {
    if (c.copy_ok) tl.mode = MODE_COPY;
    return;
  }
  if (!(it.flags & ADELL_F_TMAP)) return;
  // affine map of the tile in double: u_a(d) = U0_a + sum_b D_ab d_b   (t-space, before padding)
  double U0[3], D[3][3];
  double cc[3];
  for (int b = 0; b < 3; ++b) cc[b] = static_cast<double>(it.grid_off[b] + it.grid_sign[b] * tl.o0[b]) - static_cast<double>(c.cg[b]);
  bool finite = true;
  for (int a = 0; a < 3; ++a) {
    const double K = static_cast<double>(it.nrm[a]) * it.src_shape[a] * 0.5;
    const float* A = it.A + 4 * a;
    U0[a] = (A[0] * cc[0] + A[1] * cc[1] + A[2] * cc[2] + A[3]) * K + (it.src_shape[a] - 1) * 0.5;
    for (int b = 0; b < 3; ++b) D[a][b] = static_cast<double>(A[b]) * K * it.grid_sign[b];
    double umin = U0[a], umax = U0[a];
    for (int b = 0; b < 3; ++b) {
      const double span = D[a][b] * (min(K1_T, it.out_shape[b] - tl.o0[b]) - 1);
      if (span < 0) umin += span; else umax += span;
    }
    finite = finite && (umin > -1.0e6) && (umax < 1.0e6);
    if (!finite) break;
    tl.lo_t[a] = static_cast<int>(floor(umin - K1_EPS));
    tl.hi_t[a] = static_cast<int>(floor(umax + K1_EPS)) + 1;
  }
  if (!finite) return;
  bool all_valid = true, any_valid = true;
  int blo[3], bhi[3];  // source index interval the box must hold
  tl.rmask = 0;
  for (int a = 0; a < 3; ++a) {
    const int S = it.src_shape[a], lo = tl.lo_t[a], hi = tl.hi_t[a];
    blo[a] = lo; bhi[a] = hi;
    if (it.padding == ADELL_PAD_ZEROS) {
      if (hi < c.tlo[a] || lo >= c.thi[a]) any_valid = false;
    } else if (lo < 0 || hi > S - 1) {
      // coordinates leave [0,S): the voxel loop clamps / reflects them (ATen semantics), so the box
      // only has to hold the covering interval of the clamped / reflected indices (+1 for the hi tap)
      tl.rmask |= 1 << a;
      int rlo, rhi;
      if (it.padding == ADELL_PAD_BORDER) {
        rlo = min(max(lo, 0), S - 1); rhi = min(max(hi, 0), S - 1);
      } else if (lo >= -S && hi < 0) {            // inside the first mirrored period below
        rlo = -1 - hi; rhi = -1 - lo;
      } else if (lo >= S && hi <= 2 * S - 1) {    // inside the first mirrored period above
        rlo = 2 * S - 1 - hi; rhi = 2 * S - 1 - lo;
      } else if (lo < 0 && lo >= -S && hi <= S - 1) {   // straddles the lower edge
        rlo = 0; rhi = max(-1 - lo, hi);
      } else if (lo >= 0 && hi >= S && hi <= 2 * S - 1) {  // straddles the upper edge
        rlo = min(2 * S - 1 - hi, lo); rhi = S - 1;
      } else {
        rlo = 0; rhi = S - 1;
      }
      blo[a] = rlo; bhi[a] = rhi + 1;  // the hi tap of u' == S-1 reads cell S: TMA zero fill, weight 0
    }
    if (bhi[a] - blo[a] + 1 + (a == 2 ? 3 : 0) > it.tmap_box[a]) return;  // larger than the encoded box: direct path
    all_valid = all_valid && lo >= c.tlo[a] && hi < c.thi[a];
  }
  if (!any_valid) { tl.mode = MODE_ZERO; return; }
  // a pre offset must not leak into zero-filled (invalid) taps: such tiles use the exact path
  tl.all_valid = (all_valid && tl.rmask == 0) ? 1 : 0;
  if (tl.rmask != 0) {
    tl.all_valid = 1;
    for (int a = 0; a < 3; ++a) tl.all_valid = tl.all_valid && c.tlo[a] <= 0 && c.thi[a] >= it.src_shape[a];
  }
  for (int a = 0; a < 3; ++a) {
    tl.box[a] = it.tmap_box[a];
    tl.msign[a] = it.tmap_sign[a];
    int mo = tl.msign[a] > 0 ? blo[a] + it.tmap_off[a] : -bhi[a] + it.tmap_off[a];
    // TMA needs a 16-byte aligned start along the contiguous axis: round the box origin down to a
    // multiple of 4 elements (the encoded inner extent carries 3 spare elements for this)
    if (a == 2) mo = (mo >> 2) << 2;
    tl.mconst[a] = it.tmap_off[a] - mo;
    if (tl.rmask & (1 << a)) {
      tl.V0[a] = static_cast<float>(U0[a]);
      for (int b = 0; b < 3; ++b) tl.Dm[a][b] = static_cast<float>(D[a][b]);
      tl.rA[a] = static_cast<float>(tl.msign[a]);
      tl.rB[a] = static_cast<float>(tl.mconst[a]);
    } else {
      tl.V0[a] = static_cast<float>(tl.msign[a] * U0[a] + tl.mconst[a]);
      for (int b = 0; b < 3; ++b) tl.Dm[a][b] = static_cast<float>(tl.msign[a] * D[a][b]);
      tl.rA[a] = 1.0f; tl.rB[a] = 0.0f;
    }
  }
  tl.mode = MODE_STAGED;
}

#ifdef K1_PROFILE
__device__ unsigned long long k1_prof[8];
#define K1_PROF_T0 long long _t0 = clock64();
#define K1_PROF_ADD(i) { long long _t1 = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&k1_prof[i], (unsigned long long)(_t1 - _t0)); _t0 = _t1; }
#else
#define K1_PROF_T0
#define K1_PROF_ADD(i)
#endif

struct K1Slot {
  K1Ctx ctx;
  K1Tile tl;
  K1Fast fast;
};

// Producer: derive the tile state (and the consumers' register image) in the given slot.  The
// item itself is fetched from global memory only when the tile sequence moves on to a new item
// (tiles of one item are consecutive); otherwise it is copied from the producer's private copy.
__device__ __forceinline__ void k1_prepare(const adell_item* __restrict__ items, const int32_t* __restrict__ tile_start,
                                           int tile, int& item, int& cur_start, int& next_start, int& cached_item,
                                           K1Ctx& priv, K1Slot& sl, int lane) {
  while (tile >= next_start) { ++item; cur_start = next_start; next_start = __ldg(tile_start + item + 1); }
  uint32_t* pw = reinterpret_cast<uint32_t*>(&priv) + 32;
  if (item != cached_item) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(items + item) + 32;  // skip the tensor map
    uint32_t w0 = __ldg(src + lane), w1 = __ldg(src + lane + 32), w2 = __ldg(src + lane + 64);
    pw[lane] = w0; pw[lane + 32] = w1; pw[lane + 64] = w2;
    __syncwarp();
    if (lane == 0) k1_ctx_finish(priv);
    __syncwarp();
    cached_item = item;
  }
  {
    // whole K1Ctx (item image + derived constants), minus the unused tensor-map bytes
    constexpr int kWords = (sizeof(K1Ctx) - 128) / 4;
    uint32_t* dw = reinterpret_cast<uint32_t*>(&sl.ctx) + 32;
    for (int w = lane; w < kWords; w += 32) dw[w] = pw[w];
  }
  __syncwarp();
  if (lane == 0) {
    int n0, n1, n2;
    k1_tile_counts(sl.ctx.it.out_shape, n0, n1, n2);
    int local = tile - cur_start;
    const int b2 = local % n2; local /= n2;
    const int b1 = local % n1;
    const int b0 = local / n1;
    k1_tile_setup(sl.ctx, sl.tl, b0, b1, b2);
    sl.tl.item = item;
    if (sl.tl.mode == MODE_STAGED) k1_fast_fill(sl.ctx, sl.tl, sl.fast);
  }
  __syncwarp();
}

// Persistent, warp-specialised: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...  The
// producer warp prepares tile k+1 (item fetch + set-up) while the TMA box load of tile k is in
// flight, and issues each load the moment its ring stage is released; the 16 consumer warps
// interpolate the oldest full stage.  Rings: n_stages boxes (full/empty mbarrier pair each) and
// n_stages+1 tile-state slots, so set-up never waits for shared-memory space.
__global__ void __launch_bounds__(K1_THREADS, 1)
k1_gather(const adell_item* __restrict__ items, const int32_t* __restrict__ tile_start, int n_items, int total_tiles,
          int n_stages, int stage_bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int n_slots = n_stages + K1_PWARPS;
  K1Slot* slots = reinterpret_cast<K1Slot*>(smem + static_cast<size_t>(n_stages) * stage_bytes);
  K1Ctx* privs = reinterpret_cast<K1Ctx*>(slots + n_slots);   // one private item copy per producer warp
  uint64_t* full = reinterpret_cast<uint64_t*>(privs + K1_PWARPS);
  uint64_t* empty = full + n_stages;
  volatile int* issued = reinterpret_cast<volatile int*>(empty + n_stages);  // tiles issued so far (in order)
  if (threadIdx.x == 0) {
    *issued = 0;
    for (int s = 0; s < n_stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + s)), "r"(K1_CWARPS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int stage = 0, phase = 0, slot = 0;

  if (threadIdx.x >= K1_CTHREADS) {
    // ------------------------------------------------------------------ producer warps
    // warp p owns tiles k = p, p+P, ... of this CTA's sequence; ring positions follow k
    const int pw = (threadIdx.x - K1_CTHREADS) >> 5;
    K1Ctx& priv = privs[pw];
    int item = 0, cached_item = -1, cur_start = 0;
    int next_start = __ldg(tile_start + 1);
    int tile = blockIdx.x + pw * gridDim.x;
    int seq = pw;  // position of `tile` in this CTA's tile sequence
    for (int j = 0; j < pw; ++j) {
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
      if (++slot == n_slots) slot = 0;
    }
    while (tile < total_tiles) {
      K1_PROF_T0
      // safe to overwrite: the slot's previous tile (k - n_slots) was released before this warp's
      // previous issue (tile k - P waited for the stage of tile k - P - n_stages = k - n_slots)
      k1_prepare(items, tile_start, tile, item, cur_start, next_start, cached_item, priv, slots[slot], lane);
      K1_PROF_ADD(2)
      // issues happen strictly in tile order (the parity wait below is only meaningful for the
      // warp that is at most one ring revolution behind the consumers)
      if (lane == 0) { while (*issued != seq) { __nanosleep(20); } }
      __syncwarp();
      mbar_wait(empty + stage, phase ^ 1);
      K1_PROF_ADD(0)
      if (lane == 0) {
        const K1Slot& sl = slots[slot];
        if (sl.tl.mode == MODE_STAGED) {
          const int mo0 = sl.ctx.it.tmap_off[0] - sl.tl.mconst[0], mo1 = sl.ctx.it.tmap_off[1] - sl.tl.mconst[1],
                    mo2 = sl.ctx.it.tmap_off[2] - sl.tl.mconst[2];
          tmap_acquire(items[sl.tl.item].tmap);
          mbar_expect_tx(full + stage, static_cast<uint32_t>(sl.tl.box[0] * sl.tl.box[1] * sl.tl.box[2] * 4));
          tma_load_3d(smem + static_cast<size_t>(stage) * stage_bytes, items[sl.tl.item].tmap, full + stage, mo2, mo1, mo0);
        } else {
          mbar_arrive(full + stage);
        }
        __threadfence_block();
        *issued = seq + 1;
      }
      __syncwarp();
      K1_PROF_ADD(1)
      seq += K1_PWARPS;
      for (int j = 0; j < K1_PWARPS; ++j) {
        if (++stage == n_stages) { stage = 0; phase ^= 1; }
        if (++slot == n_slots) slot = 0;
      }
      tile += K1_PWARPS * gridDim.x;
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    K1_PROF_T0
    mbar_wait(full + stage, phase);
    K1_PROF_ADD(3)
    const K1Ctx& ctx = slots[slot].ctx;
    const K1Tile& tl = slots[slot].tl;
    const float* box = reinterpret_cast<const float*>(smem + static_cast<size_t>(stage) * stage_bytes);
    const adell_item& it = ctx.it;
    const int mode = tl.mode;
    if (mode == MODE_COPY) {
      k1_tile_copy_vec(ctx, tl);
    } else if (mode == MODE_ZERO) {
      const bool strict = (it.flags & ADELL_F_STRICT) != 0;
      k1_for_each_voxel(tl, it, [&](int, int, int, int o0, int o1, int o2) { k1_finish(it, 0.0f, o0, o1, o2, strict); });
    } else if (mode == MODE_STAGED) {
      const bool exact = (it.flags & (ADELL_F_STRICT | ADELL_F_CLIP)) != 0 || (ctx.pre_o != 0.0f && !tl.all_valid);
      if (exact) k1_tile_exact_dispatch<SmemTaps, false>(ctx, tl, box);
      else if (it.interp == ADELL_NEAREST) k1_tile_staged_nearest(ctx, tl, slots[slot].fast, box);
      else k1_tile_staged_trilinear(ctx, tl, slots[slot].fast, box);
    } else if (it.flags & ADELL_F_IDENTITY) {
      k1_tile_exact_dispatch<GlobalTaps, true>(ctx, tl, nullptr);
    } else {
      k1_tile_exact_dispatch<GlobalTaps, false>(ctx, tl, nullptr);
    }
    __syncwarp();
    K1_PROF_ADD(4)
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == n_stages) { stage = 0; phase ^= 1; }
    if (++slot == n_slots) slot = 0;
  }
}

int k1_validate(const adell_item& it) {
  for (int a = 0; a < 3; ++a) {
    if (it.out_shape[a] <= 0 || it.src_shape[a] <= 0 || it.grid_shape[a] <= 0) return ADELL_ERR_BAD_ARG;
    if (it.grid_sign[a] != 1 && it.grid_sign[a] != -1) return ADELL_ERR_BAD_ARG;
  }
  if (it.src_dtype > ADELL_U8) return ADELL_ERR_DTYPE;
  if (it.interp > ADELL_TRILINEAR || it.padding > ADELL_PAD_REFLECTION) return ADELL_ERR_BAD_ARG;
  if (it.src == nullptr || it.dst == nullptr) return ADELL_ERR_BAD_ARG;
  return ADELL_OK;
}

// ---- host: tensor-map encoding -------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn k1_get_encode() {
  // the driver entry point never changes within a process: look it up once (benign cache)
  static EncodeTiledFn cached = nullptr;
  if (cached != nullptr) return cached;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) { (void)cudaGetLastError(); return nullptr; }
  cached = reinterpret_cast<EncodeTiledFn>(fn);
  return cached;
}

// Decides staged-path eligibility for one item and, when eligible, encodes its tensor map over
// the valid source box in memory order.  Returns the box bytes (0 = not staged).
int k1_encode_item(adell_item& it, EncodeTiledFn enc) {
  it.flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
  if (it.flags & ADELL_F_IDENTITY) return 0;
  if (it.src_dtype != ADELL_F32) return 0;
  if (it.src_stride[2] != 1 && it.src_stride[2] != -1) return 0;
  int box[3];
  int64_t cells = 1;
  for (int a = 0; a < 3; ++a) {
    const double K = static_cast<double>(it.nrm[a]) * it.src_shape[a] * 0.5;
    double span = 0.0;
    for (int b = 0; b < 3; ++b) {
      const int tb = it.out_shape[b] < K1_T ? it.out_shape[b] : K1_T;
      span += fabs(static_cast<double>(it.A[4 * a + b]) * K) * (tb - 1);
    }
    if (!(span < 4096.0)) return 0;
    box[a] = static_cast<int>(floor(span + 2 * K1_EPS)) + 3;
    if (it.padding != ADELL_PAD_ZEROS) {
      // border / reflection fold the footprint back into [0,S): one spare cell for the hi tap, and a
      // thin axis is simply staged whole so that multiply-reflected tiles stay on this path
      box[a] += 1;
      if (it.src_shape[a] <= 64 && box[a] < it.src_shape[a] + 1) box[a] = it.src_shape[a] + 1;
    }
    if (a == 2) box[a] = (box[a] + 3 + 3) & ~3;  // inner extent: 16-byte multiple + slack to align the origin
    if (box[a] > 256) return 0;
    cells *= box[a];
  }
  if (cells * 4 > K1_MAX_BOX_BYTES) return 0;
  // valid source box in t-space and its origin in memory order
  cuuint64_t gdim[3], gstride[2];
  int64_t base_off = 0;
  int64_t astride[3];
  for (int a = 0; a < 3; ++a) {
    const int tlo = it.src_vlo[a] > 0 ? it.src_vlo[a] : 0;
    const int thi = it.src_vhi[a] < it.src_shape[a] ? it.src_vhi[a] : it.src_shape[a];
    if (thi <= tlo) return 0;
    const int sign = it.src_stride[a] >= 0 ? 1 : -1;
    astride[a] = it.src_stride[a] * sign;
    if (astride[a] == 0) return 0;
    it.tmap_sign[a] = sign;
    it.tmap_off[a] = sign > 0 ? -tlo : thi - 1;           // m = sign*t + off, m = 0 at the lowest address
    base_off += static_cast<int64_t>(sign > 0 ? tlo : thi - 1) * it.src_stride[a];
    gdim[2 - a] = static_cast<cuuint64_t>(thi - tlo);     // tensor-map dims are innermost first
    it.tmap_box[a] = box[a];
  }
  const uintptr_t base = reinterpret_cast<uintptr_t>(it.src) + static_cast<uintptr_t>(base_off * 4);
  gstride[0] = static_cast<cuuint64_t>(astride[1]) * 4;
  gstride[1] = static_cast<cuuint64_t>(astride[0]) * 4;
  if ((base & 15u) || (gstride[0] & 15u) || (gstride[1] & 15u)) return 0;
  if (gstride[0] < gdim[0] * 4 || gstride[1] < gstride[0]) return 0;  // rows must not overlap
  const cuuint32_t bdim[3] = {static_cast<cuuint32_t>(box[2]), static_cast<cuuint32_t>(box[1]), static_cast<cuuint32_t>(box[0])};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc == nullptr) return -1;
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(it.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   reinterpret_cast<void*>(base), gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 0;
  it.tmap_base = reinterpret_cast<const void*>(base);
  it.flags |= ADELL_F_TMAP;
  return static_cast<int>(cells * 4);
}

}  // namespace

extern "C" int adell_aug_prepare(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info) {
  if (items_host == nullptr || tile_start_host == nullptr || info == nullptr || n_items < 0) return ADELL_ERR_BAD_ARG;
  int64_t acc = 0;
  int smem = 0, staged = 0;
  EncodeTiledFn enc = nullptr;
  bool enc_tried = false;
  const char* dis = getenv("ADELL_DISABLE_STAGED");  // debugging aid: force the direct path
  const bool no_staged = dis != nullptr && dis[0] == '1';
  for (int i = 0; i < n_items; ++i) {
    int st = k1_validate(items_host[i]);
    if (st != ADELL_OK) return st;
    if (!enc_tried && !no_staged) { enc = k1_get_encode(); enc_tried = true; }
    int bytes = 0;
    if (no_staged) items_host[i].flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
    else bytes = k1_encode_item(items_host[i], enc);
    if (bytes < 0) return ADELL_ERR_NO_DRIVER;
    if (bytes > 0) { ++staged; if (bytes > smem) smem = bytes; }
    int n0, n1, n2;
    k1_tile_counts(items_host[i].out_shape, n0, n1, n2);
    tile_start_host[i] = static_cast<int32_t>(acc);
    acc += static_cast<int64_t>(n0) * n1 * n2;
    if (acc > 0x7fffffffLL) return ADELL_ERR_BAD_ARG;
  }
  tile_start_host[n_items] = static_cast<int32_t>(acc);
  info->total_tiles = acc;
  info->smem_bytes = smem;
  info->n_staged = staged;
  return ADELL_OK;
}

extern "C" int adell_aug_gather(const adell_item* items_dev, const int32_t* tile_start_dev, int n_items,
                                const adell_launch_info* info, void* stream) {
  if (n_items == 0) return ADELL_OK;
  if (info == nullptr) return ADELL_ERR_BAD_ARG;
  if (info->total_tiles == 0) return ADELL_OK;
  if (items_dev == nullptr || tile_start_dev == nullptr || n_items < 0 || info->total_tiles < 0 ||
      info->total_tiles > 0x7fffffffLL || info->smem_bytes < 0 || info->smem_bytes > K1_MAX_BOX_BYTES)
    return ADELL_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(items_dev) & 63u) != 0) return ADELL_ERR_ALIGN;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // ring of staged boxes: as many stages as fit next to the per-stage tile state
  const int stage_bytes = (info->smem_bytes + 127) & ~127;
  const int per_stage = stage_bytes + static_cast<int>(sizeof(K1Slot)) + 16;  // box + tile state + 2 mbarriers
  const int fixed = static_cast<int>(K1_PWARPS * (sizeof(K1Slot) + sizeof(K1Ctx))) + 16;
  int n_stages = (K1_SMEM_BUDGET - fixed) / per_stage;
  if (n_stages > K1_MAX_STAGES) n_stages = K1_MAX_STAGES;
  if (n_stages < 1) return ADELL_ERR_BAD_ARG;
  const int smem = n_stages * per_stage + fixed;
  e = cudaFuncSetAttribute(k1_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BUDGET);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  const int64_t grid = info->total_tiles < sms ? info->total_tiles : sms;
  k1_gather<<<static_cast<unsigned>(grid), K1_THREADS, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(
      items_dev, tile_start_dev, n_items, static_cast<int>(info->total_tiles), n_stages, stage_bytes);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_aug_gather_launches(void) { return 1; }

#ifdef K1_PROFILE
// debug builds only: cumulative cycle counters {producer wait-empty, issue, prepare, consumer wait-full, compute}
extern "C" int adell_debug_prof(unsigned long long* out8, int reset) {
  cudaMemcpyFromSymbol(out8, k1_prof, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(k1_prof, z, sizeof(z)); }
  return 0;
}
#endif
