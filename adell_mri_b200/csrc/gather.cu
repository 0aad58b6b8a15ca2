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
// This file holds the DIRECT path: taps are fetched with read-only global loads (L1/L2
// served).  It is the generic path (any padding mode, any footprint, int16/uint8 sources,
// nearest masks, identity copies).
#include "k1_math.cuh"

namespace {

constexpr int K1_THREADS = 256;
constexpr int K1_TI = 4;   // tile rows along axis 0
constexpr int K1_TJ = 8;   // tile rows along axis 1

__host__ __device__ inline int k1_kw(int o2) {
  // lanes along the contiguous axis: the largest of 32/16/8 that divides O2, else 32
  if (o2 % 32 == 0) return 32;
  if (o2 % 16 == 0) return 16;
  if (o2 % 8 == 0) return 8;
  return o2 >= 24 ? 32 : (o2 >= 12 ? 16 : 8);
}

__host__ __device__ inline void k1_tile_counts(const int32_t* O, int& n0, int& n1, int& n2, int& kw) {
  kw = k1_kw(O[2]);
  n0 = (O[0] + K1_TI - 1) / K1_TI;
  n1 = (O[1] + K1_TJ - 1) / K1_TJ;
  n2 = (O[2] + kw - 1) / kw;
}

template <int DT>
__device__ __forceinline__ float k1_tap(const K1Ctx& c, int t0, int t1, int t2) {
  int64_t idx = t0 * c.it.src_stride[0] + t1 * c.it.src_stride[1] + t2 * c.it.src_stride[2];
  if (DT == ADELL_F32) return adell_load_src_t<ADELL_F32>(c.it.src, idx);
  return adell_load_src(c.it.src, idx, c.it.src_dtype);
}

__device__ __forceinline__ bool k1_in(const K1Ctx& c, int t0, int t1, int t2) {
  return (t0 >= c.tlo[0]) & (t0 < c.thi[0]) & (t1 >= c.tlo[1]) & (t1 < c.thi[1]) & (t2 >= c.tlo[2]) &
         (t2 < c.thi[2]);
}

// One output voxel of a resampled item.  PERTAP: apply the pre map (and clip) to every tap and
// accumulate in ATen order with separate mul/add (ADELL_F_STRICT or ADELL_F_CLIP); otherwise
// accumulate sum(w*v) and sum(w_valid) with fma and apply the pre map once.
template <int INTERP, int PAD, int DT, bool PERTAP>
__device__ __forceinline__ float k1_resample_voxel(const K1Ctx& c, int g0, int g1, int g2) {
  const float c0 = static_cast<float>(g0) - c.cg[0];
  const float c1 = static_cast<float>(g1) - c.cg[1];
  const float c2 = static_cast<float>(g2) - c.cg[2];
  float u0 = k1_pad_coord<PAD>(k1_coord_exact(c, 0, c0, c1, c2), c.Sf[0], c.Sm1[0]);
  float u1 = k1_pad_coord<PAD>(k1_coord_exact(c, 1, c0, c1, c2), c.Sf[1], c.Sm1[1]);
  float u2 = k1_pad_coord<PAD>(k1_coord_exact(c, 2, c0, c1, c2), c.Sf[2], c.Sm1[2]);
  const bool clip = (c.it.flags & ADELL_F_CLIP) != 0;

  if (INTERP == ADELL_NEAREST) {
    int t0 = __float2int_rn(u0), t1 = __float2int_rn(u1), t2 = __float2int_rn(u2);
    if (!k1_in(c, t0, t1, t2)) return 0.0f;
    return k1_premap(k1_tap<DT>(c, t0, t1, t2), c.pre_s, c.pre_o, clip, c.it.clip_lo, c.it.clip_hi);
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
          float v = k1_tap<DT>(c, t0, t1, t2);
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
  return k1_premap(k1_tap<DT>(c, g0, g1, g2), c.pre_s, c.pre_o, clip, c.it.clip_lo, c.it.clip_hi);
}

template <int INTERP, int PAD, int DT, bool PERTAP, bool IDENT>
__device__ __forceinline__ void k1_tile(const K1Ctx& c, int b0, int b1, int b2, int kw) {
  const adell_item& it = c.it;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpw = 32 / kw;                 // rows per warp step
  const int r = lane / kw;                 // row within the step
  const int o2 = b2 * kw + (lane - r * kw);
  const int o0 = b0 * K1_TI + (warp >> 1);
  const int j_base = b1 * K1_TJ + (warp & 1) * 4;
  if (o0 >= it.out_shape[0] || o2 >= it.out_shape[2]) return;
  const int g0 = it.grid_off[0] + it.grid_sign[0] * o0;
  const int g2 = it.grid_off[2] + it.grid_sign[2] * o2;
  const bool gv02 = (o0 >= it.out_vlo[0]) & (o0 < it.out_vhi[0]) & (o2 >= it.out_vlo[2]) & (o2 < it.out_vhi[2]);
  const bool strict = (it.flags & ADELL_F_STRICT) != 0;
  for (int s = 0; s < 4; s += rpw) {
    const int o1 = j_base + s + r;
    if (o1 >= it.out_shape[1]) continue;
    const int g1 = it.grid_off[1] + it.grid_sign[1] * o1;
    float val = 0.0f;
    if (gv02 && o1 >= it.out_vlo[1] && o1 < it.out_vhi[1]) {
      val = IDENT ? k1_identity_voxel<DT>(c, g0, g1, g2)
                  : k1_resample_voxel<INTERP, PAD, DT, PERTAP>(c, g0, g1, g2);
    }
    // post intensity map (RandScaleIntensityd / RandShiftIntensityd) and noise
    if (strict) {
      if (it.post_scale != 1.0f) val = __fmul_rn(val, it.post_scale);
      if (it.post_offset != 0.0f) val = __fadd_rn(val, it.post_offset);
    } else {
      val = fmaf(val, it.post_scale, it.post_offset);
    }
    const int64_t olin = (static_cast<int64_t>(o0) * it.out_shape[1] + o1) * it.out_shape[2] + o2;
    if (it.noise != nullptr) val = __fadd_rn(val, __ldg(it.noise + olin));
    if (it.flags & ADELL_F_PHILOX)
      val = fmaf(it.noise_std, adell_philox_normal(it.philox_seed, it.philox_offset + olin), val);
    it.dst[o0 * it.dst_stride[0] + o1 * it.dst_stride[1] + o2 * it.dst_stride[2]] = val;
  }
}

template <int INTERP, int PAD>
__device__ __forceinline__ void k1_dispatch_dt(const K1Ctx& c, int b0, int b1, int b2, int kw) {
  const bool pertap = (c.it.flags & (ADELL_F_STRICT | ADELL_F_CLIP)) != 0;
  if (c.it.src_dtype == ADELL_F32) {
    if (pertap) k1_tile<INTERP, PAD, ADELL_F32, true, false>(c, b0, b1, b2, kw);
    else k1_tile<INTERP, PAD, ADELL_F32, false, false>(c, b0, b1, b2, kw);
  } else {
    if (pertap) k1_tile<INTERP, PAD, -1, true, false>(c, b0, b1, b2, kw);
    else k1_tile<INTERP, PAD, -1, false, false>(c, b0, b1, b2, kw);
  }
}

__global__ void __launch_bounds__(K1_THREADS)
k1_gather_direct(const adell_item* __restrict__ items, const int32_t* __restrict__ tile_start, int n_items) {
  __shared__ K1Ctx ctx;
  // block -> (item, tile): binary search in the exclusive prefix of tile counts
  const int tile = blockIdx.x;
  int lo = 0, hi = n_items;  // invariant: tile_start[lo] <= tile < tile_start[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(tile_start + mid) <= tile) lo = mid; else hi = mid;
  }
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(items + lo);
    uint32_t* d = reinterpret_cast<uint32_t*>(&ctx.it);
    if (threadIdx.x < sizeof(adell_item) / 4) d[threadIdx.x] = __ldg(s + threadIdx.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) k1_ctx_finish(ctx);
  __syncthreads();

  int n0, n1, n2, kw;
  k1_tile_counts(ctx.it.out_shape, n0, n1, n2, kw);
  int local = tile - __ldg(tile_start + lo);
  const int b2 = local % n2; local /= n2;
  const int b1 = local % n1;
  const int b0 = local / n1;

  const adell_item& it = ctx.it;
  if (it.flags & ADELL_F_IDENTITY) {
    if (it.src_dtype == ADELL_F32) k1_tile<0, 0, ADELL_F32, true, true>(ctx, b0, b1, b2, kw);
    else k1_tile<0, 0, -1, true, true>(ctx, b0, b1, b2, kw);
    return;
  }
  if (it.interp == ADELL_NEAREST) {
    if (it.padding == ADELL_PAD_ZEROS) k1_dispatch_dt<ADELL_NEAREST, ADELL_PAD_ZEROS>(ctx, b0, b1, b2, kw);
    else if (it.padding == ADELL_PAD_BORDER) k1_dispatch_dt<ADELL_NEAREST, ADELL_PAD_BORDER>(ctx, b0, b1, b2, kw);
    else k1_dispatch_dt<ADELL_NEAREST, ADELL_PAD_REFLECTION>(ctx, b0, b1, b2, kw);
  } else {
    if (it.padding == ADELL_PAD_ZEROS) k1_dispatch_dt<ADELL_TRILINEAR, ADELL_PAD_ZEROS>(ctx, b0, b1, b2, kw);
    else if (it.padding == ADELL_PAD_BORDER) k1_dispatch_dt<ADELL_TRILINEAR, ADELL_PAD_BORDER>(ctx, b0, b1, b2, kw);
    else k1_dispatch_dt<ADELL_TRILINEAR, ADELL_PAD_REFLECTION>(ctx, b0, b1, b2, kw);
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

}  // namespace

extern "C" int adell_aug_plan_tiles(const adell_item* items_host, int n_items, int32_t* tile_start_host,
                                    int64_t* total_tiles) {
  if (items_host == nullptr || tile_start_host == nullptr || n_items < 0) return ADELL_ERR_BAD_ARG;
  int64_t acc = 0;
  for (int i = 0; i < n_items; ++i) {
    int st = k1_validate(items_host[i]);
    if (st != ADELL_OK) return st;
    int n0, n1, n2, kw;
    k1_tile_counts(items_host[i].out_shape, n0, n1, n2, kw);
    tile_start_host[i] = static_cast<int32_t>(acc);
    acc += static_cast<int64_t>(n0) * n1 * n2;
    if (acc > 0x7fffffffLL) return ADELL_ERR_BAD_ARG;
  }
  tile_start_host[n_items] = static_cast<int32_t>(acc);
  if (total_tiles) *total_tiles = acc;
  return ADELL_OK;
}

extern "C" int adell_aug_gather(const adell_item* items_dev, const int32_t* tile_start_dev, int n_items,
                                int64_t total_tiles, void* stream) {
  if (n_items == 0 || total_tiles == 0) return ADELL_OK;
  if (items_dev == nullptr || tile_start_dev == nullptr || n_items < 0 || total_tiles < 0 ||
      total_tiles > 0x7fffffffLL)
    return ADELL_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(items_dev) & 63u) != 0) return ADELL_ERR_ALIGN;
  k1_gather_direct<<<static_cast<unsigned>(total_tiles), K1_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      items_dev, tile_start_dev, n_items);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_aug_gather_launches(void) { return 1; }
