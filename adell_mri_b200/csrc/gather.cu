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
// Persistent kernel, one CTA per SM, walking output tiles (extents chosen per item by
// adell_aug_prepare).  One producer warp prepares tile state and issues the TMA box loads into
// a shared-memory ring; sixteen consumer warps produce the voxels.  Paths, per tile:
//   STAGED  the tile's source footprint (a box whose extents depend only on the item's matrix)
//           is fetched by ONE TMA tensor copy (cp.async.bulk.tensor.3d, zero fill outside the
//           volume = "zeros" padding for free); taps are LDS.  Trilinear uses tile-local
//           incremental coordinates + nested lerps (<=1e-4 contract); floor / rint / float->int
//           are done with round-down adds of 1.5*2^23 on the FMA pipe (no XU conversions in the
//           loop).  Nearest re-evaluates the bit-faithful MONAI/ATen chain only when a
//           coordinate is within 1e-3 of a rounding tie => masks stay bit-exact.
//           ADELL_F_STRICT / ADELL_F_CLIP items take the bit-faithful chain for every voxel.
//   COPY    identity items (no resample fired): 128-bit vectorised flip/crop copy, 64 KiB tiles,
//           eight independent loads in flight per thread.
//   DIRECT  generic fallback: taps fetched with read-only global loads (int16/uint8 sources,
//           unaligned or oversized footprints, pad bands).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "k1_math.cuh"

namespace {

#ifndef K1_CWARPS_N
#define K1_CWARPS_N 16
#endif
constexpr int K1_CWARPS = K1_CWARPS_N;         // consumer warps per CTA
constexpr int K1_CTHREADS = 32 * K1_CWARPS;    // consumer threads
// The CTA runs K1_GROUPS independent streams: stream p = one producer warp, one hand-over warp and
// one group of consumer warps, with its own ring stages (p, p + K1_GROUPS, ...) and tile-state slots.
#define K1_GROUPS 2
#define K1_NPROD K1_GROUPS
constexpr int K1_THREADS = K1_CTHREADS + 64 * K1_GROUPS;   // consumers + per stream a hand-over and a producer warp
constexpr int K1_GWARPS = K1_CWARPS / K1_GROUPS;   // warps per group; a warp takes planes di = w, w + K1_GWARPS, ...
constexpr int K1_GTHREADS = 32 * K1_GWARPS;
constexpr int K1_T = 16;                       // base tile edge
constexpr int K1_COPY_T0 = 16;                 // identity items: 16x16x32 box copies (32 KiB: three ring stages per stream)
constexpr int K1_MAX_STAGES = 6;
constexpr int K1_SMEM_BUDGET = 227 * 1024;     // dynamic shared memory per persistent CTA (1 CTA per SM): the sm_100 opt-in maximum
constexpr int K1_MAX_BOX_BYTES = 100 * 1024;   // staged footprint limit (two stages)
constexpr int K1_TS_CACHE = 1024;               // tile-prefix entries cached in shared memory (else read from global)
constexpr double K1_EPS = 1e-3;                // coordinate slack of the fast path / tie window
constexpr float K1_MAGIC = 12582912.0f;        // 1.5 * 2^23: x + MAGIC has ulp 1 for |x| < 2^22
constexpr uint32_t K1_MAGIC_BITS = 0x4B400000u;
#ifndef K1_L2_PROMO
#define K1_L2_PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
#ifndef K1_STORE
#define K1_STORE __stcs
#endif
__device__ __forceinline__ void k1_plain_store(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void k1_wt_store(float4* p, float4 v) { __stwt(p, v); }
#ifndef K1_NP
#define K1_NP 1                                 // trilinear voxel PAIRS (packed fp32) per consumer-thread iteration (1 measured best: smaller loop body)
#endif
#ifndef K1_NV
#define K1_NV 4                                 // trilinear voxels interleaved per consumer-thread iteration
#endif

enum { MODE_DIRECT = 0, MODE_STAGED = 1, MODE_ZERO = 2, MODE_COPY = 3, MODE_DONE = 4, MODE_TSTORE = 5 };

struct K1Tile {
  int mode;
  int item;         // index of the tile's item (descriptor address for the TMA issue)
  int o0[3];        // tile origin (output index space)
  int T[3];         // tile extents (powers of two)
  int box[3];       // staged box extents, axes 0,1,2
  int mconst[3];    // box-local memory index = msign*t + mconst
  int msign[3];
  int lo_t[3], hi_t[3];
  float V0[3];      // fast path: coordinate of the tile-origin voxel — box-local memory order for
  float Dm[3][3];   // plain axes, absolute source index for axes in rmask; and its derivative
  int rmask;        // axes whose coordinates leave [0,S): border / reflection applied per voxel
  float rA[3], rB[3];  // box-local index = rA*u' + rB for the axes in rmask
  int all_valid;    // every tap of the tile lies inside the valid source region
  int fix_lo, fix_hi;  // box columns [fix_lo, fix_hi) along axis 2 hold bytes that precede the valid
                       // source box (16-byte alignment slack of the tensor map): zeroed after the load
  int next_plane;      // staged tiles: the next plane di to hand out (the group's warps draw planes from this
                       // counter instead of owning fixed ones: see k1_next_plane)
};

// Planes of a staged tile are handed out dynamically: a warp takes the next plane from a shared-memory counter
// when it has finished one.  With fixed planes (w, w + 8) the warps of a group finished a tile up to a plane's
// worth of time apart — partial tiles and sheared column groups leave some planes (nearly) empty — and since a
// stream has only two or three stages in flight, the fastest warp ran into the ring's limit and waited for the
// slowest: 16 % of the consumer cycles (k1_micro --prof, wait-full).
__device__ __forceinline__ int k1_next_plane(const K1Tile& tl) {
  int di = 0;
  if ((threadIdx.x & 31) == 0) di = atomicAdd(const_cast<int*>(&tl.next_plane), 1);
  return __shfl_sync(0xffffffffu, di, 0);
}

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
// Waits are done by the hardware, not by polling: mbarrier.try_wait takes a suspend-time hint and parks the thread
// until the phase completes (or the hint expires, then the loop retries).  Round 1 polled — try_wait without a hint
// returns after a short system-dependent time, and the nanosleep(128) between the polls of the producer / hand-over
// warps turned out to last ~14 ns — so that 11.5 % of all instructions the kernel executed were the polling loops of
// warps that had nothing to do (ncu source page, profiles/r02_k1_ncu_summary.md), competing for issue slots with the
// consumer warps of their schedulers.
#ifndef K1_WAIT_HINT_NS
#define K1_WAIT_HINT_NS 0x989680u   // 10 ms: effectively "until the barrier flips"
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(K1_WAIT_HINT_NS) : "memory");
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
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

// Shared -> global box store through the destination tensor map (bulk async group of the issuing thread).
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the committed stores of this thread have finished READING shared memory (the stage may be refilled)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  // the staged box holds the SOURCE element type (fp32 / int16 / uint8): converted at the tap
  template <int DT>
  static __device__ __forceinline__ float get(const K1Ctx& c, const K1Tile& t, const float* box, int t0, int t1, int t2) {
    int m0 = t.msign[0] * t0 + t.mconst[0], m1 = t.msign[1] * t1 + t.mconst[1], m2 = t.msign[2] * t2 + t.mconst[2];
    const int idx = (m0 * t.box[1] + m1) * t.box[2] + m2;
    const int dt = c.it.src_dtype;
    if (dt == ADELL_F32) return box[idx];
    if (dt == ADELL_I16) return static_cast<float>(reinterpret_cast<const short*>(box)[idx]);
    return static_cast<float>(reinterpret_cast<const unsigned char*>(box)[idx]);
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

// Thread -> voxel mapping of the generic paths: consecutive threads walk the tile in memory
// order (axis 2 fastest), so rows are written coalesced whatever the tile extents are.
template <class F>
__device__ __forceinline__ void k1_for_each_voxel(const K1Tile& tl, const adell_item& it, F&& body) {
  const int s2 = __ffs(tl.T[2]) - 1, s1 = __ffs(tl.T[1]) - 1;
  const int nvox = tl.T[0] << (s1 + s2);
  for (int v = threadIdx.x % K1_GTHREADS; v < nvox; v += K1_GTHREADS) {
    const int dk = v & (tl.T[2] - 1), dj = (v >> s2) & (tl.T[1] - 1), di = v >> (s1 + s2);
    const int o2 = tl.o0[2] + dk;
    const int G = (o2 >> 3) & 15;  // column group: its shift of the tile window (all zero without shear)
    const int o0 = tl.o0[0] + di - it.shear[0][G], o1 = tl.o0[1] + dj - it.shear[1][G];
    if (o0 < 0 || o1 < 0 || o0 >= it.out_shape[0] || o1 >= it.out_shape[1] || o2 >= it.out_shape[2]) continue;
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
// Tile-local incremental coordinates: v_a = V0_a + D0_a*di + D1_a*dj + D2_a*dk is the box-local
// (memory order) source coordinate; floor/frac/lerp directly on it.  The producer lays the
// consumers' register image out in 16-byte groups, so a consumer thread fetches it with a
// handful of broadcast LDS.128 per tile.
struct __align__(16) K1Fast {
  float4 ax[3];          // per source axis a: {V0_a, D0_a, D1_a, D2_a}
  float4 gb;             // gain, bias, post_offset, noise_std
  int4 n;                // T0, T1 (tile extents), n2 (voxels of the tile along axis 2), kw (16 or 32 lanes along axis 2)
  int4 lim;              // valid window of (di - s0) and (dj - s1): {-o0, O0 - o0, -o1, O1 - o1} (o = tile origin)
  float eg[4][4];        // per column group g = dk >> 3: {e_0, e_1, e_2} added to the fast coordinates,
                         // [3] = bits s0 | s1 << 16: the group's shift of the tile window (k1_item_shear)
  int4 m;                // p0, p1 (box pitches in elements), rmask | padding mode << 8, cbase: tap byte address =
                         // cbase + 4*(bits0*p0 + bits1*p1 + bits2), bits = float bits of coordinate + MAGIC
  int4 fl;               // cold (padded | noise | philox), padded, philox, -
  float4 rf[3];          // per source axis a: {1/(2 S_a), 2 S_a, S_a-1, rA_a} (axes in rmask)
  float4 rb;             // rB_0, rB_1, rB_2: box-local index = rA*u' + rB for the axes in rmask; [3] = tie threshold
  float4 ws;             // valid-weight variant (RM 3): {pre_o * post_s, post_o, -, -}; rf[a] = {lo_a - 1, hi_a + 1} then holds
                         // the box-local interval of valid taps along axis a
  int4 vlo, vhi;         // valid output range, tile-local
  float* dst;            // tile origin in the destination
  const float* noise;    // tile origin in the noise tensor (or null)
  int64_t ds0, ds1, ds2; // destination strides (elements)
  int64_t ns0, ns1;      // noise strides (contiguous [O0,O1,O2])
  uint64_t olin0;        // linear output index of the tile origin (Philox counter)
  uint64_t philox_seed, philox_offset;
};

// ATen compute_coordinates on the fast coordinate, branch-free.  Reflection about -0.5 and S-0.5 is a
// triangle wave of period 2S: with y = (u + 0.5) / 2S and f = y - rint(y) in [-0.5, 0.5], the
// reflected coordinate is |f| * 2S - 0.5 (then the clamp to [0, S-1] ATen applies as well).  The
// producer shifts the tile's coordinates by a whole number of periods so that |y| stays small
// (error of the fold: a few 1e-6 voxel).  rint is the round-to-nearest add of 1.5 * 2^23.
__device__ __forceinline__ float k1_fold_reflect(float u, float inv2S, float hinv, float twoS, float Sm1) {
  const float y = fmaf(u, inv2S, hinv);
  const float n = __fadd_rn(y, K1_MAGIC) - K1_MAGIC;
  const float x = fmaf(fabsf(y - n), twoS, -0.5f);
  return fminf(Sm1, fmaxf(x, 0.0f));
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
// Taps of integer sources (int16 / uint8 boxes staged as they are stored: half / a quarter of the bytes through
// TMA, L2 and shared memory).  The load sign- / zero-extends into a 32-bit register; the conversion to fp32 is
// the integer add of the bit pattern of 1.5 * 2^23 (exact for |x| < 2^22) followed by one FADD — no XU conversion
// instruction in the loop.  tap_bits<DT>: F32 -> the float itself, else the biased pattern (needs tap_fix).
template <int DT>
__device__ __forceinline__ float tap_raw(uint32_t addr) {
  if (DT == ADELL_F32) return lds_f32(addr);
  int v;
  if (DT == ADELL_I16) asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
  else asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return __int_as_float(v + static_cast<int>(K1_MAGIC_BITS));
}
template <int DT>
__device__ __forceinline__ float tap_f32(uint32_t addr) {
  const float r = tap_raw<DT>(addr);
  return DT == ADELL_F32 ? r : r - K1_MAGIC;
}
template <int DT> struct K1Es { static constexpr uint32_t v = DT == ADELL_F32 ? 4u : (DT == ADELL_I16 ? 2u : 1u); };
__device__ __forceinline__ uint32_t k1_es(int dt) { return dt == ADELL_F32 ? 4u : (dt == ADELL_I16 ? 2u : 1u); }
// runtime-dtype tap (cold paths)
__device__ __forceinline__ float tap_any(uint32_t addr, int dt) {
  if (dt == ADELL_F32) return tap_f32<ADELL_F32>(addr);
  if (dt == ADELL_I16) return tap_f32<ADELL_I16>(addr);
  return tap_f32<ADELL_U8>(addr);
}

// Per-thread register image of a staged tile.  RM: 0 = every coordinate of the tile stays inside
// [0,S); 1 = only source axis 2 leaves it and the padding is reflection (the common case of a thin,
// rotated volume); 2 = any axes, border or reflection (block-uniform branches per axis).
template <int RM>
struct K1Hot {
  float D1[3], P[3];      // coordinate along dj: v_a = P_a + D1_a * dj
  float gain, bias;
  uint32_t p0, p1, cbase; // tap address = cbase + 4*(bits0*p0 + bits1*p1 + bits2), bits = float bits of x + MAGIC
  float inv2S[3], hinv[3], twoS[3], Sm1[3], rA[3], rB[3];
  int rmask, pad;
  float vl[3], vh[3];     // RM 3: valid taps lie strictly between vl and vh (box-local)
  float wb, po;           // RM 3: pre_o * post_s (scaled by the valid weight), post_o
};

template <int RM>
__device__ __forceinline__ void k1_hot_plane(K1Hot<RM>& h, const K1Fast& f, int di, int dk, const float4 e) {
  const float fk = static_cast<float>(dk), fi = static_cast<float>(di);
  const float eg[3] = {e.x, e.y, e.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float4 q = f.ax[a];
    h.D1[a] = q.z;
    h.P[a] = fmaf(q.y, fi, fmaf(q.w, fk, q.x)) + eg[a];
  }
}
template <int RM>
__device__ __forceinline__ void k1_hot_load(K1Hot<RM>& h, const K1Fast& f) {
  const float4 gb = f.gb;
  const int4 m = f.m;
  h.gain = gb.x; h.bias = gb.y;
  h.p0 = static_cast<uint32_t>(m.x); h.p1 = static_cast<uint32_t>(m.y);
  h.cbase = static_cast<uint32_t>(m.w);  // computed by the producer (modulo 2^32 on purpose)
  if (RM == 3) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { const float4 q = f.rf[a]; h.vl[a] = q.x; h.vh[a] = q.y; }
    const float4 ws = f.ws;
    h.wb = ws.x; h.po = ws.y;
  } else if (RM) {
    h.rmask = m.z & 0xff; h.pad = m.z >> 8;
    const float4 rb = f.rb;
    h.rB[0] = rb.x; h.rB[1] = rb.y; h.rB[2] = rb.z;
#pragma unroll
    for (int a = (RM == 1 ? 2 : 0); a < 3; ++a) {
      const float4 q = f.rf[a];
      h.inv2S[a] = q.x; h.hinv[a] = 0.5f * q.x; h.twoS[a] = q.y; h.Sm1[a] = q.z; h.rA[a] = q.w;
    }
  }
}

// Thread -> voxels of a staged tile: warp = plane di, lanes along axis 2 (32, or 16 x 2 rows), loop
// over dj.  The thread's column group g = dk >> 3 shifts the tile window by (s0, s1) (k1_item_shear):
// output voxel (o0 + di - s0, o1 + dj - s1, o2 + dk).
struct K1Map {
  int dk, s0, s1;        // lane's column along axis 2 and its group's shifts
  int jj, jstep;         // lane's first row and row step (32 lanes along axis 2, or 16 x 2 rows)
  int jlo, jhi;          // window of dj that falls inside the output
  int ii;                // current plane: di - s0
  int j0, cnt;           // first dj of this thread and its number of voxels in the current plane
  float4 e;
};
// per tile: false when the lane has no voxels at all
__device__ __forceinline__ bool k1_lane_map(const K1Fast& f, K1Map& m) {
  const int4 n = f.n;
  const int lane = threadIdx.x & 31;
  if (n.w == 32) { m.dk = lane; m.jj = 0; m.jstep = 1; }
  else { m.dk = lane & 15; m.jj = lane >> 4; m.jstep = 2; }
  const bool in_row = m.dk < n.z;
  if (!in_row) m.dk = 0;   // (lanes beyond the row stay in the loops — the plane counter is drawn warp-wide — with no voxels)
  m.e = *reinterpret_cast<const float4*>(f.eg[m.dk >> 3]);
  const int sh = __float_as_int(m.e.w);
  m.s0 = sh & 0xffff;
  m.s1 = sh >> 16;
  const int4 lim = f.lim;
  m.jlo = max(0, lim.z + m.s1); m.jhi = min(n.y, lim.w + m.s1);
  m.j0 = m.jlo + ((m.jj - m.jlo) & (m.jstep - 1));
  m.cnt = (in_row && m.jhi > m.j0) ? (m.jhi - m.j0 + m.jstep - 1) / m.jstep : 0;
  return m.cnt > 0;
}
// per plane di: false when the plane's row of this lane lies outside the output
__device__ __forceinline__ bool k1_plane_map(const K1Fast& f, K1Map& m, int di) {
  m.ii = di - m.s0;
  return m.ii >= f.lim.x && m.ii < f.lim.y;
}

template <int RM>
__device__ __forceinline__ float k1_fast_pad_axis(const K1Hot<RM>& h, int a, float v) {
  float x;
  if (RM == 1 || h.pad == ADELL_PAD_REFLECTION) x = k1_fold_reflect(v, h.inv2S[a], h.hinv[a], h.twoS[a], h.Sm1[a]);
  else x = fminf(h.Sm1[a], fmaxf(v, 0.0f));
  return fmaf(h.rA[a], x, h.rB[a]);
}

template <int RM>
__device__ __forceinline__ void k1_fast_coords(const K1Hot<RM>& h, float fj, float& v0, float& v1, float& v2) {
  v0 = fmaf(h.D1[0], fj, h.P[0]); v1 = fmaf(h.D1[1], fj, h.P[1]); v2 = fmaf(h.D1[2], fj, h.P[2]);
  if (RM == 1) {
    v2 = k1_fast_pad_axis<RM>(h, 2, v2);
  } else if (RM == 2) {  // block-uniform: some axis leaves [0,S) inside this tile
    if (h.rmask & 1) v0 = k1_fast_pad_axis<RM>(h, 0, v0);
    if (h.rmask & 2) v1 = k1_fast_pad_axis<RM>(h, 1, v1);
    if (h.rmask & 4) v2 = k1_fast_pad_axis<RM>(h, 2, v2);
  }
}

struct K1Vox {
  float r0, r1, r2;
  uint32_t a;  // shared-memory byte address of tap (0,0,0)
  float w;     // RM 3: sum of the trilinear weights of the valid taps
};

// Sum of the two tap weights along one axis that fall on valid cells: 1 inside, a linear ramp over
// the first / last cell, 0 outside — a trapezoid of the coordinate (the valid region is a box, so
// the 3-D sum is the product of the three).
__device__ __forceinline__ float k1_valid_weight(float v, float vl, float vh) {
  return fminf(1.0f, fmaxf(fminf(v - vl, vh - v), 0.0f));
}

template <int RM>
__device__ __forceinline__ K1Vox k1_fast_vox(const K1Hot<RM>& h, float fj, const uint32_t es = 4u) {
  float v0, v1, v2;
  k1_fast_coords<RM>(h, fj, v0, v1, v2);
  // floor on the FMA pipe: round-down add of 1.5*2^23 leaves floor(x) in the low mantissa bits
  const float t0 = __fadd_rd(v0, K1_MAGIC), t1 = __fadd_rd(v1, K1_MAGIC), t2 = __fadd_rd(v2, K1_MAGIC);
  K1Vox x;
  x.r0 = v0 - (t0 - K1_MAGIC); x.r1 = v1 - (t1 - K1_MAGIC); x.r2 = v2 - (t2 - K1_MAGIC);
  x.a = h.cbase + es * (__float_as_uint(t0) * h.p0 + __float_as_uint(t1) * h.p1 + __float_as_uint(t2));
  if (RM == 3) x.w = k1_valid_weight(v0, h.vl[0], h.vh[0]) * k1_valid_weight(v1, h.vl[1], h.vh[1]) * k1_valid_weight(v2, h.vl[2], h.vh[2]);
  return x;
}

__device__ __forceinline__ float k1_lerp8(const K1Vox& x, const float* t) {
  const float x00 = fmaf(x.r2, t[1] - t[0], t[0]), x01 = fmaf(x.r2, t[3] - t[2], t[2]);
  const float x10 = fmaf(x.r2, t[5] - t[4], t[4]), x11 = fmaf(x.r2, t[7] - t[6], t[6]);
  const float y0 = fmaf(x.r1, x01 - x00, x00), y1 = fmaf(x.r1, x11 - x10, x10);
  return fmaf(x.r0, y1 - y0, y0);
}

// Trilinear, plain tiles (no output pad band, no noise): NV voxels per iteration, all their
// shared-memory taps issued before any arithmetic that depends on them (the loads are volatile
// asm so ptxas keeps them batched); no per-voxel bounds checks — a thread's voxel count is split
// into full groups and a one-at-a-time tail.
// NZ: the tile's only special feature is device (Philox) noise — one N(0,1) per output position added after the post map,
// the same fmaf as the cold loop's (the counter is the voxel's linear output index: independent of the tile shape).
template <int NV, int RMASK, int DT, bool NZ = false>
__device__ __forceinline__ void k1_tile_staged_trilinear_scalar(const K1Tile& tl, const K1Fast& f) {
  K1Map m;
  k1_lane_map(f, m);   // lanes without voxels: cnt == 0
  K1Hot<RMASK> h;
  k1_hot_load<RMASK>(h, f);
  constexpr uint32_t ES = K1Es<DT>::v;
  const uint32_t o1 = ES * h.p1, o0 = ES * h.p0;
  const int64_t ds1 = f.ds1;
  const int64_t pstep = m.jstep * ds1;
  const float fstep = static_cast<float>(m.jstep);
  const int T0 = f.n.x;
  const uint64_t nz_base = NZ ? f.philox_offset + f.olin0 : 0, nz_step = NZ ? static_cast<uint64_t>(m.jstep * f.ns1) : 0;
  const float nz_std = NZ ? f.gb.w : 0.0f;
#pragma unroll 1
  for (;;) {
    const int di = k1_next_plane(tl);
    if (di >= T0) break;
    if (!k1_plane_map(f, m, di)) continue;
    k1_hot_plane<RMASK>(h, f, di, m.dk, m.e);
    float* p = f.dst + m.ii * f.ds0 + m.dk * f.ds2 + (m.j0 - m.s1) * ds1;
    float fj = static_cast<float>(m.j0);
    int cnt = m.cnt;
    uint64_t nc = 0;
    if (NZ) nc = nz_base + static_cast<uint64_t>(m.ii * f.ns0 + (m.j0 - m.s1) * f.ns1 + m.dk);
#pragma unroll 1
    for (; cnt >= NV; cnt -= NV) {
      K1Vox x[NV];
      float t[NV][8];
#pragma unroll
      for (int u = 0; u < NV; ++u) x[u] = k1_fast_vox<RMASK>(h, fj + static_cast<float>(u) * fstep, ES);
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        t[u][0] = tap_f32<DT>(x[u].a); t[u][1] = tap_f32<DT>(x[u].a + ES);
        t[u][2] = tap_f32<DT>(x[u].a + o1); t[u][3] = tap_f32<DT>(x[u].a + o1 + ES);
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        t[u][4] = tap_f32<DT>(x[u].a + o0); t[u][5] = tap_f32<DT>(x[u].a + o0 + ES);
        t[u][6] = tap_f32<DT>(x[u].a + o0 + o1); t[u][7] = tap_f32<DT>(x[u].a + o0 + o1 + ES);
      }
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        float z = fmaf(k1_lerp8(x[u], t[u]), h.gain, h.bias);
        if (NZ) z = fmaf(nz_std, adell_philox_normal(f.philox_seed, nc + static_cast<uint64_t>(u) * nz_step), z);
        p[u * pstep] = z;
      }
      p += NV * pstep;
      fj += static_cast<float>(NV) * fstep;
      if (NZ) nc += static_cast<uint64_t>(NV) * nz_step;
    }
#pragma unroll 1
    for (; cnt > 0; --cnt) {
      const K1Vox x = k1_fast_vox<RMASK>(h, fj, ES);
      float t[8];
      t[0] = tap_f32<DT>(x.a); t[1] = tap_f32<DT>(x.a + ES); t[2] = tap_f32<DT>(x.a + o1); t[3] = tap_f32<DT>(x.a + o1 + ES);
      t[4] = tap_f32<DT>(x.a + o0); t[5] = tap_f32<DT>(x.a + o0 + ES); t[6] = tap_f32<DT>(x.a + o0 + o1); t[7] = tap_f32<DT>(x.a + o0 + o1 + ES);
      float z = fmaf(k1_lerp8(x, t), h.gain, h.bias);
      if (NZ) { z = fmaf(nz_std, adell_philox_normal(f.philox_seed, nc), z); nc += nz_step; }
      *p = z;
      p += pstep;
      fj += fstep;
    }
  }
}

// ---- packed fp32 (Blackwell FFMA2 / FADD2: two fp32 lanes per instruction in a 64-bit register pair).
// The trilinear loop is bound by instruction issue, not by the FMA pipe: doing the coordinate, floor /
// fraction and lerp arithmetic of two voxels per instruction removes about a quarter of its issue slots.
typedef unsigned long long k1_f2;
__device__ __forceinline__ k1_f2 f2_pack(float lo, float hi) { k1_f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ k1_f2 f2_dup(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(k1_f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ k1_f2 f2_fma(k1_f2 a, k1_f2 b, k1_f2 c) { k1_f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ k1_f2 f2_add(k1_f2 a, k1_f2 b) { k1_f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ k1_f2 f2_sub(k1_f2 a, k1_f2 b) { k1_f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ k1_f2 f2_add_rd(k1_f2 a, k1_f2 b) { k1_f2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ k1_f2 f2_abs(k1_f2 a) { return a & 0x7fffffff7fffffffull; }

// Two voxels (rows dj and dj + jstep of the thread) at a time.
struct K1Vox2 {
  k1_f2 r0, r1, r2;
  uint32_t aA, aB;  // shared-memory byte addresses of tap (0,0,0) of the two voxels
  k1_f2 w;          // RM 3: valid-weight sums of the two voxels
};
template <int RM>
struct K1Hot2 {
  k1_f2 D1[3], P[3];
  k1_f2 inv2S, hinv, twoS, rA, rB;   // RM == 1: reflection fold of source axis 2
  float Sm1;
  k1_f2 vl[3], vh[3];                // RM == 3: valid tap interval per axis
};
template <int RM>
__device__ __forceinline__ K1Vox2 k1_fast_vox2(const K1Hot<RM>& h, const K1Hot2<RM>& q, k1_f2 fj2, const uint32_t es = 4u) {
  const k1_f2 M = f2_dup(K1_MAGIC), NM = f2_dup(-K1_MAGIC);
  k1_f2 v0 = f2_fma(q.D1[0], fj2, q.P[0]), v1 = f2_fma(q.D1[1], fj2, q.P[1]), v2 = f2_fma(q.D1[2], fj2, q.P[2]);
  if (RM == 1) {  // k1_fold_reflect, two lanes at a time; the clamp has no packed form
    const k1_f2 y = f2_fma(v2, q.inv2S, q.hinv);
    const k1_f2 n = f2_add(f2_add(y, M), NM);
    const k1_f2 x = f2_fma(f2_abs(f2_sub(y, n)), q.twoS, f2_dup(-0.5f));
    float xa, xb;
    f2_unpack(x, xa, xb);
    xa = fminf(q.Sm1, fmaxf(xa, 0.0f)); xb = fminf(q.Sm1, fmaxf(xb, 0.0f));
    v2 = f2_fma(q.rA, f2_pack(xa, xb), q.rB);
  }
  // floor on the FMA pipe: round-down add of 1.5*2^23 leaves floor(x) in the low mantissa bits
  const k1_f2 t0 = f2_add_rd(v0, M), t1 = f2_add_rd(v1, M), t2 = f2_add_rd(v2, M);
  K1Vox2 x;
  x.r0 = f2_sub(v0, f2_add(t0, NM)); x.r1 = f2_sub(v1, f2_add(t1, NM)); x.r2 = f2_sub(v2, f2_add(t2, NM));
  float t0a, t0b, t1a, t1b, t2a, t2b;
  f2_unpack(t0, t0a, t0b); f2_unpack(t1, t1a, t1b); f2_unpack(t2, t2a, t2b);
  x.aA = h.cbase + es * (__float_as_uint(t0a) * h.p0 + __float_as_uint(t1a) * h.p1 + __float_as_uint(t2a));
  x.aB = h.cbase + es * (__float_as_uint(t0b) * h.p0 + __float_as_uint(t1b) * h.p1 + __float_as_uint(t2b));
  if (RM == 3) {  // valid-weight product, the subtractions packed, min / max per lane
    float wa = 1.0f, wb = 1.0f;
    const k1_f2 vv[3] = {v0, v1, v2};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float la, lb, ha, hb;
      f2_unpack(f2_sub(vv[a], q.vl[a]), la, lb);
      f2_unpack(f2_sub(q.vh[a], vv[a]), ha, hb);
      wa *= fminf(1.0f, fmaxf(fminf(la, ha), 0.0f));
      wb *= fminf(1.0f, fmaxf(fminf(lb, hb), 0.0f));
    }
    x.w = f2_pack(wa, wb);
  }
  return x;
}
__device__ __forceinline__ k1_f2 k1_lerp8x2(const K1Vox2& x, const k1_f2* t) {
  const k1_f2 x00 = f2_fma(x.r2, f2_sub(t[1], t[0]), t[0]), x01 = f2_fma(x.r2, f2_sub(t[3], t[2]), t[2]);
  const k1_f2 x10 = f2_fma(x.r2, f2_sub(t[5], t[4]), t[4]), x11 = f2_fma(x.r2, f2_sub(t[7], t[6]), t[6]);
  const k1_f2 y0 = f2_fma(x.r1, f2_sub(x01, x00), x00), y1 = f2_fma(x.r1, f2_sub(x11, x10), x10);
  return f2_fma(x.r0, f2_sub(y1, y0), y0);
}

// Trilinear, plain tiles, RM = 0 / 1: NP pairs of voxels per iteration with packed arithmetic, a
// one-voxel-at-a-time tail.  Same operations (and roundings) per voxel as the scalar loop.
// two taps (one of each voxel of the pair) -> packed fp32; integer sources: one packed FADD2 removes the bias of both
template <int DT>
__device__ __forceinline__ k1_f2 tap_pair(uint32_t aA, uint32_t aB) {
  const k1_f2 r = f2_pack(tap_raw<DT>(aA), tap_raw<DT>(aB));
  return DT == ADELL_F32 ? r : f2_add(r, f2_dup(-K1_MAGIC));
}

template <int NP, int RMASK, int DT, bool NZ = false>
__device__ __forceinline__ void k1_tile_staged_trilinear(const K1Tile& tl, const K1Fast& f) {
  K1Map m;
  k1_lane_map(f, m);   // lanes without voxels: cnt == 0
  K1Hot<RMASK> h;
  k1_hot_load<RMASK>(h, f);
  constexpr uint32_t ES = K1Es<DT>::v;
  const uint32_t o1 = ES * h.p1, o0 = ES * h.p0;
  const int64_t ds1 = f.ds1;
  const int64_t pstep = m.jstep * ds1;
  const float fstep = static_cast<float>(m.jstep);
  const int T0 = f.n.x;
  const k1_f2 G2 = f2_dup(h.gain), B2 = f2_dup(h.bias);
  const uint64_t nz_base = NZ ? f.philox_offset + f.olin0 : 0, nz_step = NZ ? static_cast<uint64_t>(m.jstep * f.ns1) : 0;
  const float nz_std = NZ ? f.gb.w : 0.0f;
  K1Hot2<RMASK> q;
  if (RMASK == 1) {
    q.inv2S = f2_dup(h.inv2S[2]); q.hinv = f2_dup(h.hinv[2]); q.twoS = f2_dup(h.twoS[2]);
    q.rA = f2_dup(h.rA[2]); q.rB = f2_dup(h.rB[2]); q.Sm1 = h.Sm1[2];
  }
  k1_f2 WB2 = 0, PO2 = 0;
  if (RMASK == 3) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { q.vl[a] = f2_dup(h.vl[a]); q.vh[a] = f2_dup(h.vh[a]); }
    WB2 = f2_dup(h.wb); PO2 = f2_dup(h.po);
  }
#pragma unroll 1
  for (;;) {
    const int di = k1_next_plane(tl);
    if (di >= T0) break;
    if (!k1_plane_map(f, m, di)) continue;
    k1_hot_plane<RMASK>(h, f, di, m.dk, m.e);
#pragma unroll
    for (int a = 0; a < 3; ++a) { q.D1[a] = f2_dup(h.D1[a]); q.P[a] = f2_dup(h.P[a]); }
    float* p = f.dst + m.ii * f.ds0 + m.dk * f.ds2 + (m.j0 - m.s1) * ds1;
    float fj = static_cast<float>(m.j0);
    int cnt = m.cnt;
    uint64_t nc = 0;
    if (NZ) nc = nz_base + static_cast<uint64_t>(m.ii * f.ns0 + (m.j0 - m.s1) * f.ns1 + m.dk);
#pragma unroll 1
    for (; cnt >= 2 * NP; cnt -= 2 * NP) {
      K1Vox2 x[NP];
      k1_f2 t[NP][8];
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        const float fa = fj + static_cast<float>(2 * u) * fstep;
        x[u] = k1_fast_vox2<RMASK>(h, q, f2_pack(fa, fa + fstep), ES);
      }
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        t[u][0] = tap_pair<DT>(x[u].aA, x[u].aB); t[u][1] = tap_pair<DT>(x[u].aA + ES, x[u].aB + ES);
        t[u][2] = tap_pair<DT>(x[u].aA + o1, x[u].aB + o1); t[u][3] = tap_pair<DT>(x[u].aA + o1 + ES, x[u].aB + o1 + ES);
      }
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        t[u][4] = tap_pair<DT>(x[u].aA + o0, x[u].aB + o0); t[u][5] = tap_pair<DT>(x[u].aA + o0 + ES, x[u].aB + o0 + ES);
        t[u][6] = tap_pair<DT>(x[u].aA + o0 + o1, x[u].aB + o0 + o1);
        t[u][7] = tap_pair<DT>(x[u].aA + o0 + o1 + ES, x[u].aB + o0 + o1 + ES);
      }
#pragma unroll
      for (int u = 0; u < NP; ++u) {
        float za, zb;
        // RM 3: gain * sum(w v) + (pre_o post_s) * sum(w valid) + post_o
        f2_unpack(f2_fma(k1_lerp8x2(x[u], t[u]), G2, RMASK == 3 ? f2_fma(x[u].w, WB2, PO2) : B2), za, zb);
        if (NZ) {
          za = fmaf(nz_std, adell_philox_normal(f.philox_seed, nc + static_cast<uint64_t>(2 * u) * nz_step), za);
          zb = fmaf(nz_std, adell_philox_normal(f.philox_seed, nc + static_cast<uint64_t>(2 * u + 1) * nz_step), zb);
        }
        p[(2 * u) * pstep] = za;
        p[(2 * u + 1) * pstep] = zb;
      }
      p += 2 * NP * pstep;
      fj += static_cast<float>(2 * NP) * fstep;
      if (NZ) nc += static_cast<uint64_t>(2 * NP) * nz_step;
    }
#pragma unroll 1
    for (; cnt > 0; --cnt) {
      const K1Vox x = k1_fast_vox<RMASK>(h, fj, ES);
      float t[8];
      t[0] = tap_f32<DT>(x.a); t[1] = tap_f32<DT>(x.a + ES); t[2] = tap_f32<DT>(x.a + o1); t[3] = tap_f32<DT>(x.a + o1 + ES);
      t[4] = tap_f32<DT>(x.a + o0); t[5] = tap_f32<DT>(x.a + o0 + ES); t[6] = tap_f32<DT>(x.a + o0 + o1); t[7] = tap_f32<DT>(x.a + o0 + o1 + ES);
      float z = fmaf(k1_lerp8(x, t), h.gain, RMASK == 3 ? fmaf(x.w, h.wb, h.po) : h.bias);
      if (NZ) { z = fmaf(nz_std, adell_philox_normal(f.philox_seed, nc), z); nc += nz_step; }
      *p = z;
      p += pstep;
      fj += fstep;
    }
  }
}

// Nearest, plain tiles: fast coordinates, exact replay inside the tie window.
template <int RMASK, int DT>
__device__ __forceinline__ void k1_tile_staged_nearest(const K1Ctx& c, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box) {
  constexpr uint32_t ES = K1Es<DT>::v;
  K1Map m;
  k1_lane_map(f, m);   // lanes without voxels: cnt == 0
  K1Hot<RMASK> h;
  k1_hot_load<RMASK>(h, f);
  const float tie = f.rb.w;
  const int64_t ds1 = f.ds1;
  const int64_t pstep = m.jstep * ds1;
  const int T0 = f.n.x;
  // Voxels inside the tie window of a rounding tie take the bit-faithful replay (k1_exact_nearest_smem: several hundred
  // instructions, and the whole warp waits for the one lane that needs it).  A few per mille of the voxels do — about
  // one warp iteration in ten — which made a mask tile cost 1.6-2x a trilinear one.  They are DEFERRED instead: a lane
  // remembers its tie voxel (tile-local plane and row, biased into 16 bits each) and the warp replays them together
  // once, after the tile's planes (a lane that meets a second one first replays the remembered voxel on the spot).
  int pend = -1;
  auto replay = [&](int pk) {
    const int ii = (pk >> 16) - 0x4000, jj = (pk & 0xffff) - 0x4000;
    f.dst[ii * f.ds0 + m.dk * f.ds2 + jj * ds1] = k1_exact_nearest_smem(c, tl, box, ii, jj, m.dk);
  };
#pragma unroll 1
  for (;;) {
    const int di = k1_next_plane(tl);
    if (di >= T0) break;
    if (!k1_plane_map(f, m, di)) continue;
    k1_hot_plane<RMASK>(h, f, di, m.dk, m.e);
    float* p = f.dst + m.ii * f.ds0 + m.dk * f.ds2 + (m.j0 - m.s1) * ds1;
    int dj = m.j0;
    int cnt = m.cnt;
    while (cnt > 0) {
      bool full = false;   // the lane met a second tie voxel: leave the voxel loop (which holds no call), replay, come back
#pragma unroll 2
      for (; cnt > 0; --cnt, dj += m.jstep, p += pstep) {
        float v0, v1, v2;
        k1_fast_coords<RMASK>(h, static_cast<float>(dj), v0, v1, v2);
        // rint (ties to even) on the FMA pipe: x + 1.5*2^23 rounds to the nearest integer
        const float t0 = __fadd_rn(v0, K1_MAGIC), t1 = __fadd_rn(v1, K1_MAGIC), t2 = __fadd_rn(v2, K1_MAGIC);
        const float n0 = t0 - K1_MAGIC, n1 = t1 - K1_MAGIC, n2 = t2 - K1_MAGIC;
        // the fast tap is loaded unconditionally, off the tie check's dependency chain (inside the tie window the fast
        // cell is one of the two tied cells: both lie in the staged box); only the store depends on the check
        const uint32_t a = h.cbase + ES * (__float_as_uint(t0) * h.p0 + __float_as_uint(t1) * h.p1 + __float_as_uint(t2));
        float val = fmaf(tap_f32<DT>(a), h.gain, h.bias);
        if (RMASK == 3) {  // an invalid (zero-filled) tap is a literal 0 of the padded volume: no pre offset
          const bool ok = (n0 > h.vl[0]) & (n0 < h.vh[0]) & (n1 > h.vl[1]) & (n1 < h.vh[1]) & (n2 > h.vl[2]) & (n2 < h.vh[2]);
          if (!ok) val = h.po;
        }
        if (fmaxf(fmaxf(fabsf(v0 - n0), fabsf(v1 - n1)), fabsf(v2 - n2)) > tie) {   // inside the tie window of a rounding tie
          if (pend >= 0) { full = true; break; }
          pend = ((m.ii + 0x4000) << 16) | (dj - m.s1 + 0x4000);
        } else {
          *p = val;
        }
      }
      if (full) { replay(pend); pend = -1; }   // (the voxel that did not fit is evaluated again and remembered)
    }
  }
  if (pend >= 0) replay(pend);
}

// Rare tiles: output pad band (SpatialPadd after the resample), injected or Philox noise.  Same
// arithmetic as the plain loops, one voxel at a time, out of line.
__device__ __noinline__ void k1_tile_staged_cold(const K1Ctx& c, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box) {
  K1Map m;
  k1_lane_map(f, m);   // lanes without voxels: cnt == 0
  const int dk = m.dk;
  K1Hot<2> h;
  k1_hot_load<2>(h, f);
  const bool nearest = c.it.interp == ADELL_NEAREST;
  const float tie = f.rb.w;
  const int dt = c.it.src_dtype;
  const uint32_t es = k1_es(dt);
  const uint32_t o1 = es * h.p1, o0 = es * h.p0;
  const int4 vlo = f.vlo, vhi = f.vhi, fl = f.fl;
  const float4 gb = f.gb;
  const int T0 = f.n.x;
  for (;;) {
    const int di = k1_next_plane(tl);
    if (di >= T0) break;
    if (!k1_plane_map(f, m, di)) continue;
    k1_hot_plane<2>(h, f, di, dk, m.e);
    int dj = m.j0;
    for (int cnt = m.cnt; cnt > 0; --cnt, dj += m.jstep) {
      const int jo = dj - m.s1;  // tile-local output position (ii, jo, dk)
      float val;
      if (nearest) {
        float v0, v1, v2;
        k1_fast_coords<2>(h, static_cast<float>(dj), v0, v1, v2);
        const float t0 = __fadd_rn(v0, K1_MAGIC), t1 = __fadd_rn(v1, K1_MAGIC), t2 = __fadd_rn(v2, K1_MAGIC);
        if (fabsf(v0 - (t0 - K1_MAGIC)) > tie || fabsf(v1 - (t1 - K1_MAGIC)) > tie || fabsf(v2 - (t2 - K1_MAGIC)) > tie)
          val = k1_exact_nearest_smem(c, tl, box, m.ii, jo, dk);
        else
          val = fmaf(tap_any(h.cbase + es * (__float_as_uint(t0) * h.p0 + __float_as_uint(t1) * h.p1 + __float_as_uint(t2)), dt), h.gain, h.bias);
      } else {
        const K1Vox x = k1_fast_vox<2>(h, static_cast<float>(dj), es);
        float t[8];
        t[0] = tap_any(x.a, dt); t[1] = tap_any(x.a + es, dt); t[2] = tap_any(x.a + o1, dt); t[3] = tap_any(x.a + o1 + es, dt);
        t[4] = tap_any(x.a + o0, dt); t[5] = tap_any(x.a + o0 + es, dt); t[6] = tap_any(x.a + o0 + o1, dt); t[7] = tap_any(x.a + o0 + o1 + es, dt);
        val = fmaf(k1_lerp8(x, t), h.gain, h.bias);
      }
      if (fl.y) {
        const bool ov = (m.ii >= vlo.x) & (m.ii < vhi.x) & (jo >= vlo.y) & (jo < vhi.y) & (dk >= vlo.z) & (dk < vhi.z);
        if (!ov) val = gb.z;
      }
      const int64_t rel = m.ii * f.ns0 + jo * f.ns1 + dk;
      if (f.noise != nullptr) val = __fadd_rn(val, __ldg(f.noise + rel));
      if (fl.z) val = fmaf(gb.w, adell_philox_normal(f.philox_seed, f.philox_offset + f.olin0 + rel), val);
      f.dst[m.ii * f.ds0 + jo * f.ds1 + dk * f.ds2] = val;
    }
  }
}

// Trilinear fp32 tiles whose only special feature is device noise (RandGaussianNoised of the SSL views: a fifth of them):
// the plain loops with the noise epilogue instead of the one-voxel-at-a-time cold loop.  Out of line: one call per tile.
__device__ __noinline__ void k1_tile_staged_noisy(const K1Tile& tl, const K1Fast& f, int rm) {
  if (rm == 0) k1_tile_staged_trilinear<K1_NP, 0, ADELL_F32, true>(tl, f);
  else if (rm == 1) k1_tile_staged_trilinear<K1_NP, 1, ADELL_F32, true>(tl, f);
  else k1_tile_staged_trilinear_scalar<K1_NV, 2, ADELL_F32, true>(tl, f);
}

// Plain staged tile (no output pad band, no noise) of a source of element type DT.  fp32 sources are inlined
// into the kernel's main loop; the integer-source variants are out of line (one call per tile) so that the three
// copies of the voxel loops do not share the instruction cache footprint of the hot one.
template <int DT>
__device__ __forceinline__ void k1_staged_plain_body(const K1Ctx& ctx, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box, int rm) {
  if (ctx.it.interp == ADELL_NEAREST) {
    if (rm == 0) k1_tile_staged_nearest<0, DT>(ctx, tl, f, box);
    else if (rm == 1) k1_tile_staged_nearest<1, DT>(ctx, tl, f, box);
    else if (rm == 3) k1_tile_staged_nearest<3, DT>(ctx, tl, f, box);
    else k1_tile_staged_nearest<2, DT>(ctx, tl, f, box);
  } else {
    if (rm == 0) k1_tile_staged_trilinear<K1_NP, 0, DT>(tl, f);
    else if (rm == 1) k1_tile_staged_trilinear<K1_NP, 1, DT>(tl, f);
    else if (rm == 3) k1_tile_staged_trilinear<K1_NP, 3, DT>(tl, f);
    else k1_tile_staged_trilinear_scalar<K1_NV, 2, DT>(tl, f);
  }
}
template <int DT>
__device__ __noinline__ void k1_staged_plain_ool(const K1Ctx& ctx, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box, int rm) {
  k1_staged_plain_body<DT>(ctx, tl, f, box, rm);
}
template <int DT>
__device__ __forceinline__ void k1_staged_plain(const K1Ctx& ctx, const K1Tile& tl, const K1Fast& f, const float* __restrict__ box, int rm) {
  if (DT == ADELL_F32) k1_staged_plain_body<DT>(ctx, tl, f, box, rm);
  else k1_staged_plain_ool<DT>(ctx, tl, f, box, rm);
}

// Identity items (flip / crop only, everything valid): the producer brings the tile's source box
// in with one TMA load like any staged tile — the load latency is carried by the TMA queue, several
// tiles deep, instead of by the consumer threads — and the consumers move it out with 128-bit
// shared loads and 128-bit streaming global stores.  One 32x16x32 tile = 64 KiB; each thread
// moves eight float4 (rows di0 + 4r of column quad q).  A flip is a sign in the box index; a flip
// along the contiguous axis reverses the quad in registers.
// quad of four consecutive source elements starting at box element index e -> four floats (memory order)
template <int DT>
__device__ __forceinline__ float4 k1_load_quad(const float* __restrict__ box, int e, bool vec) {
  if (DT == ADELL_F32) {
    const float* p = box + e;
    return vec ? *reinterpret_cast<const float4*>(p) : make_float4(p[0], p[1], p[2], p[3]);
  } else if (DT == ADELL_I16) {
    const short* p = reinterpret_cast<const short*>(box) + e;
    if (vec) {  // 8-byte aligned: one 64-bit load
      const short4 q = *reinterpret_cast<const short4*>(p);
      return make_float4(static_cast<float>(q.x), static_cast<float>(q.y), static_cast<float>(q.z), static_cast<float>(q.w));
    }
    return make_float4(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2]), static_cast<float>(p[3]));
  } else {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(box) + e;
    if (vec) {
      const uchar4 q = *reinterpret_cast<const uchar4*>(p);
      return make_float4(static_cast<float>(q.x), static_cast<float>(q.y), static_cast<float>(q.z), static_cast<float>(q.w));
    }
    return make_float4(static_cast<float>(p[0]), static_cast<float>(p[1]), static_cast<float>(p[2]), static_cast<float>(p[3]));
  }
}

template <int DT>
__device__ __forceinline__ void k1_tile_copy_box(const K1Ctx& c, const K1Tile& tl, const float* __restrict__ box) {
  const adell_item& it = c.it;
  constexpr int PS = K1_GTHREADS / 128;  // planes a group covers per step (a plane = 16 rows x 8 quads)
  const int gtid = threadIdx.x % K1_GTHREADS;
  if (gtid >= 128 * PS) return;          // (a group that is not a multiple of 128 threads: the rest would repeat rows)
  const int q = gtid & 7, dj = (gtid >> 3) & 15, di0 = gtid >> 7;
  const int o0 = tl.o0[0] + di0, o1 = tl.o0[1] + dj, o2 = tl.o0[2] + 4 * q;
  const int O0 = it.out_shape[0];
  if (o1 >= it.out_shape[1] || o2 >= it.out_shape[2] || o0 >= O0) return;
  // box index of output voxel o along axis a: s_a*o_a + M_a
  const int s0 = tl.msign[0] * it.grid_sign[0], s1 = tl.msign[1] * it.grid_sign[1], s2 = tl.msign[2] * it.grid_sign[2];
  const int M0 = tl.msign[0] * it.grid_off[0] + tl.mconst[0], M1 = tl.msign[1] * it.grid_off[1] + tl.mconst[1],
            M2 = tl.msign[2] * it.grid_off[2] + tl.mconst[2];
  const bool rev = s2 < 0;
  const int m2 = rev ? M2 - (o2 + 3) : M2 + o2;  // lowest box column of the quad
  const int p1 = tl.box[2], p0 = tl.box[1] * tl.box[2];
  const int sp = (s0 * o0 + M0) * p0 + (s1 * o1 + M1) * p1 + m2;   // box element index of the thread's first quad
  const int sstep = PS * s0 * p0;
  const bool vec = (m2 & 3) == 0;  // block-uniform: the same for every quad of every tile of an item
  const bool clip = (it.flags & ADELL_F_CLIP) != 0;
  const float pre_s = c.pre_s, pre_o = c.pre_o, post_s = it.post_scale, post_o = it.post_offset;
  const float clo = it.clip_lo, chi = it.clip_hi;
  const float gain = pre_s * post_s, bias = fmaf(pre_o, post_s, post_o);
  const bool plain = !clip && gain == 1.0f && bias == 0.0f;
  const int64_t dstep = PS * it.dst_stride[0];
  float* dp = it.dst + o0 * it.dst_stride[0] + o1 * it.dst_stride[1] + o2;
  const int nrow = (min(O0 - o0, tl.T[0] - di0) + PS - 1) / PS;  // rows o0 + PS*r of this thread inside the tile
  auto fix = [&](float4 x) {
    if (rev) { float t = x.x; x.x = x.w; x.w = t; t = x.y; x.y = x.z; x.z = t; }
    if (plain) return x;
    if (clip) {
      x.x = fmaf(k1_premap(x.x, pre_s, pre_o, true, clo, chi), post_s, post_o);
      x.y = fmaf(k1_premap(x.y, pre_s, pre_o, true, clo, chi), post_s, post_o);
      x.z = fmaf(k1_premap(x.z, pre_s, pre_o, true, clo, chi), post_s, post_o);
      x.w = fmaf(k1_premap(x.w, pre_s, pre_o, true, clo, chi), post_s, post_o);
    } else {
      x.x = fmaf(x.x, gain, bias); x.y = fmaf(x.y, gain, bias); x.z = fmaf(x.z, gain, bias); x.w = fmaf(x.w, gain, bias);
    }
    return x;
  };
  if (vec) {
    // eight independent vector loads in flight per thread, then their stores
#pragma unroll 1
    for (int r0 = 0; r0 < nrow; r0 += 8) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r0 + r < nrow)
          K1_STORE(reinterpret_cast<float4*>(dp + (r0 + r) * dstep), fix(k1_load_quad<DT>(box, sp + (r0 + r) * sstep, true)));
    }
  } else {
#pragma unroll 2
    for (int r = 0; r < nrow; ++r)
      K1_STORE(reinterpret_cast<float4*>(dp + r * dstep), fix(k1_load_quad<DT>(box, sp + r * sstep, false)));
  }
}

// Tile whose every tap lies outside the valid source (zeros padding: the corners of a rotated volume):
// each voxel is post(0) (+ noise).  Without noise that is one constant per item: the tile is a fill with
// the item's fields held in registers (the generic per-voxel epilogue re-reads them from shared memory
// for every voxel: ~9 000 cycles per tile against ~1 000 here; a quarter of config E's tiles are such tiles).
__device__ __forceinline__ void k1_tile_zero(const K1Ctx& c, const K1Tile& tl) {
  const adell_item& it = c.it;
  const bool strict = (it.flags & ADELL_F_STRICT) != 0;
  if (it.noise != nullptr || (it.flags & ADELL_F_PHILOX)) {
    k1_for_each_voxel(tl, it, [&](int, int, int, int o0, int o1, int o2) { k1_finish(it, 0.0f, o0, o1, o2, strict); });
    return;
  }
  float val = 0.0f;
  if (strict) {
    if (it.post_scale != 1.0f) val = __fmul_rn(val, it.post_scale);
    if (it.post_offset != 0.0f) val = __fadd_rn(val, it.post_offset);
  } else {
    val = fmaf(val, it.post_scale, it.post_offset);
  }
  const int T1 = tl.T[1], T2 = tl.T[2];
  const int s2 = __ffs(T2) - 1, s1 = __ffs(T1) - 1;
  const int nvox = tl.T[0] << (s1 + s2);
  const int b0 = tl.o0[0], b1 = tl.o0[1], b2 = tl.o0[2];
  const int O0 = it.out_shape[0], O1 = it.out_shape[1], O2 = it.out_shape[2];
  const int64_t ds0 = it.dst_stride[0], ds1 = it.dst_stride[1], ds2 = it.dst_stride[2];
  float* __restrict__ dst = it.dst;
  // the lane's column (and with it its column group's shifts) is the same in every iteration: T2 divides the group size
  const int gt = threadIdx.x % K1_GTHREADS;
  const int dk = gt & (T2 - 1), o2 = b2 + dk;
  if (o2 >= O2) return;
  const int G = (o2 >> 3) & 15;
  const int sh0 = it.shear[0][G], sh1 = it.shear[1][G];
  float* __restrict__ col = dst + o2 * ds2;
#pragma unroll 4
  for (int v = gt; v < nvox; v += K1_GTHREADS) {
    const int dj = (v >> s2) & (T1 - 1), di = v >> (s1 + s2);
    const int o0 = b0 + di - sh0, o1 = b1 + dj - sh1;
    if (o0 < 0 || o1 < 0 || o0 >= O0 || o1 >= O1) continue;
    col[o0 * ds0 + o1 * ds1] = val;
  }
}

// ------------------------------------------------------------------------- per-tile set-up
struct K1Slot {
  K1Ctx ctx;
  K1Tile tl;
  K1Fast fast;
};

#ifdef K1_PROFILE
// producer sub-phases (debug builds): cycles of lane 0 between checkpoints, summed over tiles
__device__ unsigned long long k1_prof2[16];
#define K1_P2_DECL long long _p2 = clock64();
#define K1_P2(i) { if (lane == 0) { const long long _t = clock64(); atomicAdd(&k1_prof2[i], static_cast<unsigned long long>(_t - _p2)); _p2 = _t; } else { _p2 = clock64(); } }
#else
#define K1_P2_DECL
#define K1_P2(i)
#endif

// Warp-parallel: lanes 0..2 own one source axis each (footprint, padding fold, box placement,
// fast coordinates), lane 3 fills the scalar part of the consumers' register image.  Everything
// per-item (the fp64 coordinate of output voxel 0 and its derivative, the footprint of a full
// tile) was computed on the host by adell_aug_prepare.
// The column-group terms of a tile depend on the item and on the tile's position along axis 2 only: the producer keeps
// those of its last (item, b2) in registers (a 32-deep volume has ONE tile along axis 2: every tile of the item reuses
// them).  The group loop was the longest phase of the set-up (profiles/r02_k1_ncu_summary.md: ~1 800 of ~8 500 cycles).
struct K1GroupCache {
  int item = -1, b2 = -1;
  double emin = 0.0, emax = 0.0;
  float egv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int shv[4] = {0, 0, 0, 0};
  int sa_max = 0;
};

__device__ __forceinline__ void k1_tile_setup(const K1Ctx& c, K1Slot& sl, int b0, int b1, int b2, uint32_t box_addr, int lane,
                                              K1GroupCache& gc, int item_id, bool slot_fresh) {
  const adell_item& it = c.it;
  K1Tile& tl = sl.tl;
  const unsigned FULL = 0xffffffffu;
  const bool ax = lane < 3;
  const int a = ax ? lane : 0;
  K1_P2_DECL
  const int o00 = b0 * it.tile_dim[0], o01 = b1 * it.tile_dim[1], o02 = b2 * it.tile_dim[2];
  const int o0a = a == 0 ? o00 : (a == 1 ? o01 : o02);
  if (ax) { tl.o0[a] = o0a; tl.T[a] = it.tile_dim[a]; }
  if (it.kind >= ADELL_KIND_VCOPY) {
    // source box of the tile = its voxels, in memory order (integer flip / crop only)
    if (ax) {
      const int na = min(static_cast<int>(it.tile_dim[a]), it.out_shape[a] - o0a);
      const int g0 = it.grid_off[a] + it.grid_sign[a] * o0a, g1 = it.grid_off[a] + it.grid_sign[a] * (o0a + na - 1);
      const int msign = it.tmap_sign[a];
      int mo = msign > 0 ? min(g0, g1) + it.tmap_off[a] : -max(g0, g1) + it.tmap_off[a];
      if (a == 2) mo &= ~(static_cast<int>(16u / k1_es(it.src_dtype)) - 1);  // 16-byte aligned start along the contiguous axis
      tl.box[a] = it.tmap_box[a]; tl.msign[a] = msign; tl.mconst[a] = it.tmap_off[a] - mo;
      if (a == 2) { tl.fix_lo = 0; tl.fix_hi = 0; }
    }
    if (lane == 0) tl.mode = it.kind == ADELL_KIND_VCOPY ? MODE_COPY : MODE_TSTORE;
    return;
  }
  if (it.kind != ADELL_KIND_STAGED) {
    if (lane == 0) tl.mode = MODE_DIRECT;
    return;
  }
  // un-padded source coordinate of the tile-origin voxel along axis a, and the tile's footprint
  const double U0 = fma(it.fp_D[3 * a + 2], static_cast<double>(o02),
                        fma(it.fp_D[3 * a + 1], static_cast<double>(o01),
                            fma(it.fp_D[3 * a + 0], static_cast<double>(o00), it.fp_U0[a])));
  K1_P2(0)   // tile origin coordinate (fp64)
  // column groups of this tile: group term e(g) = D_a2*8g - D_a0*s0(G) - D_a1*s1(G) (see k1_item_shear);
  // fp_smin/fp_smax hold the footprint of ONE group (di, dj over the tile, 8 voxels along axis 2)
  if (gc.item != item_id || gc.b2 != b2) {   // (warp-uniform)
    const int G0 = o02 >> 3;
    const int ng = min(static_cast<int>(it.tile_dim[2]) >> 3, (it.out_shape[2] - o02 + 7) >> 3);
    gc.item = item_id; gc.b2 = b2;
    gc.emin = 0.0; gc.emax = 0.0; gc.sa_max = 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      gc.egv[g] = 0.0f; gc.shv[g] = 0;
      if (g < ng) {
        const int s0 = it.shear[0][(G0 + g) & 15], s1 = it.shear[1][(G0 + g) & 15];
        gc.sa_max = max(gc.sa_max, a == 0 ? s0 : (a == 1 ? s1 : 0));
        const double sh = -(it.fp_D[3 * a + 0] * s0 + it.fp_D[3 * a + 1] * s1);
        const double ev = fma(it.fp_D[3 * a + 2], 8.0 * g, sh);
        gc.emin = g == 0 ? ev : fmin(gc.emin, ev);
        gc.emax = g == 0 ? ev : fmax(gc.emax, ev);
        gc.egv[g] = static_cast<float>(sh);   // -D_a0*s0(G) - D_a1*s1(G): added to the group's fast coordinates
        gc.shv[g] = s0 | (s1 << 16);
      }
    }
  }
  const double emin = gc.emin, emax = gc.emax;
  float egv[4] = {gc.egv[0], gc.egv[1], gc.egv[2], gc.egv[3]};
  const int shv[4] = {gc.shv[0], gc.shv[1], gc.shv[2], gc.shv[3]};
  const int sa_max = gc.sa_max;  // largest shift of this tile's groups along output axis a (axes 0, 1)
  K1_P2(1)   // column-group loop
  const double umin = U0 + emin + static_cast<double>(it.fp_smin[a]), umax = U0 + emax + static_cast<double>(it.fp_smax[a]);
  const bool finite = (umin > -1.0e6) && (umax < 1.0e6);
  if (!__all_sync(FULL, finite)) { if (lane == 0) tl.mode = MODE_DIRECT; return; }
  const int S = it.src_shape[a];
  const int lo = static_cast<int>(floor(umin - K1_EPS)), hi = static_cast<int>(floor(umax + K1_EPS)) + 1;
  int blo = lo, bhi = hi;  // source index interval the box must hold
  bool rm = false, anyv = true;
  if (it.padding == ADELL_PAD_ZEROS) {
    if (hi < c.tlo[a] || lo >= c.thi[a]) anyv = false;
  } else if (lo < 0 || hi > S - 1) {
    // coordinates leave [0,S): the voxel loop clamps / reflects them (ATen semantics), so the box
    // only has to hold the covering interval of the clamped / reflected indices (+1 for the hi tap)
    rm = true;
    int rlo, rhi;
    if (it.padding == ADELL_PAD_BORDER) {
      // a coordinate clamped onto cell 0 still READS cell 1 (weight 0): a tile that lies entirely below the
      // volume needs [0, 1], not [0, 0] — on an axis that runs backwards in memory cell 1 would otherwise sit
      // at box index -1 (an illegal address for the bit-faithful path; found by tools/fuzz_parity.py)
      rlo = min(max(lo, 0), S - 1); rhi = min(max(hi, min(1, S - 1)), S - 1);
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
    // the fast path clamps u' just below S-1: the hi tap never reads cell S, and the lo tap of a
    // coordinate clamped at the upper edge is S-2
    blo = min(rlo, max(S - 2, 0)); bhi = rhi + (S == 1 ? 1 : 0);
    // an axis that runs backwards in memory (box index = bhi - t): a coordinate clamped / reflected onto the lowest
    // cell blo puts the lo tap on the LAST needed box index and the hi tap (weight 0, still read) one cell past it.
    // Stage that cell too (t = blo - 1: real data, or TMA's zero fill outside the volume) — the mirror image of the
    // clamp one ulp below S-1 on forward axes, at no cost in the voxel loops.
    if (it.tmap_sign[a] < 0) blo -= 1;
  }
  // cells [blo, bhi] plus, along the contiguous axis, those lost to aligning the box origin down to 16 bytes
  const int es = static_cast<int>(k1_es(it.src_dtype));
  const int qm = 16 / es - 1;   // elements per 16 bytes, minus one
  const int mo_raw = it.tmap_sign[a] > 0 ? blo + it.tmap_off[a] : -bhi + it.tmap_off[a];
  const bool fits = bhi - blo + 1 + (a == 2 ? (mo_raw & qm) : 0) <= it.tmap_box[a];  // else larger than the encoded box
  const bool allv = lo >= c.tlo[a] && hi < c.thi[a];
  const bool win_full = c.tlo[a] <= 0 && c.thi[a] >= S;
  if (!__all_sync(FULL, fits)) { if (lane == 0) tl.mode = MODE_DIRECT; return; }
  if (!__all_sync(FULL, anyv)) { if (lane == 0) tl.mode = MODE_ZERO; return; }
  K1_P2(2)   // footprint interval, padding case analysis, fit checks
  const int rmask = static_cast<int>(__ballot_sync(FULL, ax && rm) & 7u);
  const bool allv_all = __all_sync(FULL, allv), win_all = __all_sync(FULL, win_full);
  // a pre offset must not leak into zero-filled (invalid) taps: such tiles use the exact path
  const int all_valid = rmask ? (win_all ? 1 : 0) : (allv_all ? 1 : 0);
  const int msign = it.tmap_sign[a];
  int mo = msign > 0 ? blo + it.tmap_off[a] : -bhi + it.tmap_off[a];
  // TMA needs a 16-byte aligned start along the contiguous axis: round the box origin down to a
  // multiple of 16 bytes (the encoded inner extent carries the spare elements for this)
  if (a == 2) mo &= ~qm;
  const int mconst = it.tmap_off[a] - mo;
  float V0, Dm0, Dm1, Dm2, rA, rB;
  if (rm) {
    // reflection is periodic in 2S: move the tile's coordinates next to the origin (whole periods,
    // exact in fp64) so that the fold's fp32 arithmetic keeps its precision far from the volume
    double sh = 0.0;
    if (it.padding == ADELL_PAD_REFLECTION) sh = 2.0 * S * rint((0.5 * (umin + umax) + 0.5) / (2.0 * S));
    V0 = static_cast<float>(U0 - sh);  // (group terms egv stay in source orientation, like Dm)
    Dm0 = static_cast<float>(it.fp_D[3 * a + 0]); Dm1 = static_cast<float>(it.fp_D[3 * a + 1]); Dm2 = static_cast<float>(it.fp_D[3 * a + 2]);
    rA = static_cast<float>(msign); rB = static_cast<float>(mconst);
  } else {
    V0 = static_cast<float>(msign * U0 + mconst);
#pragma unroll
    for (int g = 0; g < 4; ++g) egv[g] *= static_cast<float>(msign);
    Dm0 = static_cast<float>(msign * it.fp_D[3 * a + 0]); Dm1 = static_cast<float>(msign * it.fp_D[3 * a + 1]);
    Dm2 = static_cast<float>(msign * it.fp_D[3 * a + 2]);
    rA = 1.0f; rB = 0.0f;
  }
  K1_P2(3)   // votes, box placement, fast coordinate of the tile origin
  K1Fast& f = sl.fast;
  const int na = min(static_cast<int>(it.tile_dim[a]), it.out_shape[a] - o0a);
  const int vlo = it.out_vlo[a] - o0a, vhi = it.out_vhi[a] - o0a;
  // does any voxel of the tile (its window may be shifted down by up to sa_max) lie in an output pad band?
  const bool padded_a = it.out_vlo[a] > max(0, o0a - sa_max) || it.out_vhi[a] < min(it.out_shape[a], o0a + static_cast<int>(it.tile_dim[a]));
  const bool padded = __any_sync(FULL, ax && padded_a);
  const int n0 = __shfl_sync(FULL, na, 0), n1 = __shfl_sync(FULL, na, 1), n2 = __shfl_sync(FULL, na, 2);
  const float rB0 = __shfl_sync(FULL, rB, 0), rB1 = __shfl_sync(FULL, rB, 1), rB2 = __shfl_sync(FULL, rB, 2);
  // tie window of the nearest fast path: grows with the magnitude of the tile's (un-padded) coordinates,
  // where the reference's own fp32 chain is coarser (far translations under border / reflection)
  const float mag_a = ax ? fmaxf(fabsf(static_cast<float>(umin)), fabsf(static_cast<float>(umax))) : 0.0f;
  const float mag = fmaxf(fmaxf(__shfl_sync(FULL, mag_a, 0), __shfl_sync(FULL, mag_a, 1)), __shfl_sync(FULL, mag_a, 2));
  const float tie = fminf(c.tie, 0.5f - 2.0e-6f * mag);
  const int vl0 = __shfl_sync(FULL, vlo, 0), vl1 = __shfl_sync(FULL, vlo, 1), vl2 = __shfl_sync(FULL, vlo, 2);
  const int vh0 = __shfl_sync(FULL, vhi, 0), vh1 = __shfl_sync(FULL, vhi, 1), vh2 = __shfl_sync(FULL, vhi, 2);
  K1_P2(4)   // shuffles
  if (ax) {
    tl.lo_t[a] = lo; tl.hi_t[a] = hi;
    tl.box[a] = it.tmap_box[a]; tl.msign[a] = msign; tl.mconst[a] = mconst;
    tl.V0[a] = V0; tl.Dm[a][0] = Dm0; tl.Dm[a][1] = Dm1; tl.Dm[a][2] = Dm2; tl.rA[a] = rA; tl.rB[a] = rB;
    f.ax[a] = make_float4(V0, Dm0, Dm1, Dm2);
#pragma unroll
    for (int g = 0; g < 4; ++g) f.eg[g][a] = egv[g];
    // clamp just below S-1 (one ulp): floor(u') + 1 <= S-1, the lerp error is <= 2^-23 of the step
    if (rm) {
      f.rf[a] = make_float4(0.5f / c.Sf[a], 2.0f * c.Sf[a], c.Sm1[a] > 0.0f ? __uint_as_float(__float_as_uint(c.Sm1[a]) - 1u) : 0.0f, rA);
    } else {
      // box-local interval [m_lo, m_hi] of the valid taps (m = msign * t + mconst, t in [tlo, thi - 1])
      const int ma = msign * c.tlo[a] + mconst, mb = msign * (c.thi[a] - 1) + mconst;
      f.rf[a] = make_float4(static_cast<float>(min(ma, mb) - 1), static_cast<float>(max(ma, mb) + 1), 0.0f, 0.0f);
    }
    if (a == 2) {  // alignment slack columns of the tensor map that fall inside this box
      tl.fix_lo = max(0, -mo);
      tl.fix_hi = it.fp_fix > 0 ? min(it.tmap_box[2], it.fp_fix - mo) : 0;
    }
  }
  if (lane == 3) {
    tl.mode = MODE_STAGED;
    tl.rmask = rmask; tl.all_valid = all_valid;
    const int philox = (it.flags & ADELL_F_PHILOX) ? 1 : 0;
    if (slot_fresh) {   // the item-constant part: written when the slot receives this item's context (see k1_prepare)
      f.gb = make_float4(c.pre_s * it.post_scale, fmaf(c.pre_o, it.post_scale, it.post_offset), it.post_offset, it.noise_std);
      f.ws = make_float4(c.pre_o * it.post_scale, it.post_offset, 0.0f, 0.0f);
      f.ds0 = it.dst_stride[0]; f.ds1 = it.dst_stride[1]; f.ds2 = it.dst_stride[2];
      f.ns1 = it.out_shape[2]; f.ns0 = static_cast<int64_t>(it.out_shape[1]) * it.out_shape[2];
      f.philox_seed = it.philox_seed; f.philox_offset = it.philox_offset;
    }
    f.n = make_int4(it.tile_dim[0], it.tile_dim[1], n2, it.tile_dim[2]);
    f.lim = make_int4(-o00, it.out_shape[0] - o00, -o01, it.out_shape[1] - o01);
#pragma unroll
    for (int g = 0; g < 4; ++g) f.eg[g][3] = __int_as_float(shv[g]);
    const uint32_t p0 = static_cast<uint32_t>(it.tmap_box[1] * it.tmap_box[2]), p1 = static_cast<uint32_t>(it.tmap_box[2]);
    f.m = make_int4(static_cast<int>(p0), static_cast<int>(p1), rmask | (it.padding << 8),
                    static_cast<int>(box_addr - static_cast<uint32_t>(es) * (K1_MAGIC_BITS * (p0 + p1 + 1u))));
    f.rb = make_float4(rB0, rB1, rB2, tie);
    f.fl = make_int4((padded || it.noise != nullptr || philox) ? 1 : 0, padded ? 1 : 0, philox, 0);
    f.vlo = make_int4(vl0, vl1, vl2, 0);
    f.vhi = make_int4(vh0, vh1, vh2, 0);
    f.dst = it.dst + o00 * it.dst_stride[0] + o01 * it.dst_stride[1] + o02 * it.dst_stride[2];
    const uint64_t olin0 = (static_cast<uint64_t>(o00) * it.out_shape[1] + o01) * it.out_shape[2] + o02;
    f.olin0 = olin0;
    f.noise = it.noise ? it.noise + olin0 : nullptr;
  }
  K1_P2(5)   // stores of the tile state and the consumers' register image
}

#ifdef K1_PROFILE
__device__ unsigned long long k1_prof[16];
// cycle counters accumulate in registers and are flushed once per warp (low perturbation)
#define K1_PROF_DECL long long _acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define K1_PROF_T0 long long _t0 = clock64();
#define K1_PROF_ADD(i) { long long _t1 = clock64(); _acc[i] += _t1 - _t0; _t0 = _t1; }
#define K1_PROF_FLUSH if ((threadIdx.x & 31) == 0) { for (int _i = 0; _i < 16; ++_i) if (_acc[_i]) atomicAdd(&k1_prof[_i], (unsigned long long)_acc[_i]); }
#else
#define K1_PROF_DECL
#define K1_PROF_T0
#define K1_PROF_ADD(i)
#define K1_PROF_FLUSH
#endif

// Producer: derive the tile state (and the consumers' register image) in the given slot.  The
// item itself is fetched from global memory only when the tile sequence moves on to a new item.
struct K1Walk {
  int prev_tile = -2;      // last tile this producer prepared
  int b0 = 0, b1 = 0, b2 = 0;  // its tile coordinates inside the item
  int fresh = 0;           // slots of the stream's ring that already hold the current item's context
  K1GroupCache gc;         // column-group terms of the last (item, tile position along axis 2)
  unsigned const_mask = 0; // ring slots whose register image already holds the current item's constant part
};
__device__ __forceinline__ void k1_prepare(const adell_item* __restrict__ items, const int32_t* ts, int n_items,
                                           int tile, int& item, int& cur_start, int& next_start, int& cached_item,
                                           K1Ctx& priv, K1Slot& sl, uint32_t box_addr, int lane, K1Walk& wk, int ring_slots, int slot_pos) {
  K1_P2_DECL
  // monotone walk over the per-item tile prefix, 32 entries per step (one load latency per step,
  // not one per item skipped); `ts` is the shared-memory copy of the prefix when it fits
  while (tile >= next_start) {
    const int idx = item + 1 + lane;
    const int v = idx <= n_items ? ts[idx] : 0x7fffffff;
    const int adv = __popc(__ballot_sync(0xffffffffu, tile >= v));  // prefix is non-decreasing: a run of leading lanes
    item += adv;
    cur_start = __shfl_sync(0xffffffffu, v, adv - 1);
    const int nxt = __shfl_sync(0xffffffffu, v, adv & 31);
    next_start = adv < 32 ? nxt : (item + 1 <= n_items ? ts[item + 1] : 0x7fffffff);
  }
  K1_P2(8)   // walk over the tile prefix
  constexpr int kItemWords = (sizeof(adell_item) - 128) / 4;  // without the tensor map
  uint32_t* pw = reinterpret_cast<uint32_t*>(&priv) + 32;
  const bool new_item = item != cached_item;
  if (new_item) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(items + item) + 32;
#pragma unroll
    for (int w = 0; w < (kItemWords + 31) / 32; ++w)
      if (w * 32 + lane < kItemWords) pw[w * 32 + lane] = __ldg(src + w * 32 + lane);
    __syncwarp();
    if (lane == 0) k1_ctx_finish(priv);
    __syncwarp();
    cached_item = item;
    wk.fresh = 0;
    wk.const_mask = 0u;
  }
  // whole K1Ctx (item image + derived constants), minus the unused tensor-map bytes — only while some
  // slot of the ring (visited round-robin) still holds another item's context: an item has hundreds of tiles
  if (wk.fresh < ring_slots) {
    constexpr int kWords = (sizeof(K1Ctx) - 128) / 4;
    uint32_t* dw = reinterpret_cast<uint32_t*>(&sl.ctx) + 32;
#pragma unroll
    for (int w = 0; w < (kWords + 31) / 32; ++w)
      if (w * 32 + lane < kWords) dw[w * 32 + lane] = pw[w * 32 + lane];
    ++wk.fresh;
    __syncwarp();
  }
  K1_P2(9)   // item fetch / context copies
  // tile coordinates inside the item: the successor of the previous tile by carries, else two divisions
  const int n1 = priv.it.n_tiles[1], n2 = priv.it.n_tiles[2];
  if (!new_item && tile == wk.prev_tile + 1 && tile > cur_start) {
    if (++wk.b2 == n2) { wk.b2 = 0; if (++wk.b1 == n1) { wk.b1 = 0; ++wk.b0; } }
  } else {
    int local = tile - cur_start;
    wk.b2 = local % n2; local /= n2;
    wk.b1 = local % n1;
    wk.b0 = local / n1;
  }
  wk.prev_tile = tile;
  K1_P2(10)  // tile coordinates
  const bool need_const = !(wk.const_mask >> slot_pos & 1u);
#ifdef K1_EXP_FREE_PRODUCER
  // TIMING EXPERIMENT (wrong voxels): 7 of 8 tiles re-use the state this slot holds of an earlier tile of the same item
  if (!need_const && (tile & 7) != 0) { if (lane == 0) sl.tl.next_plane = 0; __syncwarp(); return; }
#endif
  k1_tile_setup(sl.ctx, sl, wk.b0, wk.b1, wk.b2, box_addr, lane, wk.gc, item, need_const);
  if (lane == 0) { sl.tl.item = item; sl.tl.next_plane = 0; }
  __syncwarp();
  if (need_const && sl.tl.mode == MODE_STAGED) wk.const_mask |= 1u << slot_pos;   // (only a staged tile writes the image)
}

// Producer, after the TMA load of a tile whose box contains alignment-slack columns completed:
// zero them (they hold bytes that precede the valid source box) and hand the stage to the consumers.
__device__ __forceinline__ void k1_fix_columns(float* box, const K1Tile& tl, int lane, int dt) {
  const int w = tl.fix_hi - tl.fix_lo, p1 = tl.box[2];
  const int rows = tl.box[0] * tl.box[1];
  if (dt == ADELL_F32) {
    for (int e = lane; e < rows * w; e += 32) box[(e / w) * p1 + tl.fix_lo + e % w] = 0.0f;
  } else if (dt == ADELL_I16) {
    short* b = reinterpret_cast<short*>(box);
    for (int e = lane; e < rows * w; e += 32) b[(e / w) * p1 + tl.fix_lo + e % w] = 0;
  } else {
    unsigned char* b = reinterpret_cast<unsigned char*>(box);
    for (int e = lane; e < rows * w; e += 32) b[(e / w) * p1 + tl.fix_lo + e % w] = 0;
  }
  fence_proxy_async_smem();
  __syncwarp();
}

// Persistent, warp-specialised, dynamically scheduled.  One CTA per SM runs K1_GROUPS independent
// streams.  Stream p: its producer warp pulls chunks of `chunk` consecutive tiles from a global queue
// (an atomic counter in the caller's launch buffer: SMs that run faster simply take more chunks, so
// no SM idles at the tail), prepares the tile state and issues one TMA box load per tile into the
// stream's next free ring stage; its hand-over warp sits between the TMA completion and the
// consumers (alignment-slack columns); its K1_GWARPS consumer warps produce the voxels.  Per stream:
// n_stages / K1_GROUPS ring stages (full / empty / landed mbarriers each) and one more tile-state
// slot than stages, so set-up never waits for shared-memory space.  A tile with mode MODE_DONE ends
// the stream.
struct K1Ring {
  int n, i, phase;   // stages (or slots) of the stream, position, phase bit of the current lap
  __device__ __forceinline__ void init(int n_) { n = n_; i = 0; phase = 0; }
  __device__ __forceinline__ void next() { if (++i == n) { i = 0; phase ^= 1; } }
};

__global__ void __launch_bounds__(K1_THREADS, 1)
k1_gather(const adell_item* __restrict__ items, const int32_t* __restrict__ tile_start, int n_items, int total_tiles,
          int n_stages, int stage_bytes, int chunk, int n_big_r, int n_big_c, int first_copy,
          unsigned int* __restrict__ sched) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int n_slots = n_stages + K1_GROUPS;
  K1Slot* slots = reinterpret_cast<K1Slot*>(smem + static_cast<size_t>(n_stages) * stage_bytes);
  K1Ctx* priv = reinterpret_cast<K1Ctx*>(slots + n_slots);   // the producers' private item copies
  uint64_t* full = reinterpret_cast<uint64_t*>(priv + K1_GROUPS);
  uint64_t* empty = full + n_stages;
  uint64_t* landed = empty + n_stages;   // TMA completion, seen by the hand-over warp first
  int32_t* ts_s = reinterpret_cast<int32_t*>(landed + n_stages);
  const bool ts_cached = n_items + 1 <= K1_TS_CACHE;
  if (ts_cached)
    for (int i = threadIdx.x; i <= n_items; i += K1_THREADS) ts_s[i] = __ldg(tile_start + i);
  const int32_t* ts = ts_cached ? ts_s : tile_start;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + s)), "r"(K1_GWARPS));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(landed + s)), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // stream of this warp, its stages p, p + K1_GROUPS, ... and slots p, p + K1_GROUPS, ...
  const int strm = warp < K1_CWARPS ? warp / K1_GWARPS : (warp - K1_CWARPS) % K1_GROUPS;
  K1Ring rs, rl;
  rs.init(n_stages / K1_GROUPS);
  rl.init(n_slots / K1_GROUPS);

  if (warp >= K1_CWARPS + K1_GROUPS) {
    // ------------------------------------------------------------------ producer warp of stream `strm`
    int item = 0, cached_item = -1, cur_start = 0;
    int next_start = ts[1];
    int acq_item = -1;                  // item whose tensor map this warp acquired last
    int64_t cur = 0, cur_end = 0;       // tiles left of the current chunk
    // Two queues: 0 = tiles of resampled / generic items [0, first_copy), 1 = tiles of box-copy items
    // [first_copy, total).  Stream 0 drains the resampled queue first, stream 1 the copy queue, each
    // moving on to the other queue when its own is empty: most of the time an SM then works on a
    // compute-bound and a memory-bound tile at once instead of all SMs going through the same phases.
    int q = strm & 1;
    bool switched = false;
    K1Walk wk;
    const bool idle = (chunk >> 16) == strm + 1;   // measurement aid (ADELL_K1_IDLE_STREAM)
    chunk &= 0xffff;
    K1_PROF_DECL
    for (;; rs.next(), rl.next()) {
      const int stage = strm + K1_GROUPS * rs.i, slot = strm + K1_GROUPS * rl.i;
      K1_PROF_T0
      while (cur >= cur_end) {          // next unit from the current queue, or from the other one
        if (idle) { cur = cur_end = total_tiles; break; }
        unsigned int c = 0;
        if (lane == 0) c = atomicAdd(sched + q, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        const int64_t lo = q == 0 ? 0 : first_copy, hi = q == 0 ? first_copy : total_tiles;
        const unsigned int nb = static_cast<unsigned int>(q == 0 ? n_big_r : n_big_c);
        // the first nb units of a queue are chunks of `chunk` tiles, the rest single tiles (balanced tail)
        if (c < nb) { cur = lo + static_cast<int64_t>(c) * chunk; cur_end = cur + chunk; }
        else { cur = lo + static_cast<int64_t>(nb) * chunk + (c - nb); cur_end = cur + 1; }
        cur_end = min(cur_end, hi);
        if (cur < hi) break;
        if (switched) { cur = cur_end = total_tiles; break; }   // both queues drained
        switched = true;
        q ^= 1;
        item = 0; cur_start = 0; next_start = ts[1];            // the other queue's tiles may lie behind: restart the walk
        cur = cur_end = 0;
      }
      if (cur >= total_tiles || cur >= cur_end) {         // queues drained: end-of-stream marker
        if (lane == 0) slots[slot].tl.mode = MODE_DONE;
        __syncwarp();
        mbar_wait_relaxed(empty + stage, rs.phase ^ 1);
        if (lane == 0) mbar_arrive(landed + stage);
        break;
      }
      const int tile = static_cast<int>(cur++);
      // safe to overwrite: the stream has one more slot than stages, and this warp's previous issue
      // waited for the release of the stage of the slot's previous tile
      k1_prepare(items, ts, n_items, tile, item, cur_start, next_start, cached_item, priv[strm], slots[slot],
                 smem_u32(smem + static_cast<size_t>(stage) * stage_bytes), lane, wk, rl.n, rl.i);
      K1_PROF_ADD(2)
      mbar_wait_relaxed(empty + stage, rs.phase ^ 1);
      K1_PROF_ADD(0)
      if (lane == 0) {
        const K1Slot& sl = slots[slot];
        if (sl.tl.mode == MODE_STAGED || sl.tl.mode == MODE_COPY || sl.tl.mode == MODE_TSTORE) {  // tiles with a TMA box load
          const int mo0 = sl.ctx.it.tmap_off[0] - sl.tl.mconst[0], mo1 = sl.ctx.it.tmap_off[1] - sl.tl.mconst[1],
                    mo2 = sl.ctx.it.tmap_off[2] - sl.tl.mconst[2];
          if (sl.tl.item != acq_item) { tmap_acquire(items[sl.tl.item].tmap); acq_item = sl.tl.item; }
          mbar_expect_tx(landed + stage, static_cast<uint32_t>(sl.tl.box[0] * sl.tl.box[1] * sl.tl.box[2]) * k1_es(sl.ctx.it.src_dtype));
          tma_load_3d(smem + static_cast<size_t>(stage) * stage_bytes, items[sl.tl.item].tmap, landed + stage, mo2, mo1, mo0);
        } else {
          mbar_arrive(landed + stage);
        }
      }
      __syncwarp();
      K1_PROF_ADD(1)
    }
    // the last producer of the grid to drain the queue re-arms it for the next launch of this buffer
    if (lane == 0) {
      const unsigned int done = atomicAdd(sched + 2, 1u);
      if (done == K1_GROUPS * gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; sched[2] = 0u; }
    }
    K1_PROF_FLUSH
    return;
  }

  if (warp >= K1_CWARPS) {
    // ------------------------------------------------------------------ hand-over warp of stream `strm`
    // Sits between the TMA completion (landed) and the consumers (full): zeroes the alignment-slack
    // columns of boxes that have any (crop windows that start mid-row), off the producer's path, so
    // that the producer never waits for a load to land.
    int acq_dst = -1;   // item whose destination tensor map this warp acquired last
    for (;; rs.next(), rl.next()) {
      const int stage = strm + K1_GROUPS * rs.i, slot = strm + K1_GROUPS * rl.i;
      mbar_wait_relaxed(landed + stage, rs.phase);
      const K1Tile& tl = slots[slot].tl;
      const int mode = tl.mode;
      if (mode == MODE_STAGED && tl.fix_hi > tl.fix_lo)
        k1_fix_columns(reinterpret_cast<float*>(smem + static_cast<size_t>(stage) * stage_bytes), tl, lane, slots[slot].ctx.it.src_dtype);
      if (mode == MODE_TSTORE) {
        // Plain copy tile: the box that just landed goes straight back out through the destination tensor
        // map — no consumer instructions, no LSU traffic.  The box is in SOURCE memory order: an axis whose
        // net direction is reversed (a flip) is stored slice by slice to the mirrored coordinate (planes for
        // axis 0, rows for axis 1; the item's map was encoded with that box shape: kind - ADELL_KIND_TSTORE).
        const adell_item& it = slots[slot].ctx.it;
        if (tl.item != acq_dst) { tmap_acquire(items[tl.item].dmap); acq_dst = tl.item; }
        const void* dmap = items[tl.item].dmap;
        const uint32_t base = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        const int split = it.kind - ADELL_KIND_TSTORE;
        const int n0 = min(tl.T[0], it.out_shape[0] - tl.o0[0]), n1 = min(tl.T[1], it.out_shape[1] - tl.o0[1]);
        const bool r0 = tl.msign[0] * it.grid_sign[0] < 0, r1 = tl.msign[1] * it.grid_sign[1] < 0;
        const uint32_t row = static_cast<uint32_t>(tl.box[2]) * 4u, plane = row * static_cast<uint32_t>(tl.box[1]);
        bool issued = false;
        if (split == 0) {
          if (lane == 0) { tma_store_3d(dmap, base, tl.o0[2], tl.o0[1], tl.o0[0]); issued = true; }
        } else if (split == 1) {
          if (lane < n0) {
            const int b0 = r0 ? n0 - 1 - lane : lane;
            tma_store_3d(dmap, base + b0 * plane, tl.o0[2], tl.o0[1], tl.o0[0] + lane);
            issued = true;
          }
        } else {
          for (int idx = lane; idx < n0 * n1; idx += 32) {
            const int d0 = idx / n1, d1 = idx - d0 * n1;
            const int b0 = r0 ? n0 - 1 - d0 : d0, b1 = r1 ? n1 - 1 - d1 : d1;
            tma_store_3d(dmap, base + b0 * plane + b1 * row, tl.o0[2], tl.o0[1] + d1, tl.o0[0] + d0);
            issued = true;
          }
        }
        if (issued) { tma_store_commit(); tma_store_wait_read(); }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full + stage);
      if (mode == MODE_DONE) break;
    }
    tma_store_wait_all();   // every store of this warp is complete before the kernel ends
    return;
  }

  // -------------------------------------------------------------------- consumer warps of stream `strm`
  K1_PROF_DECL
#ifdef K1_PROFILE
  const long long _tstart = clock64();
#endif
  for (;; rs.next(), rl.next()) {
    const int stage = strm + K1_GROUPS * rs.i, slot = strm + K1_GROUPS * rl.i;
    K1_PROF_T0
    mbar_wait(full + stage, rs.phase);
    K1_PROF_ADD(3)
    const K1Ctx& ctx = slots[slot].ctx;
    const K1Tile& tl = slots[slot].tl;
    const float* box = reinterpret_cast<const float*>(smem + static_cast<size_t>(stage) * stage_bytes);
    const adell_item& it = ctx.it;
    const int mode = tl.mode;
    if (mode == MODE_DONE) break;
    if (mode == MODE_COPY) {
      if (it.src_dtype == ADELL_F32) k1_tile_copy_box<ADELL_F32>(ctx, tl, box);
      else if (it.src_dtype == ADELL_I16) k1_tile_copy_box<ADELL_I16>(ctx, tl, box);
      else k1_tile_copy_box<ADELL_U8>(ctx, tl, box);
    } else if (mode == MODE_TSTORE) {
      // stored by the hand-over warp (TMA): the consumers only pass the stage on
    } else if (mode == MODE_ZERO) {
      k1_tile_zero(ctx, tl);
#ifdef K1_EXP_NO_CONSUME
    } else if (mode == MODE_STAGED) {   // TIMING EXPERIMENT (no voxels written): what the producer + TMA side alone sustains
#endif
    } else if (mode == MODE_STAGED) {
      // a pre offset must not leak into zero-filled (invalid) taps: the fast loops then scale it by the
      // sum of the valid tap weights (variant 3); where that is not available, the exact path
      const K1Fast& f = slots[slot].fast;
      const bool leak = ctx.pre_o != 0.0f && !tl.all_valid;
      const bool exact = (it.flags & (ADELL_F_STRICT | ADELL_F_CLIP)) != 0 || (leak && (tl.rmask != 0 || f.fl.x));
      if (exact) k1_tile_exact_dispatch<SmemTaps, false>(ctx, tl, box);
      else if (f.fl.x) {
        // (leak is false here, so the variant is 0, 1 or 2)
        const int rm = tl.rmask == 0 ? 0 : ((tl.rmask == 4 && it.padding == ADELL_PAD_REFLECTION) ? 1 : 2);
        if (f.fl.z && !f.fl.y && f.noise == nullptr && it.interp != ADELL_NEAREST && it.src_dtype == ADELL_F32)
          k1_tile_staged_noisy(tl, f, rm);
        else
          k1_tile_staged_cold(ctx, tl, f, box);
      } else {
        // block-uniform variant: 0 = no padding arithmetic, 1 = reflection on the thin axis only, 2 = general,
        // 3 = no padding arithmetic but invalid taps under a pre offset (valid-weight sum)
        const int rm = tl.rmask == 0 ? (leak ? 3 : 0) : ((tl.rmask == 4 && it.padding == ADELL_PAD_REFLECTION) ? 1 : 2);
        if (it.src_dtype == ADELL_F32) k1_staged_plain<ADELL_F32>(ctx, tl, f, box, rm);
        else if (it.src_dtype == ADELL_I16) k1_staged_plain<ADELL_I16>(ctx, tl, f, box, rm);
        else k1_staged_plain<ADELL_U8>(ctx, tl, f, box, rm);
      }
    } else if (it.flags & ADELL_F_IDENTITY) {
      k1_tile_exact_dispatch<GlobalTaps, true>(ctx, tl, nullptr);
    } else {
      k1_tile_exact_dispatch<GlobalTaps, false>(ctx, tl, nullptr);
    }
    __syncwarp();
#ifdef K1_PROFILE
    { long long _t1 = clock64(); _acc[4] += _t1 - _t0; if (threadIdx.x % K1_GTHREADS == 0) _acc[7] += 1;
      // per tile kind (first warp of each group): cycles and tiles of resampled [8,9] / consumer-copy [10,11] / TMA-store [12,13] tiles
      if (threadIdx.x % K1_GTHREADS == 0) { const int _k = mode == MODE_STAGED ? 8 : (mode == MODE_COPY ? 10 : (mode == MODE_TSTORE ? 12 : 14)); _acc[_k] += _t1 - _t0; _acc[_k + 1] += 1; } }
#endif
    if (lane == 0) mbar_arrive(empty + stage);
  }
#ifdef K1_PROFILE
  if (threadIdx.x % K1_GTHREADS == 0) {  // whole-kernel time of this consumer group: sum and max over groups
    const unsigned long long dt = static_cast<unsigned long long>(clock64() - _tstart);
    atomicAdd(&k1_prof[5], dt);
    atomicMax(&k1_prof[6], dt);
  }
#endif
  K1_PROF_FLUSH
}

// Shared-memory plan of the persistent CTA: ring stages (box + tile state + 3 mbarriers each) next to
// the fixed part (extra tile-state slots and item copies of the producers, barriers, tile prefix).
int k1_smem_fixed() { return K1_GROUPS * static_cast<int>(sizeof(K1Slot) + sizeof(K1Ctx)) + 64 + K1_TS_CACHE * 4; }
int k1_smem_per_stage(int stage_bytes) { return stage_bytes + static_cast<int>(sizeof(K1Slot)) + 24; }
// largest box that still leaves room for four ring stages (two per stream: one at work, one in flight)
int k1_pref_box_bytes() {
  const int b = ((K1_SMEM_BUDGET - k1_smem_fixed()) / 4 - static_cast<int>(sizeof(K1Slot)) - 24) & ~127;
  return b < K1_MAX_BOX_BYTES ? b : K1_MAX_BOX_BYTES;
}

int k1_validate(const adell_item& it) {
  for (int a = 0; a < 3; ++a) {
    if (it.out_shape[a] <= 0 || it.src_shape[a] <= 0 || it.grid_shape[a] <= 0) return ADELL_ERR_BAD_ARG;
    if (it.grid_sign[a] != 1 && it.grid_sign[a] != -1) return ADELL_ERR_BAD_ARG;
  }
  if (it.src_dtype > ADELL_U8) return ADELL_ERR_DTYPE;
  if (it.interp > ADELL_TRILINEAR || it.padding > ADELL_PAD_REFLECTION) return ADELL_ERR_BAD_ARG;
  if (it.src == nullptr || it.dst == nullptr) return ADELL_ERR_BAD_ARG;
  if (it.flags & ADELL_F_WIN_DEV) {
    if (it.win_dev == nullptr) return ADELL_ERR_BAD_ARG;
    for (int a = 0; a < 3; ++a)
      if (it.src_vlo[a] != 0 || it.src_vhi[a] < it.src_shape[a]) return ADELL_ERR_BAD_ARG;   // src_vhi = the parent's extents
    // zeros padding would have to stop at the window's edge: only the parent's edge is known to the staged box
    if (!(it.flags & ADELL_F_IDENTITY) && it.padding == ADELL_PAD_ZEROS) return ADELL_ERR_UNSUPPORTED;
  }
  return ADELL_OK;
}

// ---- host: per-item policy -------------------------------------------------------------------
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

// stands in for cuTensorMapEncodeTiled in adell_aug_plan: accepts every layout, encodes nothing
CUresult k1_encode_nothing(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill) { return CUDA_SUCCESS; }

// Identity item eligible for the box copy: unit step along axis 2, 16-byte aligned
// destination rows, nothing invalid, no noise, no strict-order post map.  (The source needs no
// alignment: it arrives through a TMA box whose origin is rounded down to 16 bytes.)
int k1_host_es(int dt) { return dt == ADELL_F32 ? 4 : (dt == ADELL_I16 ? 2 : 1); }

bool k1_vcopy_ok(const adell_item& it) {
  if (!(it.flags & ADELL_F_IDENTITY)) return false;   // (integer sources are converted on the way out of the box)
  if ((it.src_stride[2] != 1 && it.src_stride[2] != -1) || it.dst_stride[2] != 1 || it.noise != nullptr) return false;
  if ((it.flags & (ADELL_F_PHILOX | ADELL_F_STRICT)) || (it.out_shape[2] & 3) != 0) return false;
  for (int a = 0; a < 3; ++a) {
    if (it.out_vlo[a] > 0 || it.out_vhi[a] < it.out_shape[a]) return false;
    const int tlo = it.src_vlo[a] > 0 ? it.src_vlo[a] : 0;
    const int thi = it.src_vhi[a] < it.src_shape[a] ? it.src_vhi[a] : it.src_shape[a];
    const int ga = it.grid_off[a], gb = it.grid_off[a] + it.grid_sign[a] * (it.out_shape[a] - 1);
    if ((ga < gb ? ga : gb) < tlo || (ga > gb ? ga : gb) >= thi) return false;
  }
  if ((reinterpret_cast<uintptr_t>(it.dst) & 15u) || (it.dst_stride[0] & 3) || (it.dst_stride[1] & 3)) return false;
  return true;
}

// fp64 coordinate of output voxel (0,0,0) and its derivative (same expressions the device used to
// evaluate per tile): u_a(o) = (A_a . (g(o) - cg) + A_a3) * nrm_a*S_a/2 + (S_a-1)/2, g = off + sign*o.
void k1_item_map(adell_item& it) {
  double cc[3];
  for (int b = 0; b < 3; ++b)
    cc[b] = static_cast<double>(it.grid_off[b]) - static_cast<double>(static_cast<float>(it.grid_shape[b] - 1) * 0.5f);
  for (int a = 0; a < 3; ++a) {
    const double K = static_cast<double>(it.nrm[a]) * it.src_shape[a] * 0.5;
    const float* A = it.A + 4 * a;
    it.fp_U0[a] = (A[0] * cc[0] + A[1] * cc[1] + A[2] * cc[2] + A[3]) * K + (it.src_shape[a] - 1) * 0.5;
    for (int b = 0; b < 3; ++b) it.fp_D[3 * a + b] = static_cast<double>(A[b]) * K * it.grid_sign[b];
  }
}

// Tuning / debugging knobs from the environment, read once per process (getenv walks the whole
// environment: per item it cost more than the policy itself).
struct K1Tuning {
  bool no_shear, no_staged, no_tstore;
  int copy_t0, pref_box, tile_pref, chunk, tail, idle_stream, tstore_split;
  int64_t tile_cost;
  K1Tuning() {
    auto num = [](const char* name, int lo, int hi, int dflt) {
      const char* e = getenv(name);
      if (e == nullptr || e[0] == '\0') return dflt;   // unset or empty
      const int v = atoi(e);
      return (v >= lo && v <= hi) ? v : dflt;
    };
    const char* e = getenv("ADELL_K1_NO_SHEAR");
    no_shear = e != nullptr && e[0] == '1';
    e = getenv("ADELL_DISABLE_STAGED");                       // debugging aid: force the direct path
    no_staged = e != nullptr && e[0] == '1';
    e = getenv("ADELL_K1_NO_TSTORE");                         // measurement aid: copy tiles through the consumer warps
    no_tstore = e != nullptr && e[0] == '1';
    tstore_split = num("ADELL_K1_TSTORE_SPLIT", 0, 2, 1);    // finest split that still takes the TMA-store path
    copy_t0 = num("ADELL_K1_COPY_T0", 8, 32, K1_COPY_T0);
    if (copy_t0 != 8 && copy_t0 != 16 && copy_t0 != 32) copy_t0 = K1_COPY_T0;
    pref_box = num("ADELL_K1_PREF_BOX", 1024, K1_MAX_BOX_BYTES, -1);
    tile_pref = num("ADELL_K1_TILE", 0, 8, -1);               // 0 = 16x16x32, 1 = 16x32x16, 2 = 8x16x32, 3 = 16x16x16, 4.. the small shapes only
    chunk = num("ADELL_K1_CHUNK", 1, 64, 4);
    tail = num("ADELL_K1_TAIL", 0, 64, 3);
    idle_stream = num("ADELL_K1_IDLE_STREAM", 0, K1_GROUPS - 1, -1);   // measurement aid: that stream of every CTA takes no tiles
    tile_cost = num("ADELL_K1_TILE_COST", 0, 1 << 30, 3072);
  }
};
const K1Tuning& k1_tuning() {
  static const K1Tuning t;
  return t;
}

// Column-group shear of the tile grid.  A tile spans 16 or 32 voxels of axis 2 (the lanes); under a
// rotation that couples axis 2 into source axes 0/1 (a thin volume tilted about an in-plane axis)
// the footprint of such a column is slanted and its bounding box mostly empty.  Shifting the
// tile's (axis 0, axis 1) window per group of 8 voxels along axis 2 (8 fp32 = one 32-byte sector:
// row stores stay sector-complete) straightens the footprint: the shifts s(G) solve
//   [D00 D01; D10 D11] s = (D02, D12) * 8G   (rounded to integers, made non-negative).
// Any integer shifts give a valid partition of the output (each column group is tiled by its own
// shifted grid).  Returns true when a non-trivial shear was written to it.shear.
bool k1_item_shear(adell_item& it) {
  memset(it.shear, 0, sizeof(it.shear));
  if (k1_tuning().no_shear) return false;
  const int nG = (it.out_shape[2] + 7) >> 3;
  if (nG < 2 || nG > 16) return false;
  const double* D = it.fp_D;
  const double det = D[0] * D[4] - D[1] * D[3];
  if (!(fabs(det) > 0.25)) return false;
  const double a0 = (D[4] * D[2] - D[1] * D[5]) / det * 8.0;
  const double a1 = (-D[3] * D[2] + D[0] * D[5]) / det * 8.0;
  long s[2][16], mn[2] = {0, 0}, mx[2] = {0, 0};
  for (int G = 0; G < nG; ++G) {
    s[0][G] = lrint(a0 * G);
    s[1][G] = lrint(a1 * G);
    for (int b = 0; b < 2; ++b) {
      if (s[b][G] < mn[b]) mn[b] = s[b][G];
      if (s[b][G] > mx[b]) mx[b] = s[b][G];
    }
  }
  if (mx[0] - mn[0] > 127 || mx[1] - mn[1] > 127) return false;
  if (mx[0] == mn[0] && mx[1] == mn[1]) return false;
  for (int G = 0; G < nG; ++G)
    for (int b = 0; b < 2; ++b) it.shear[b][G] = static_cast<int8_t>(s[b][G] - mn[b]);
  return true;
}

int k1_max_shear(const adell_item& it, int b) {
  int m = 0;
  for (int G = 0; G < 16; ++G) if (it.shear[b][G] > m) m = it.shear[b][G];
  return m;
}

// Extent along source axis a of the group terms e_a(G) = D_a2*8g - D_a0*s0(G) - D_a1*s1(G) over the
// column groups g of one tile (g = G - first group of the tile), maximised over the tile rows.
double k1_group_range(const adell_item& it, int a, int T2) {
  const int nG = (it.out_shape[2] + 7) >> 3, ng = T2 >> 3;
  double range = 0.0;
  for (int G0 = 0; G0 < nG; G0 += (ng > 0 ? ng : 1)) {
    double lo = 0.0, hi = 0.0;
    for (int g = 0; g < ng && G0 + g < nG; ++g) {
      const int G = (G0 + g) & 15;
      const double ev = it.fp_D[3 * a + 2] * 8.0 * g - it.fp_D[3 * a + 0] * it.shear[0][G] - it.fp_D[3 * a + 1] * it.shear[1][G];
      if (g == 0 || ev < lo) lo = ev;
      if (g == 0 || ev > hi) hi = ev;
    }
    if (hi - lo > range) range = hi - lo;
  }
  return range;
}

// Extent of the tile along output axis b that the footprint has to cover.
int k1_tile_extent(const adell_item& it, const int* T, int b, bool sheared) {
  if (b == 2) return it.out_shape[2] < 8 ? it.out_shape[2] : 8;  // one column group (staged tiles span 16 or 32)
  if (sheared) return T[b];  // a shifted window can hold a full tile even where out_shape < T
  return it.out_shape[b] < T[b] ? it.out_shape[b] : T[b];
}

// Footprint box of one T-shaped output tile; returns its bytes (0: not stageable).  Needs
// tmap_sign / tmap_off (k1_tmap_layout).
int64_t k1_box_for_tile(const adell_item& it, const int* T, int* box, bool sheared) {
  int64_t cells = 1;
  const int es = k1_host_es(it.src_dtype), q = 16 / es;   // element size; elements per 16 bytes
  for (int a = 0; a < 3; ++a) {
    double span = k1_group_range(it, a, T[2]);
    for (int b = 0; b < 3; ++b) span += fabs(it.fp_D[3 * a + b]) * (k1_tile_extent(it, T, b, sheared) - 1);
    if (!(span < 4096.0)) return 0;
    box[a] = static_cast<int>(floor(span + 2 * K1_EPS)) + 3;
    bool whole = false;
    if (it.padding != ADELL_PAD_ZEROS) {
      // border / reflection fold the coordinates back into [0, S-1) (the fast path clamps just below
      // S-1, so the hi tap never leaves the volume), and a thin axis is simply staged whole so that
      // multiply-reflected tiles stay on this path
      if (it.src_shape[a] <= 64 || box[a] >= it.src_shape[a]) { box[a] = it.src_shape[a] > 1 ? it.src_shape[a] : 2; whole = true; }
      if (it.tmap_sign[a] < 0) box[a] += 1;   // reversed axis: the cell below the staged interval (see k1_tile_setup)
    }
    if (a == 2) {
      // inner extent: a 16-byte multiple, plus the cells lost to aligning the box origin down to 16 bytes
      // (known exactly when the axis is staged whole)
      int lead = q - 1;
      if (whole && !(it.flags & ADELL_F_WIN_DEV)) {   // (a device-side window can start anywhere in its 16 bytes)
        const int mo = it.tmap_sign[2] > 0 ? it.tmap_off[2] : -(it.src_shape[2] - 1) + it.tmap_off[2];
        lead = mo & (q - 1);
      }
      box[a] = (box[a] + lead + q - 1) & ~(q - 1);
    }
    if (box[a] > 256) return 0;
    cells *= box[a];
  }
  return cells * es;
}

// Memory-order layout of the item's valid source box: fills tmap_sign / tmap_off / fp_fix and the
// tensor-map geometry.  Returns false when the layout cannot be expressed as a tensor map.
struct K1Layout {
  cuuint64_t gdim[3], gstride[2];
  uintptr_t base;
  int es = 4;   // element size in bytes: 4 fp32, 2 int16, 1 uint8
};
bool k1_tmap_layout(adell_item& it, K1Layout& L) {
  int64_t base_off = 0;
  int64_t astride[3];
  for (int a = 0; a < 3; ++a) {
    const int tlo = it.src_vlo[a] > 0 ? it.src_vlo[a] : 0;
    const int thi = it.src_vhi[a] < it.src_shape[a] ? it.src_vhi[a] : it.src_shape[a];
    if (thi <= tlo) return false;
    const int sign = it.src_stride[a] >= 0 ? 1 : -1;
    astride[a] = it.src_stride[a] * sign;
    if (astride[a] == 0) return false;
    it.tmap_sign[a] = sign;
    it.tmap_off[a] = sign > 0 ? -tlo : thi - 1;           // m = sign*t + off, m = 0 at the lowest address
    base_off += static_cast<int64_t>(sign > 0 ? tlo : thi - 1) * it.src_stride[a];
    // tensor-map dims are innermost first; a device-side window (ADELL_F_WIN_DEV) maps the whole parent: the
    // kernel adds the window's start to the box coordinates
    L.gdim[2 - a] = static_cast<cuuint64_t>((it.flags & ADELL_F_WIN_DEV) ? it.src_vhi[a] : thi - tlo);
  }
  const int es = k1_host_es(it.src_dtype);
  L.es = es;
  uintptr_t base = reinterpret_cast<uintptr_t>(it.src) + static_cast<uintptr_t>(base_off * es);
  L.gstride[0] = static_cast<cuuint64_t>(astride[1]) * es;
  L.gstride[1] = static_cast<cuuint64_t>(astride[0]) * es;
  if ((base & static_cast<uintptr_t>(es - 1)) || (L.gstride[0] & 15u) || (L.gstride[1] & 15u)) return false;
  if (L.gstride[0] < L.gdim[0] * es || L.gstride[1] < L.gstride[0]) return false;  // rows must not overlap
  // a crop window may start anywhere in a row: align the tensor-map base down to 16 bytes; the
  // elements this prepends to every row (up to 3 fp32 / 7 int16 / 15 uint8) are zeroed in shared memory
  // after each load (fp_fix)
  const int slack = static_cast<int>(base & 15u) / es;
  base -= static_cast<uintptr_t>(slack) * es;
  L.gdim[0] += static_cast<cuuint64_t>(slack);
  it.tmap_off[2] += slack;
  it.fp_fix = slack;
  L.base = base;
  return true;
}

// Encodes one 3-D fp32 tensor map (innermost dimension first) into out_map.  Returns 1, 0 (the driver
// refused the layout) or -1 (no driver).
// A device-resident cache hands the same volumes back every epoch and the outputs are written into the
// same batch tensors, and a map only depends on the memory layout (flips are signs in the box index):
// the last encodings are remembered per host thread instead of asking the driver again (~0.4 us each).
int k1_encode_map(uint8_t* out_map, const K1Layout& L, const cuuint32_t* bdim, EncodeTiledFn enc) {
  const cuuint32_t estr[3] = {1, 1, 1};
  if (enc == nullptr) return -1;
  struct Entry { uintptr_t base; cuuint64_t gdim[3], gstride[2]; cuuint32_t box[3]; int es; bool valid; uint8_t map[128]; };
  constexpr int kEntries = 2048;
  static thread_local Entry* cache = nullptr;
  if (cache == nullptr) cache = static_cast<Entry*>(calloc(kEntries, sizeof(Entry)));
  Entry* e = nullptr;
  if (cache != nullptr && enc != &k1_encode_nothing) {
    uint64_t h = (L.base >> 4) * 0x9E3779B97F4A7C15ull ^ (L.gdim[0] * 31 + L.gdim[1] * 131 + L.gdim[2] * 1031 + bdim[0] * 7 + bdim[1] * 11 + bdim[2] * 13);
    e = cache + (h >> 40) % kEntries;
    if (e->valid && e->base == L.base && e->gdim[0] == L.gdim[0] && e->gdim[1] == L.gdim[1] && e->gdim[2] == L.gdim[2] &&
        e->gstride[0] == L.gstride[0] && e->gstride[1] == L.gstride[1] && e->box[0] == bdim[0] && e->box[1] == bdim[1] &&
        e->box[2] == bdim[2] && e->es == L.es) {
      memcpy(out_map, e->map, 128);
      return 1;
    }
  }
  const CUtensorMapDataType dt = L.es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (L.es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out_map), dt, 3,
                   reinterpret_cast<void*>(L.base), L.gdim, L.gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, K1_L2_PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return 0;
  if (e != nullptr) {
    e->base = L.base;
    for (int i = 0; i < 3; ++i) { e->gdim[i] = L.gdim[i]; e->box[i] = bdim[i]; }
    e->gstride[0] = L.gstride[0]; e->gstride[1] = L.gstride[1];
    e->es = L.es;
    memcpy(e->map, out_map, 128);
    e->valid = true;
  }
  return 1;
}

// Encodes the source tensor map for a staged box of the given extents (axes 0,1,2).
int k1_encode_tmap(adell_item& it, const K1Layout& L, const int* box, EncodeTiledFn enc) {
  it.flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
  for (int a = 0; a < 3; ++a) it.tmap_box[a] = box[a];
  const cuuint32_t bdim[3] = {static_cast<cuuint32_t>(box[2]), static_cast<cuuint32_t>(box[1]), static_cast<cuuint32_t>(box[0])};
  const int r = k1_encode_map(it.tmap, L, bdim, enc);
  if (r <= 0) return r;
  it.flags |= ADELL_F_TMAP;
  return 1;
}

// Plain copy item (nothing but integer flips / crops, no intensity map, contiguous axis not reversed,
// 16-byte aligned rows on both sides): its tiles can leave through a destination tensor map (TMA
// store) without touching the consumer warps.  Fills it.dmap and returns the split (0: whole boxes,
// 1: plane by plane — axis 0 reversed, 2: row by row — axis 1 reversed), -1 when not eligible, -2 without
// a driver.  Needs tmap_sign (k1_tmap_layout) and the encoded source box.
int k1_encode_dst(adell_item& it, const int* T, const int* box, EncodeTiledFn enc) {
  if (k1_tuning().no_tstore) return -1;
  if (it.src_dtype != ADELL_F32) return -1;   // integer sources are converted by the consumer warps
  if (it.flags & ADELL_F_WIN_DEV) return -1;  // window alignment unknown on the host
  if (it.flags & (ADELL_F_CLIP | ADELL_F_PRE_DEV)) return -1;
  if (it.pre_scale != 1.0f || it.pre_offset != 0.0f || it.post_scale != 1.0f || it.post_offset != 0.0f) return -1;
  if (it.tmap_sign[2] * it.grid_sign[2] < 0) return -1;          // a flip along the contiguous axis reverses elements
  if (box[2] != T[2] || box[1] != T[1] || box[0] != T[0]) return -1;  // unaligned window: the box carries slack columns
  if (it.fp_fix != 0) return -1;
  K1Layout L;
  L.base = reinterpret_cast<uintptr_t>(it.dst);
  if (L.base & 15u) return -1;
  if (it.dst_stride[2] != 1 || it.dst_stride[1] < it.out_shape[2] || it.dst_stride[0] < it.dst_stride[1] * it.out_shape[1]) return -1;
  L.gdim[0] = static_cast<cuuint64_t>(it.out_shape[2]);
  L.gdim[1] = static_cast<cuuint64_t>(it.out_shape[1]);
  L.gdim[2] = static_cast<cuuint64_t>(it.out_shape[0]);
  L.gstride[0] = static_cast<cuuint64_t>(it.dst_stride[1]) * 4;
  L.gstride[1] = static_cast<cuuint64_t>(it.dst_stride[0]) * 4;
  if ((L.gstride[0] & 15u) || (L.gstride[1] & 15u)) return -1;
  const bool r0 = it.tmap_sign[0] * it.grid_sign[0] < 0, r1 = it.tmap_sign[1] * it.grid_sign[1] < 0;
  const int split = r1 ? 2 : (r0 ? 1 : 0);
  // row-by-row stores (128 B each, 256 per tile) are slower than the consumer copy: measured
  if (split > k1_tuning().tstore_split) return -1;
  const cuuint32_t bdim[3] = {static_cast<cuuint32_t>(T[2]), static_cast<cuuint32_t>(split == 2 ? 1 : T[1]),
                              static_cast<cuuint32_t>(split == 0 ? T[0] : 1)};
  const int r = k1_encode_map(it.dmap, L, bdim, enc);
  if (r < 0) return -2;
  return r == 0 ? -1 : split;
}

// Identity item: tensor map for the 32x16x32 box copy.  Returns the box bytes (0 = not eligible).
int k1_encode_copy(adell_item& it, EncodeTiledFn enc) {
  if (!k1_vcopy_ok(it)) return 0;
  int T[3] = {k1_tuning().copy_t0, 16, 32};
  int box[3] = {T[0], T[1], T[2]};
  K1Layout L;
  if (!k1_tmap_layout(it, L)) return 0;
  int r = k1_encode_tmap(it, L, box, enc);
  if (r <= 0) return r;
  // tensor coordinate of the first tile's lowest element along axis 2: if it is not a multiple of
  // four the box origin is rounded down per tile and the box needs four more columns
  const int n2 = it.out_shape[2] < T[2] ? it.out_shape[2] : T[2];
  const int g0 = it.grid_off[2], g1 = it.grid_off[2] + it.grid_sign[2] * (n2 - 1);
  const int lo = it.tmap_sign[2] > 0 ? (g0 < g1 ? g0 : g1) + it.tmap_off[2] : -(g0 > g1 ? g0 : g1) + it.tmap_off[2];
  const int q = 16 / L.es;
  if ((lo & (q - 1)) || (it.flags & ADELL_F_WIN_DEV)) {   // (a device-side window: its alignment is not known here)
    box[2] += q;
    r = k1_encode_tmap(it, L, box, enc);
    if (r <= 0) return r;
  }
  for (int a = 0; a < 3; ++a) it.tile_dim[a] = static_cast<uint8_t>(T[a]);
  const int split = k1_encode_dst(it, T, box, enc);
  if (split == -2) return -1;
  it.kind = static_cast<uint8_t>(split < 0 ? ADELL_KIND_VCOPY : ADELL_KIND_TSTORE + split);
  return box[0] * box[1] * box[2] * L.es;
}

// Decides staged-path eligibility for one item and, when eligible, picks its tile shape, encodes
// its tensor map over the valid source box in memory order and fills the derived fields.
// Returns the box bytes (0 = not staged, -1 = no driver).
int k1_encode_item(adell_item& it, EncodeTiledFn enc, int tile_pref) {
  it.flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
  if (it.flags & ADELL_F_IDENTITY) return 0;
  if (it.src_stride[2] != 1 && it.src_stride[2] != -1) return 0;
  k1_item_map(it);
  const bool sheared = k1_item_shear(it);
  const int smx[2] = {k1_max_shear(it, 0), k1_max_shear(it, 1)};
  K1Layout L;
  if (!k1_tmap_layout(it, L)) { memset(it.shear, 0, sizeof(it.shear)); return 0; }
  // tile shapes: among those whose box leaves room for three ring stages (two consumer groups at
  // work + one load in flight) take the cheapest one: voxels of all its tiles, padding included (idle
  // lanes cost as much as busy ones), plus a fixed cost per tile (producer set-up, consumer prologue:
  // worth ~3000 voxels, measured); ties go to the shape listed first.  Only when none fits, the same
  // choice among the boxes that fit twice.
  // Shapes 4.. are small tiles for footprints that the regular ones cannot stage within the preferred box (strong
  // zooms: the workhorse's scale members draw factors around 2): ONE item with a ~100 KB box would leave the whole
  // launch with a single ring stage per stream (no load in flight while a tile is consumed).
  static const int kShapes[9][3] = {{16, 16, 32}, {16, 32, 16}, {8, 16, 32}, {16, 16, 16}, {8, 16, 16}, {8, 8, 32}, {8, 8, 16}, {4, 8, 16}, {4, 4, 16}};
  const int pref_bytes = k1_tuning().pref_box > 0 ? k1_tuning().pref_box : k1_pref_box_bytes();
  int box[3] = {0, 0, 0}, T[3] = {16, 16, 16};
  int64_t bytes = 0;
  const int64_t tile_cost = k1_tuning().tile_cost;
  // pass 0: the regular shapes within the preferred box; pass 1: the small ones within it; pass 2: any shape up to
  // the largest box the kernel can hold twice
  for (int pass = 0; pass < 3 && bytes == 0; ++pass) {
    const int limit = pass < 2 ? pref_bytes : K1_MAX_BOX_BYTES;
    int64_t best_cover = -1;
    for (int s = (pass == 1 ? 4 : 0); s < (pass == 0 ? 4 : 9); ++s) {
      if (tile_pref >= 0 && s != tile_pref) continue;
      if (kShapes[s][2] == 32 && it.out_shape[2] <= 16) continue;
      if (s == 1 && it.out_shape[1] <= 16) continue;
      int bx[3];
      const int64_t b = k1_box_for_tile(it, kShapes[s], bx, sheared);
      if (b == 0 || b > limit) continue;
      int64_t cover = 1, tiles = 1;
      for (int a = 0; a < 3; ++a) {
        const int64_t nt = (it.out_shape[a] + (a < 2 ? smx[a] : 0) + kShapes[s][a] - 1) / kShapes[s][a];
        tiles *= nt;
        cover *= nt * kShapes[s][a];
      }
      cover += tiles * tile_cost;
      if (pass == 2) cover = b;   // nothing fits the preferred box: the SMALLEST box (the launch keeps the most stages)
      if (best_cover >= 0 && cover >= best_cover) continue;
      best_cover = cover;
      bytes = b;
      for (int a = 0; a < 3; ++a) { box[a] = bx[a]; T[a] = kShapes[s][a]; }
    }
  }
  if (bytes == 0) { memset(it.shear, 0, sizeof(it.shear)); return 0; }
  const int r = k1_encode_tmap(it, L, box, enc);
  if (r <= 0) { memset(it.shear, 0, sizeof(it.shear)); return r; }
  for (int a = 0; a < 3; ++a) {
    it.tile_dim[a] = static_cast<uint8_t>(T[a]);
    double smin = 0.0, smax = 0.0;  // footprint of one 8-voxel column group of a tile (di, dj, kk in [0,8))
    for (int b = 0; b < 3; ++b) {
      const double span = it.fp_D[3 * a + b] * (k1_tile_extent(it, T, b, sheared) - 1);
      if (span < 0) smin += span; else smax += span;
    }
    // rounded outwards so the fp32 copies never under-estimate the footprint
    it.fp_smin[a] = nextafterf(static_cast<float>(smin), -INFINITY);
    it.fp_smax[a] = nextafterf(static_cast<float>(smax), INFINITY);
  }
  it.kind = ADELL_KIND_STAGED;
  return static_cast<int>(bytes);
}

}  // namespace

namespace {
int k1_prepare_impl(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info, bool plan_only);
}  // namespace

// adell_item is declared 64-byte aligned and the host code is compiled against that: a misaligned host buffer is refused
// (it used to be silent undefined behaviour that only vector moves of a wider ISA would have turned into a fault).
static inline bool k1_items_misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 63u) != 0; }

extern "C" int adell_aug_prepare(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info) {
  if (n_items > 0 && k1_items_misaligned(items_host)) return ADELL_ERR_ALIGN;
  return k1_prepare_impl(items_host, n_items, tile_start_host, info, false);
}

extern "C" int adell_aug_plan(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info) {
  if (n_items > 0 && k1_items_misaligned(items_host)) return ADELL_ERR_ALIGN;
  const int st = k1_prepare_impl(items_host, n_items, tile_start_host, info, true);
  if (st == ADELL_OK) {
    for (int i = 0; i < n_items; ++i) items_host[i].flags &= static_cast<uint8_t>(~ADELL_F_TMAP);  // nothing was encoded
    info->n_staged = -1 - info->n_staged;   // marks the plan as not launchable (adell_aug_gather refuses it)
  }
  return st;
}

namespace {
int k1_prepare_impl(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info, bool plan_only) {
  if (items_host == nullptr || tile_start_host == nullptr || info == nullptr || n_items < 0) return ADELL_ERR_BAD_ARG;
  int64_t acc = 0;
  int smem = 0, staged = 0;
  EncodeTiledFn enc = plan_only ? &k1_encode_nothing : nullptr;
  bool enc_tried = plan_only;
  const bool no_staged = k1_tuning().no_staged;
  const int tile_pref = k1_tuning().tile_pref;
  for (int i = 0; i < n_items; ++i) {
    adell_item& it = items_host[i];
    int st = k1_validate(it);
    if (st != ADELL_OK) return st;
    if (!enc_tried && !no_staged) { enc = k1_get_encode(); enc_tried = true; }
    it.kind = ADELL_KIND_GENERIC;
    it.tile_dim[0] = it.tile_dim[1] = it.tile_dim[2] = K1_T;
    it.fp_fix = 0;
    memset(it.shear, 0, sizeof(it.shear));
    int bytes = 0;
    it.flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
    if (!no_staged) {
      bytes = (it.flags & ADELL_F_IDENTITY) ? k1_encode_copy(it, enc) : k1_encode_item(it, enc, tile_pref);
      if (bytes == 0) {  // not staged after all: generic 16^3 tiles
        memset(it.shear, 0, sizeof(it.shear));
        it.flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
        it.kind = ADELL_KIND_GENERIC;
        it.tile_dim[0] = it.tile_dim[1] = it.tile_dim[2] = K1_T;
        it.fp_fix = 0;
      }
    }
    if (bytes < 0) return ADELL_ERR_NO_DRIVER;
    if (bytes > 0) { ++staged; if (bytes > smem) smem = bytes; }
    int64_t n = 1;
    for (int a = 0; a < 3; ++a) {
      const int ext = it.out_shape[a] + ((a < 2 && it.kind == ADELL_KIND_STAGED) ? k1_max_shear(it, a) : 0);
      it.n_tiles[a] = (ext + it.tile_dim[a] - 1) / it.tile_dim[a];
      n *= it.n_tiles[a];
    }
    tile_start_host[i] = static_cast<int32_t>(n);  // tile count for now; the prefix follows the reordering
    acc += n;
    if (acc > 0x7fffffffLL) return ADELL_ERR_BAD_ARG;
  }
  // Identity (box-copy) items go behind the others (stable): tiles [0, first_copy) then feed the
  // "resampled" queue and the rest the "copy" queue, so that every SM can keep a compute-bound and a
  // memory-bound tile in flight at once.  Items are independent: their order does not matter.
  int n_copy = 0;
  for (int i = 0; i < n_items; ++i) n_copy += items_host[i].kind >= ADELL_KIND_VCOPY;
  if (n_copy > 0 && n_copy < n_items) {
    adell_item* tmp = static_cast<adell_item*>(malloc(sizeof(adell_item) * static_cast<size_t>(n_items)));
    int32_t* cnt = static_cast<int32_t*>(malloc(sizeof(int32_t) * static_cast<size_t>(n_items)));
    if (tmp == nullptr || cnt == nullptr) { free(tmp); free(cnt); return ADELL_ERR_BAD_ARG; }
    int w = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int i = 0; i < n_items; ++i)
        if ((items_host[i].kind >= ADELL_KIND_VCOPY) == (pass == 1)) { memcpy(tmp + w, items_host + i, sizeof(adell_item)); cnt[w++] = tile_start_host[i]; }
    memcpy(items_host, tmp, sizeof(adell_item) * static_cast<size_t>(n_items));
    memcpy(tile_start_host, cnt, sizeof(int32_t) * static_cast<size_t>(n_items));
    free(tmp); free(cnt);
  }
  int64_t run = 0, first_copy = -1;
  for (int i = 0; i < n_items; ++i) {
    if (first_copy < 0 && items_host[i].kind >= ADELL_KIND_VCOPY) first_copy = run;
    const int32_t n = tile_start_host[i];
    tile_start_host[i] = static_cast<int32_t>(run);
    run += n;
  }
  tile_start_host[n_items] = static_cast<int32_t>(acc);
  for (int q = 1; q <= 4; ++q) tile_start_host[n_items + q] = 0;  // chunk queues of the launch (see adell_aug_gather)
  info->first_copy_tile = first_copy < 0 ? acc : first_copy;
  info->total_tiles = acc;
  info->smem_bytes = smem;
  info->n_staged = staged;
  return ADELL_OK;
}
}  // namespace

extern "C" int adell_aug_prepare_steps(void* buf_host, int n_steps, const int32_t* n_items, const int64_t* item_off,
                                       const int64_t* tile_off, adell_launch_info* infos) {
  if (buf_host == nullptr || n_items == nullptr || item_off == nullptr || tile_off == nullptr || infos == nullptr || n_steps < 0)
    return ADELL_ERR_BAD_ARG;
  uint8_t* base = static_cast<uint8_t*>(buf_host);
  for (int k = 0; k < n_steps; ++k) {
    const int st = adell_aug_prepare(reinterpret_cast<adell_item*>(base + item_off[k]), n_items[k],
                                     reinterpret_cast<int32_t*>(base + tile_off[k]), infos + k);
    if (st != ADELL_OK) return st;
  }
  return ADELL_OK;
}

extern "C" int adell_aug_gather(const adell_item* items_dev, const int32_t* tile_start_dev, int n_items,
                                const adell_launch_info* info, void* stream) {
  if (n_items == 0) return ADELL_OK;
  if (info == nullptr) return ADELL_ERR_BAD_ARG;
  if (info->total_tiles == 0) return ADELL_OK;
  if (items_dev == nullptr || tile_start_dev == nullptr || n_items < 0 || info->total_tiles < 0 ||
      info->total_tiles > 0x7fffffffLL || info->smem_bytes < 0 || info->smem_bytes > K1_MAX_BOX_BYTES || info->n_staged < 0)
    return ADELL_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(items_dev) & 63u) != 0) return ADELL_ERR_ALIGN;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // ring of staged boxes: as many stages as fit next to the per-stage tile state
  const int stage_bytes = (info->smem_bytes + 127) & ~127;
  const int per_stage = k1_smem_per_stage(stage_bytes);
  const int fixed = k1_smem_fixed();
  int n_stages = (K1_SMEM_BUDGET - fixed) / per_stage;
  if (n_stages > K1_MAX_STAGES) n_stages = K1_MAX_STAGES;
  n_stages -= n_stages % K1_GROUPS;  // the same number of stages for every stream
  if (n_stages < K1_GROUPS) return ADELL_ERR_BAD_ARG;
  const int smem = n_stages * per_stage + fixed;
  e = cudaFuncSetAttribute(k1_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BUDGET);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // consecutive tiles a producer takes from the queue at a time (they share the item and neighbouring
  // source boxes): small enough that the tail of the launch stays balanced
  const int chunk = k1_tuning().chunk | ((k1_tuning().idle_stream + 1) << 16);
  // the last ~3 tiles per stream of each queue are handed out one by one
  const int64_t streams = static_cast<int64_t>(sms) * K1_GROUPS;
  const int64_t tail = k1_tuning().tail * streams;
  const int64_t first_copy = info->first_copy_tile < 0 || info->first_copy_tile > info->total_tiles ? info->total_tiles : info->first_copy_tile;
  const int64_t tiles_q[2] = {first_copy, info->total_tiles - first_copy};
  int64_t n_big[2], n_units = 0;
  for (int q = 0; q < 2; ++q) {
    n_big[q] = tiles_q[q] > tail ? (tiles_q[q] - tail) / (chunk & 0xffff) : 0;
    n_units += n_big[q] + (tiles_q[q] - n_big[q] * (chunk & 0xffff));
  }
  const int64_t n_ctas = (n_units + K1_GROUPS - 1) / K1_GROUPS;
  const int64_t grid = n_ctas < sms ? n_ctas : sms;
  // the words after the tile prefix are the launch's chunk queues {next resampled unit, next copy unit,
  // drained producers}: zero on upload (adell_aug_prepare), re-armed by the kernel itself when it finishes
  unsigned int* sched = reinterpret_cast<unsigned int*>(const_cast<int32_t*>(tile_start_dev) + n_items + 1);
  k1_gather<<<static_cast<unsigned>(grid), K1_THREADS, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(
      items_dev, tile_start_dev, n_items, static_cast<int>(info->total_tiles), n_stages, stage_bytes, chunk,
      static_cast<int>(n_big[0]), static_cast<int>(n_big[1]), static_cast<int>(first_copy), sched);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_aug_gather_launches(void) { return 1; }

#ifdef K1_PROFILE
// debug builds only: cumulative cycle counters {producer wait-empty, issue, prepare, consumer wait-full, compute}
extern "C" int adell_debug_prof(unsigned long long* out16, int reset) {
  cudaMemcpyFromSymbol(out16, k1_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(k1_prof, z, sizeof(z)); }
  return 0;
}
extern "C" int adell_debug_prof2(unsigned long long* out16, int reset) {
  cudaMemcpyFromSymbol(out16, k1_prof2, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(k1_prof2, z, sizeof(z)); }
  return 0;
}
#endif
