// Host-side chain composer: turns per-volume chain descriptors (adell_chain) into canonical
// adell_item arrays and prepares the launches — the native form of what plan.BatchPlan does in numpy
// for the single-pass chains of the reference's three pipelines:
//
//   parent -> SpatialCrop (RandSpatialCropd / RandCropByPosNegLabeld window)            crop0
//          -> flips BEFORE the resample (get_augmentations_class: OneOf(RandFlipd) first) flip0
//          -> RandAffined (one resample, or none)                                        A
//          -> flips AFTER the resample (get_augmentations_unet: RandFlipd per axis)      flip1
//          -> CenterSpatialCropd                                                         crop1
//          -> intensity map
//   (/root/reference/adell_mri/transform_factory/augmentations.py:98-176,255-320,427-515)
//
// No device work here: integer index algebra per volume, then adell_aug_prepare per step.  Exists so
// that the per-step host cost of a training loop is a few microseconds instead of ~0.1 ms of numpy
// (one Python rank per GPU shares the host's cores with 7 others: host composition was what held
// 8-GPU scaling at 0.72 in round 1).
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

// Per-axis integer index map (plan.IntMap): index t of the current space [0,size) reads underlying
// index off + sign*t and is valid iff vlo <= t < vhi (else literal 0).
struct IntMap3 {
  int64_t size[3], off[3], sign[3], vlo[3], vhi[3];
  void init(const int32_t* s) {
    for (int a = 0; a < 3; ++a) { size[a] = s[a]; off[a] = 0; sign[a] = 1; vlo[a] = 0; vhi[a] = s[a]; }
  }
  void init64(const int64_t* s) {
    for (int a = 0; a < 3; ++a) { size[a] = s[a]; off[a] = 0; sign[a] = 1; vlo[a] = 0; vhi[a] = s[a]; }
  }
  static int64_t clip(int64_t v, int64_t lo, int64_t hi) { return v < lo ? lo : (v > hi ? hi : v); }
  // SpatialCrop: keep [start, start + n) — start and n already clipped to the current extent
  void crop(const int64_t* start, const int64_t* n) {
    for (int a = 0; a < 3; ++a) {
      off[a] += sign[a] * start[a];
      vlo[a] = clip(vlo[a] - start[a], 0, n[a]);
      vhi[a] = clip(vhi[a] - start[a], 0, n[a]);
      size[a] = n[a];
    }
  }
  void flip(unsigned mask) {
    for (int a = 0; a < 3; ++a) {
      if (!(mask >> a & 1u)) continue;
      off[a] += sign[a] * (size[a] - 1);
      const int64_t nlo = size[a] - vhi[a], nhi = size[a] - vlo[a];
      sign[a] = -sign[a];
      vlo[a] = nlo; vhi[a] = nhi;
    }
  }
};

int elsize(int dtype) { return dtype == ADELL_F32 ? 4 : (dtype == ADELL_I16 ? 2 : 1); }

int compose_one(const adell_chain& c, adell_item& it) {
  if (c.src == nullptr || c.dst == nullptr || c.src_dtype > ADELL_U8) return ADELL_ERR_BAD_ARG;
  memset(&it, 0, sizeof(it));
  IntMap3 pre, post;
  pre.init(c.src_shape);
  const bool affine = (c.flags & ADELL_CHAIN_AFFINE) != 0;
  int64_t st[3], n[3];
  bool has_crop0 = false;
  for (int a = 0; a < 3; ++a) has_crop0 |= c.crop0_size[a] > 0;
  if (has_crop0) {   // BatchPlan.crop: start clipped to [0, cur], size to cur - start
    for (int a = 0; a < 3; ++a) {
      st[a] = IntMap3::clip(c.crop0_start[a], 0, pre.size[a]);
      const int64_t want = c.crop0_size[a] > 0 ? c.crop0_size[a] : pre.size[a];
      n[a] = want < pre.size[a] - st[a] ? want : pre.size[a] - st[a];
    }
    pre.crop(st, n);
  }
  pre.flip(c.flip0);
  IntMap3* cur = &pre;
  if (affine) { post.init64(pre.size); cur = &post; }
  cur->flip(c.flip1);
  bool has_crop1 = false;
  for (int a = 0; a < 3; ++a) has_crop1 |= c.crop1_size[a] > 0;
  if (has_crop1) {   // CenterSpatialCrop: start = max(size/2 - roi/2, 0), roi <= 0 keeps the axis
    for (int a = 0; a < 3; ++a) {
      int64_t roi = c.crop1_size[a] <= 0 ? cur->size[a] : (c.crop1_size[a] < cur->size[a] ? c.crop1_size[a] : cur->size[a]);
      int64_t s0 = cur->size[a] / 2 - roi / 2;
      if (s0 < 0) s0 = 0;
      st[a] = IntMap3::clip(s0, 0, cur->size[a]);
      n[a] = roi < cur->size[a] - st[a] ? roi : cur->size[a] - st[a];
    }
    cur->crop(st, n);
  }
  // plan.BatchPlan._fill_items
  int64_t elem_off = 0;
  for (int a = 0; a < 3; ++a) elem_off += pre.off[a] * c.src_stride[a];
  it.src = static_cast<const char*>(c.src) + elem_off * elsize(c.src_dtype);
  it.dst = c.dst;
  it.pre_dev = c.pre_dev;
  for (int a = 0; a < 3; ++a) {
    it.src_stride[a] = pre.sign[a] * c.src_stride[a];
    it.dst_stride[a] = c.dst_stride[a];
    it.src_shape[a] = static_cast<int32_t>(pre.size[a]);
    it.src_vlo[a] = static_cast<int32_t>(pre.vlo[a]);
    it.src_vhi[a] = static_cast<int32_t>(pre.vhi[a]);
    it.out_shape[a] = static_cast<int32_t>(affine ? post.size[a] : pre.size[a]);
    it.grid_shape[a] = static_cast<int32_t>(pre.size[a]);
    it.grid_off[a] = static_cast<int32_t>(affine ? post.off[a] : 0);
    it.grid_sign[a] = static_cast<int32_t>(affine ? post.sign[a] : 1);
    it.out_vlo[a] = static_cast<int32_t>(affine ? post.vlo[a] : 0);
    it.out_vhi[a] = static_cast<int32_t>(affine ? post.vhi[a] : pre.size[a]);
    const int64_t m = pre.size[a] > 2 ? pre.size[a] : 2;
    it.nrm[a] = static_cast<float>(2.0 / static_cast<double>(m));
  }
  if (affine) {
    memcpy(it.A, c.A, sizeof(it.A));
  } else {
    static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(it.A, I, sizeof(it.A));
  }
  it.pre_scale = c.pre_scale; it.pre_offset = c.pre_offset;
  it.post_scale = c.post_scale; it.post_offset = c.post_offset;
  it.src_dtype = c.src_dtype;
  it.interp = affine ? c.interp : static_cast<uint8_t>(ADELL_TRILINEAR);
  it.padding = affine ? c.padding : static_cast<uint8_t>(ADELL_PAD_ZEROS);
  uint8_t flags = 0;
  if (!affine) flags |= ADELL_F_IDENTITY;
  if (c.flags & ADELL_CHAIN_STRICT) flags |= ADELL_F_STRICT;
  if (c.pre_dev != nullptr) flags |= ADELL_F_PRE_DEV;
  it.flags = flags;
  return ADELL_OK;
}

}  // namespace

extern "C" int adell_chain_size(void) { return static_cast<int>(sizeof(adell_chain)); }

extern "C" int adell_chain_compose(const adell_chain* chains, int n, adell_item* items_host) {
  if (n < 0 || (n > 0 && (chains == nullptr || items_host == nullptr))) return ADELL_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) {
    const int st = compose_one(chains[i], items_host[i]);
    if (st != ADELL_OK) return st;
  }
  return ADELL_OK;
}

extern "C" int adell_chain_prepare_steps(const adell_chain* chains, void* buf_host, int n_steps, const int32_t* n_items,
                                         const int64_t* item_off, const int64_t* tile_off, adell_launch_info* infos,
                                         int plan_only) {
  if (chains == nullptr || buf_host == nullptr || n_items == nullptr || item_off == nullptr || tile_off == nullptr ||
      infos == nullptr || n_steps < 0)
    return ADELL_ERR_BAD_ARG;
  uint8_t* base = static_cast<uint8_t*>(buf_host);
  int64_t first = 0;
  for (int k = 0; k < n_steps; ++k) {
    adell_item* items = reinterpret_cast<adell_item*>(base + item_off[k]);
    int32_t* tiles = reinterpret_cast<int32_t*>(base + tile_off[k]);
    int st = adell_chain_compose(chains + first, n_items[k], items);
    if (st != ADELL_OK) return st;
    st = plan_only ? adell_aug_plan(items, n_items[k], tiles, infos + k) : adell_aug_prepare(items, n_items[k], tiles, infos + k);
    if (st != ADELL_OK) return st;
    first += n_items[k];
  }
  return ADELL_OK;
}
