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
#include <stdlib.h>
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
  const bool win = c.win_dev != nullptr;   // BatchPlan.crop_from_device: the window's start lives in device memory
  if (win) {
    if (!has_crop0) return ADELL_ERR_BAD_ARG;
    for (int a = 0; a < 3; ++a) {
      if (c.crop0_start[a] != 0) return ADELL_ERR_BAD_ARG;
      st[a] = 0;
      const int64_t want = c.crop0_size[a] > 0 ? c.crop0_size[a] : pre.size[a];
      n[a] = want < pre.size[a] ? want : pre.size[a];
    }
    pre.crop(st, n);
  } else if (has_crop0) {   // BatchPlan.crop: start clipped to [0, cur], size to cur - start
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
    if (win) {
      // ADELL_F_WIN_DEV: src_vhi carries the parent's extent counted from the window's lowest element at start 0
      const int64_t lo = pre.sign[a] > 0 ? pre.off[a] : pre.off[a] - (pre.size[a] - 1);
      it.src_vhi[a] = static_cast<int32_t>(c.src_shape[a] - lo);
    }
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
  if (win) { flags |= ADELL_F_WIN_DEV; it.win_dev = c.win_dev; }
  it.flags = flags;
  return ADELL_OK;
}

}  // namespace

extern "C" int adell_chain_size(void) { return static_cast<int>(sizeof(adell_chain)); }

extern "C" int adell_chain_compose(const adell_chain* chains, int n, adell_item* items_host) {
  if (n < 0 || (n > 0 && (chains == nullptr || items_host == nullptr))) return ADELL_ERR_BAD_ARG;
  if (n > 0 && (reinterpret_cast<uintptr_t>(items_host) & 63u) != 0) return ADELL_ERR_ALIGN;  // adell_item is 64-byte aligned
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

// ---------------------------------------------------------------------------------------------------------------
// Sequence composer (adell_seq): plan.BatchPlan's multi-pass semantics for ordered member lists, in C.
// The state of one volume's OPEN pass is plan._Stage restricted to what the members can produce (no clip, no
// device-side pre map, no injected noise tensor, no device-side window: chains that need those stay on BatchPlan).
namespace {

struct SeqStage {
  IntMap3 pre, post;
  bool has_affine;
  float A[12];
  uint8_t interp, padding;
  double pre_s, pre_o, post_s, post_o;
  float philox_std;
  uint64_t philox_seed, philox_off;
  int64_t grid[3];
  void init(const int64_t* size) {
    pre.init64(size);
    post.init64(size);
    has_affine = false;
    static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    memcpy(A, I, sizeof(A));
    interp = ADELL_TRILINEAR;
    padding = ADELL_PAD_ZEROS;
    pre_s = 1.0; pre_o = 0.0; post_s = 1.0; post_o = 0.0;
    philox_std = 0.0f; philox_seed = 0; philox_off = 0;
    for (int a = 0; a < 3; ++a) grid[a] = size[a];
  }
  bool noise() const { return philox_std != 0.0f; }
  const int64_t* out_size() const { return has_affine ? post.size : pre.size; }
  bool pre_invalid() const {
    for (int a = 0; a < 3; ++a)
      if (pre.vlo[a] > 0 || pre.vhi[a] < pre.size[a]) return true;
    return false;
  }
};

struct SeqVol {
  SeqStage st;
  uint64_t parent;       // address of the current parent volume (the caller's, or the scratch of the previous pass)
  int64_t pstride[3];
  uint8_t pdtype;
  int level;             // closed passes so far
};

// plan.BatchPlan._fill_items for one volume
void seq_fill_item(const SeqStage& st, uint64_t parent, const int64_t* pstride, uint8_t pdtype, uint64_t dst,
                   const int64_t* dst_stride, bool strict, adell_item& it) {
  memset(&it, 0, sizeof(it));
  int64_t elem_off = 0;
  for (int a = 0; a < 3; ++a) elem_off += st.pre.off[a] * pstride[a];
  it.src = reinterpret_cast<const void*>(parent + static_cast<uint64_t>(elem_off * elsize(pdtype)));
  it.dst = reinterpret_cast<float*>(dst);
  const bool ha = st.has_affine;
  for (int a = 0; a < 3; ++a) {
    it.src_stride[a] = st.pre.sign[a] * pstride[a];
    it.dst_stride[a] = dst_stride[a];
    it.src_shape[a] = static_cast<int32_t>(st.pre.size[a]);
    it.src_vlo[a] = static_cast<int32_t>(st.pre.vlo[a]);
    it.src_vhi[a] = static_cast<int32_t>(st.pre.vhi[a]);
    it.out_shape[a] = static_cast<int32_t>(ha ? st.post.size[a] : st.pre.size[a]);
    it.grid_shape[a] = static_cast<int32_t>(ha ? st.grid[a] : st.pre.size[a]);
    it.grid_off[a] = static_cast<int32_t>(ha ? st.post.off[a] : 0);
    it.grid_sign[a] = static_cast<int32_t>(ha ? st.post.sign[a] : 1);
    it.out_vlo[a] = static_cast<int32_t>(ha ? st.post.vlo[a] : 0);
    it.out_vhi[a] = static_cast<int32_t>(ha ? st.post.vhi[a] : st.pre.size[a]);
    const int64_t m = st.pre.size[a] > 2 ? st.pre.size[a] : 2;
    it.nrm[a] = static_cast<float>(2.0 / static_cast<double>(m));
  }
  memcpy(it.A, st.A, sizeof(it.A));
  it.pre_scale = static_cast<float>(st.pre_s); it.pre_offset = static_cast<float>(st.pre_o);
  it.post_scale = static_cast<float>(st.post_s); it.post_offset = static_cast<float>(st.post_o);
  it.noise_std = st.philox_std;
  it.philox_seed = st.philox_seed; it.philox_offset = st.philox_off;
  it.src_dtype = pdtype;
  it.interp = st.interp; it.padding = st.padding;
  uint8_t flags = 0;
  if (!ha) flags |= ADELL_F_IDENTITY;
  if (strict) flags |= ADELL_F_STRICT;
  if (st.philox_std != 0.0f) flags |= ADELL_F_PHILOX;
  it.flags = flags;
}

// Growable array of items per level (malloc: no C++ runtime containers across this file's C-style code)
struct ItemVec {
  adell_item* p = nullptr;
  int n = 0, cap = 0;
  adell_item* push() {
    if (n == cap) {
      const int nc = cap ? 2 * cap : 64;
      adell_item* q = static_cast<adell_item*>(realloc(p, sizeof(adell_item) * static_cast<size_t>(nc)));
      if (q == nullptr) return nullptr;
      p = q; cap = nc;
    }
    return p + n++;
  }
  ~ItemVec() { free(p); }
};

struct SeqStep {
  SeqVol* vols;
  int n;
  bool strict;
  ItemVec levels[ADELL_SEQ_MAX_OPS * 2 + 1];   // a slot closes a volume at most twice (before the affine, before the fold)
  int max_level = 0;
  uint64_t scratch_dev;
  int64_t scratch_off = 0;     // elements
  bool oom = false;

  // BatchPlan._close for one volume: its open pass becomes an item writing a scratch volume
  void close(int v) {
    SeqVol& x = vols[v];
    const int64_t* os = x.st.out_size();
    const int64_t size[3] = {os[0], os[1], os[2]};
    const int64_t vox = size[0] * size[1] * size[2];
    const uint64_t tptr = scratch_dev + 4ull * static_cast<uint64_t>(scratch_off);
    scratch_off += (vox + 63) / 64 * 64;   // every scratch volume starts on a 256-byte boundary
    const int64_t cstride[3] = {size[1] * size[2], size[2], 1};
    adell_item* it = levels[x.level].push();
    if (it == nullptr) { oom = true; return; }
    seq_fill_item(x.st, x.parent, x.pstride, x.pdtype, tptr, cstride, strict, *it);
    if (x.level + 1 > max_level) max_level = x.level + 1;
    ++x.level;
    x.st.init(size);
    x.parent = tptr;
    for (int a = 0; a < 3; ++a) x.pstride[a] = cstride[a];
    x.pdtype = ADELL_F32;
  }
};

int seq_compose_step(const adell_seq* seqs, int n, uint64_t scratch_dev, SeqStep& S) {
  S.n = n;
  S.scratch_dev = scratch_dev;
  S.vols = static_cast<SeqVol*>(malloc(sizeof(SeqVol) * static_cast<size_t>(n > 0 ? n : 1)));
  if (S.vols == nullptr) return ADELL_ERR_BAD_ARG;
  bool strict = n > 0, fast = false;
  int max_ops = 0;
  for (int v = 0; v < n; ++v) {
    const adell_seq& q = seqs[v];
    if (q.src == nullptr || q.dst == nullptr || q.src_dtype > ADELL_U8 || q.n_ops > ADELL_SEQ_MAX_OPS) return ADELL_ERR_BAD_ARG;
    strict = strict && (q.flags & ADELL_SEQ_STRICT);
    fast = fast || (q.flags & ADELL_SEQ_FAST);
    if (q.n_ops > max_ops) max_ops = q.n_ops;
    SeqVol& x = S.vols[v];
    int64_t size[3] = {q.src_shape[0], q.src_shape[1], q.src_shape[2]};
    x.st.init(size);
    x.parent = reinterpret_cast<uint64_t>(q.src);
    for (int a = 0; a < 3; ++a) x.pstride[a] = q.src_stride[a];
    x.pdtype = q.src_dtype;
    x.level = 0;
    bool has_crop0 = false;
    for (int a = 0; a < 3; ++a) has_crop0 |= q.crop0_size[a] > 0;
    if (has_crop0) {   // BatchPlan.crop: start clipped to [0, cur], size to cur - start
      int64_t st[3], nn[3];
      for (int a = 0; a < 3; ++a) {
        st[a] = IntMap3::clip(q.crop0_start[a], 0, size[a]);
        const int64_t want = q.crop0_size[a] > 0 ? q.crop0_size[a] : size[a];
        nn[a] = want < size[a] - st[a] ? want : size[a] - st[a];
      }
      x.st.pre.crop(st, nn);
      for (int a = 0; a < 3; ++a) x.st.grid[a] = x.st.pre.size[a];
    }
  }
  S.strict = strict;   // (BatchPlan: one default for the whole plan)
  unsigned char* mark = static_cast<unsigned char*>(malloc(static_cast<size_t>(n > 0 ? n : 1)));
  if (mark == nullptr) return ADELL_ERR_BAD_ARG;
  for (int s = 0; s < max_ops; ++s) {
    // ---- affines of this slot (BatchPlan.affine)
    for (int v = 0; v < n; ++v) {
      mark[v] = 0;
      const adell_seq& q = seqs[v];
      if (s >= q.n_ops || q.ops[s].kind != ADELL_OP_AFFINE) continue;
      SeqStage& st = S.vols[v].st;
      mark[v] = 1;
      if (fast && st.has_affine && !st.noise()) {
        bool untouched = st.post_s == 1.0 && st.post_o == 0.0;
        for (int a = 0; a < 3; ++a)
          untouched = untouched && st.post.off[a] == 0 && st.post.sign[a] == 1 && st.post.size[a] == st.post.vhi[a] &&
                      st.post.vlo[a] == 0 && st.post.size[a] == st.pre.size[a];
        if (untouched) {   // M = M_prev @ A in float64, rows 0..2 back to fp32
          const float* B = q.ops[s].A;
          float out[12];
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) {
              double acc = 0.0;
              for (int k = 0; k < 3; ++k) acc += static_cast<double>(st.A[4 * r + k]) * static_cast<double>(B[4 * k + c]);
              if (c == 3) acc += static_cast<double>(st.A[4 * r + 3]);
              out[4 * r + c] = static_cast<float>(acc);
            }
          memcpy(st.A, out, sizeof(out));
          mark[v] = 0;
        }
      }
    }
    for (int v = 0; v < n; ++v)
      if (mark[v] && (S.vols[v].st.has_affine || S.vols[v].st.noise())) S.close(v);
    // a pending post map precedes the resample: folded into the per-tap pre map — unless zero-filled (invalid)
    // taps would pick up its offset: those volumes are closed first
    for (int v = 0; v < n; ++v) {
      if (!mark[v]) continue;
      const SeqStage& st = S.vols[v].st;
      const bool foldable = st.post_s != 1.0 || st.post_o != 0.0;
      if (foldable && st.pre_invalid() && st.post_o != 0.0) S.close(v);
    }
    for (int v = 0; v < n; ++v) {
      if (!mark[v]) continue;
      SeqStage& st = S.vols[v].st;
      const adell_seq_op& op = seqs[v].ops[s];
      if (st.post_s != 1.0 || st.post_o != 0.0) {
        st.pre_o = st.pre_o * st.post_s + st.post_o;
        st.pre_s = st.pre_s * st.post_s;
        st.post_s = 1.0; st.post_o = 0.0;
      }
      st.has_affine = true;
      memcpy(st.A, op.A, sizeof(st.A));
      st.interp = op.interp; st.padding = op.padding;
      for (int a = 0; a < 3; ++a) st.grid[a] = st.pre.size[a];
      st.post.init64(st.pre.size);
    }
    // ---- intensity maps of this slot (BatchPlan.intensity)
    for (int v = 0; v < n; ++v) {
      const adell_seq& q = seqs[v];
      mark[v] = s < q.n_ops && q.ops[s].kind == ADELL_OP_INTENSITY;
      if (mark[v] && S.vols[v].st.noise()) S.close(v);
    }
    for (int v = 0; v < n; ++v) {
      if (!mark[v]) continue;
      SeqStage& st = S.vols[v].st;
      const double scale = seqs[v].ops[s].scale, offset = seqs[v].ops[s].offset;
      if (st.has_affine || st.pre_invalid()) {
        st.post_o = st.post_o * scale + offset;
        st.post_s = st.post_s * scale;
      } else {
        st.pre_o = st.pre_o * scale + offset;
        st.pre_s = st.pre_s * scale;
      }
    }
    // ---- noise members of this slot (BatchPlan.add_philox_noise)
    for (int v = 0; v < n; ++v) {
      const adell_seq& q = seqs[v];
      mark[v] = s < q.n_ops && q.ops[s].kind == ADELL_OP_PHILOX;
      if (mark[v] && S.vols[v].st.noise()) S.close(v);
    }
    for (int v = 0; v < n; ++v) {
      if (!mark[v]) continue;
      SeqStage& st = S.vols[v].st;
      const adell_seq_op& op = seqs[v].ops[s];
      st.philox_std = op.philox_std; st.philox_seed = op.philox_seed; st.philox_off = op.philox_offset;
    }
    if (S.oom) { free(mark); return ADELL_ERR_BAD_ARG; }
  }
  free(mark);
  return ADELL_OK;
}

}  // namespace

extern "C" int adell_seq_size(void) { return static_cast<int>(sizeof(adell_seq)); }

extern "C" int adell_seq_prepare_steps(const adell_seq* seqs, int n_steps, const int32_t* n_vols, uint64_t scratch_dev,
                                       int64_t scratch_elems, void* buf_host, int64_t buf_bytes, adell_seq_launch* launches,
                                       int max_launches, int32_t* n_launches, int64_t* bytes_used, int64_t* scratch_used,
                                       int prepare_mode) {
  if (seqs == nullptr || n_vols == nullptr || buf_host == nullptr || launches == nullptr || n_launches == nullptr ||
      bytes_used == nullptr || scratch_used == nullptr || n_steps < 0 || buf_bytes < 0 || max_launches < 0)
    return ADELL_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(buf_host) & 63u) != 0) return ADELL_ERR_ALIGN;  // items are written at 64-byte multiples of it
  uint8_t* base = static_cast<uint8_t*>(buf_host);
  int64_t off = 0, smax = 0, first = 0;
  int nl = 0;
  bool fits = true;
  for (int k = 0; k < n_steps; ++k) {
    const int n = n_vols[k];
    if (n < 0) return ADELL_ERR_BAD_ARG;
    SeqStep S;
    int st = seq_compose_step(seqs + first, n, scratch_dev, S);
    if (st != ADELL_OK) { free(S.vols); return st; }
    if (S.scratch_off > smax) smax = S.scratch_off;
    const bool scratch_ok = S.scratch_off <= scratch_elems;
    for (int l = 0; l <= S.max_level; ++l) {
      const bool last = l == S.max_level;
      const int cnt = last ? n : S.levels[l].n;
      if (cnt == 0) continue;
      const int64_t bytes = (static_cast<int64_t>(cnt) * static_cast<int64_t>(sizeof(adell_item)) + 4 * (cnt + 5) + 127) / 128 * 128;
      const bool room = fits && scratch_ok && off + bytes <= buf_bytes && nl < max_launches;
      if (room) {
        adell_item* items = reinterpret_cast<adell_item*>(base + off);
        if (last) {
          for (int v = 0; v < n; ++v) {
            const SeqVol& x = S.vols[v];
            seq_fill_item(x.st, x.parent, x.pstride, x.pdtype, reinterpret_cast<uint64_t>(seqs[first + v].dst),
                          seqs[first + v].dst_stride, S.strict, items[v]);
          }
        } else {
          memcpy(items, S.levels[l].p, sizeof(adell_item) * static_cast<size_t>(cnt));
        }
        int32_t* tiles = reinterpret_cast<int32_t*>(base + off + static_cast<int64_t>(cnt) * static_cast<int64_t>(sizeof(adell_item)));
        adell_seq_launch& L = launches[nl];
        L.item_off = off; L.n_items = cnt; L.step = k;
        memset(&L.info, 0, sizeof(L.info));
        if (prepare_mode == 0) st = adell_aug_prepare(items, cnt, tiles, &L.info);
        else if (prepare_mode == 1) st = adell_aug_plan(items, cnt, tiles, &L.info);
        else st = ADELL_OK;
        if (st != ADELL_OK) { free(S.vols); return st; }
      } else {
        fits = false;
      }
      off += bytes;
      ++nl;
    }
    free(S.vols);
    if (!scratch_ok) fits = false;
    first += n;
  }
  *n_launches = nl;
  *bytes_used = off;
  *scratch_used = smax;
  return fits ? ADELL_OK : ADELL_ERR_NO_SPACE;
}
