/*
 * adell_b200.h — C ABI of the B200-native volumetric augmentation hot path.
 *
 * The reference (CCIG-Champalimaud/adell-mri) is pure Python and has NO FFI layer;
 * the seam this ABI replaces is the per-sample MONAI transform chain that
 * adell_mri/transform_factory assembles and that DataLoader workers execute:
 *
 *   adell_aug_gather          <- RandAffined + RandFlipd + RandSpatialCropd/SpatialPadd/
 *                                CenterSpatialCropd + RandGaussianNoised/RandScale/ShiftIntensityd
 *                                + ConcatItemsd + safe_collate
 *                                (adell_mri/transform_factory/augmentations.py:98-176,255-301,427-515;
 *                                 adell_mri/modules/augmentations.py:165-256;
 *                                 adell_mri/transform_factory/transforms.py:169-214,463-509,787-820;
 *                                 adell_mri/utils/utils.py:308-377)
 *   adell_minmax              <- the two reductions inside monai ScaleIntensityd(minv=0,maxv=1)
 *                                (transforms.py:143-148,430-435,772-777) and
 *                                ConditionalRescalingd/Offsetd (utils/monai_transforms/image_intensity_ops.py:71-74,119-121)
 *   adell_intensity_map       <- the elementwise part of ScaleIntensityd / ConditionalRescalingd /
 *                                Offsetd / ScaleIntensityRange (exact fp32 op order)
 *   adell_hist_pass/_select/
 *   adell_percentile_finalize <- np.percentile inside monai ScaleIntensityRangePercentilesd
 *                                (named by BASELINE.json north_star; new capability, the reference
 *                                 itself only uses the (0,100) special case = min-max)
 *
 * Conventions: every pointer named *_dev is caller-owned DEVICE memory (e.g. from the
 * torch allocator); nothing is allocated, freed or synchronised inside; work is enqueued
 * on `stream` (a cudaStream_t passed as void*).  Every entry point returns ADELL_OK (0)
 * or a negative status and never throws across the ABI.  Re-entrant, no global mutable
 * state.  There is NO CPU fallback: without a CUDA device the compute entry points return
 * ADELL_ERR_NO_DEVICE / ADELL_ERR_LAUNCH.
 */
#ifndef ADELL_B200_H_
#define ADELL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADELL_ABI_VERSION 7

/* status codes */
#define ADELL_OK 0
#define ADELL_ERR_BAD_ARG (-1)
#define ADELL_ERR_DTYPE (-2)
#define ADELL_ERR_ALIGN (-3)
#define ADELL_ERR_LAUNCH (-4)
#define ADELL_ERR_NO_DEVICE (-5)
#define ADELL_ERR_NO_DRIVER (-6)
#define ADELL_ERR_UNSUPPORTED (-7)
#define ADELL_ERR_NO_SPACE (-8) /* a caller-provided buffer is too small (the needed size is reported) */

/* source element types */
#define ADELL_F32 0
#define ADELL_I16 1
#define ADELL_U8 2

/* interpolation (MONAI mode "nearest" / "bilinear") */
#define ADELL_NEAREST 0
#define ADELL_TRILINEAR 1

/* padding_mode of grid_sample */
#define ADELL_PAD_ZEROS 0
#define ADELL_PAD_BORDER 1
#define ADELL_PAD_REFLECTION 2

/* item flags */
#define ADELL_F_IDENTITY 0x01 /* no resample fired: pure integer flip/crop/pad copy (bit-exact)     */
#define ADELL_F_CLIP 0x02     /* clamp each tap to [clip_lo, clip_hi] after the pre map              */
#define ADELL_F_STRICT 0x04   /* ATen operation order for the trilinear sum (mul, add; no fma) and
                                 separate mul/add for post_scale/post_offset                         */
#define ADELL_F_PHILOX 0x08   /* add N(0, noise_std) from Philox4x32-10 (fast mode; not comparable
                                 with the reference's RandomState stream)                            */
#define ADELL_F_PRE_DEV 0x10  /* read {pre_scale, pre_offset} from pre_dev (device-side statistics)  */
#define ADELL_F_TMAP 0x20     /* tmap holds a valid CUtensorMap of the valid source box (staged path)*/
#define ADELL_F_FASTCOORD 0x40 /* trilinear only: incremental coordinates (<=1e-4 contract), no
                                  bit-faithful replay of the MONAI/ATen fp32 coordinate chain         */
#define ADELL_F_WIN_DEV 0x80  /* the resample domain is a WINDOW of a parent volume whose start is only known
                                 on the device (RandCropByPosNegLabeld centres selected by adell_posneg_starts):
                                 src / src_stride describe the window as if it started at parent index (0,0,0),
                                 src_shape is the window's extent, src_vlo must be 0 and src_vhi holds the
                                 PARENT's extents; the kernel reads win_dev[0..2] (int32, 0 <= start <=
                                 parent - window) when it fetches the item and shifts the source by it.
                                 Resampled items need border / reflection padding (zeros padding would have to
                                 stop at the window, which the staged box cannot know: such items are refused) */

/*
 * One unit of work = one (sample, key): the whole reference chain for one volume,
 * composed on the host into canonical form
 *
 *   parent --(integer: crop/pad/flip BEFORE the resample)--> resample domain S
 *          --(RandAffined: A, grid G, interp, padding)-----> grid G
 *          --(integer: flip/crop/pad AFTER the resample)---> output O
 *          --(post intensity map, noise)--------------------> dst
 *
 * Resample domain index t (per axis, 0 <= t < S) addresses  src + sum_a t_a*src_stride[a]
 * (strides are SIGNED: a flip before the resample is a negative stride) and reads literal 0
 * when t_a is outside [src_vlo[a], src_vhi[a]) (constant SpatialPadd band).  Output index o
 * maps to grid index g_a = grid_off[a] + grid_sign[a]*o_a and reads literal 0 when o_a is
 * outside [out_vlo[a], out_vhi[a]) (SpatialPadd after the resample).  Source coordinates follow MONAI/ATen bit for bit:
 *   c_a = g_a - (G_a-1)/2 ;  x_a = fma(A[a][3],1, fma(A[a][2],c_2, fma(A[a][1],c_1, A[a][0]*c_0)))
 *   n_a = x_a * nrm[a]    ;  u_a = ((n_a + 1) * S_a - 1) / 2       (each op rounded to fp32)
 * then padding_mode / floor / rint exactly as ATen's grid_sampler_3d CPU kernel.
 * Axis order is MONAI's [H, W, D] = axes 0,1,2; axis 2 is the contiguous one.
 */
typedef struct __attribute__((aligned(64))) adell_item {
  uint8_t tmap[128];       /* opaque CUtensorMap (see adell_item_encode_tensormap)                  */
  const void* src;         /* element t=(0,0,0); may lie outside the allocation, never read there  */
  float* dst;              /* fp32 output, element o=(0,0,0)                                        */
  const float* noise;      /* optional injected noise, contiguous [O0,O1,O2] fp32, or NULL          */
  const float* pre_dev;    /* optional device {scale, offset} (ADELL_F_PRE_DEV)                     */
  const int32_t* win_dev;  /* ADELL_F_WIN_DEV: device int32[3], start of the source window in the parent  */
  int64_t src_stride[3];   /* signed, in elements                                                   */
  int64_t dst_stride[3];   /* in elements                                                           */
  int32_t src_shape[3];    /* S                                                                      */
  int32_t src_vlo[3];
  int32_t src_vhi[3];
  int32_t out_shape[3];    /* O                                                                      */
  int32_t grid_shape[3];   /* G                                                                      */
  int32_t grid_off[3];
  int32_t grid_sign[3];    /* +1 or -1                                                               */
  int32_t out_vlo[3];
  int32_t out_vhi[3];
  int32_t tmap_off[3];     /* t-space index of the tmap box origin (staged path)                    */
  int32_t tmap_sign[3];    /* +1/-1: box index = tmap_sign*(t - tmap_off) (staged path)             */
  int32_t tmap_box[3];     /* staged box extents actually encoded (z,y,x order = axes 2,1,0)        */
  float A[12];             /* rows 0..2 of the fp32 4x4 MONAI affine (row-major 3x4)                 */
  float nrm[3];            /* (float)(2.0 / max(2, S_a)) — MONAI Resample norm_coords factor        */
  float pre_scale, pre_offset, clip_lo, clip_hi;
  float post_scale, post_offset, noise_std;
  uint64_t philox_seed, philox_offset;
  uint8_t src_dtype, interp, padding, flags;
  /* ---- derived by adell_aug_prepare; callers leave these zero ------------------------------ */
  uint8_t tile_dim[3];     /* output tile extents of this item along axes 0,1,2                     */
  uint8_t kind;            /* ADELL_KIND_*: which K1 path the item's tiles take                     */
  int32_t n_tiles[3];      /* ceil((out_shape + max shear) / tile_dim)                              */
  float fp_smin[3];        /* staged path: footprint of a full tile relative to its origin voxel,   */
  float fp_smax[3];        /*   per source axis (sum of negative / positive D*(tile_dim-1) terms)    */
  int32_t fp_fix;          /* staged path: leading columns (axis 2, box order) of the tensor map
                              that lie before the valid source box (16-byte alignment slack); they
                              are zeroed in shared memory after the TMA load                         */
  double fp_U0[3];         /* un-padded source coordinate of output voxel (0,0,0), fp64             */
  double fp_D[9];          /* d(source coordinate a)/d(output index b), row-major [a][b], fp64      */
  int8_t shear[2][16];     /* staged path: per 8-voxel column group G = o_2 >> 3 along axis 2, the shift
                              (>= 0) of the output tile grid along axes 0 and 1: voxel o belongs to
                              tile ((o_0 + shear[0][G]) / tile_dim[0], (o_1 + shear[1][G]) / tile_dim[1],
                              o_2 / tile_dim[2]).  Chosen so that a column tile's source footprint
                              stays compact under rotations that couple axis 2 into axes 0/1; all
                              zero = plain tile grid.                                                */
  uint8_t dmap[128];       /* opaque CUtensorMap of the destination volume (ADELL_KIND_TSTORE items: their
                              tiles are written by TMA stores).  Completes the struct to 768 bytes.    */
} adell_item;

/* adell_item.kind */
#define ADELL_KIND_GENERIC 0 /* bit-faithful per-voxel path, taps from global memory               */
#define ADELL_KIND_STAGED 1  /* TMA-staged source footprint, taps from shared memory               */
#define ADELL_KIND_VCOPY 2   /* identity item: 128-bit vectorised flip/crop copy                   */
#define ADELL_KIND_TSTORE 3  /* plain identity item (no intensity map, contiguous axis not reversed, aligned
                                rows): TMA box load -> TMA box store, no consumer instructions.  kind =
                                ADELL_KIND_TSTORE + split: 3 whole boxes, 4 plane by plane (axis 0 flipped),
                                5 row by row (axis 1 flipped)                                          */

/* -- library / device ---------------------------------------------------------------- */
int adell_abi_version(void);
const char* adell_status_string(int status);
int adell_item_size(void);           /* sizeof(adell_item), for bindings to cross-check      */
int adell_device_sm_count(int* out); /* number of SMs of the current device                   */

/* Host-only: out[b] = mats[b][0] @ ... @ mats[b][k-1] (4x4 fp32, row-major), every product an
 * fp32 FMA chain in k order = what torch's CPU `@` yields for MONAI's AffineGrid composition
 * (monai AffineGrid: affine = eye @ rotate @ shear @ translate @ scale). */
int adell_mat4_chain(const float* mats, int batch, int k, float* out);

/* Host-only: the MONAI AffineGrid matrix eye @ Rx @ Ry @ Rz @ shear @ translate @ scale of `batch` parameter sets
 * in one call (k_* parameters given per set, 0 = that factor is absent; sin_r / cos_r are the fp32 sines / cosines of
 * the k_rot <= 3 leading rotation angles, evaluated by the caller with torch like MONAI's create_rotate); products are
 * the FMA chains of adell_mat4_chain.  out: [batch][16] row-major. */
int adell_affine_compose(const float* sin_r, const float* cos_r, int k_rot, const float* shear, int k_shear,
                         const float* translate, int k_trans, const float* scale, int k_scale, int batch, float* out);

/* -- K1: fused gather ----------------------------------------------------------------- */
/* What the host learned while preparing one launch. */
typedef struct adell_launch_info {
  int64_t total_tiles; /* output tiles over all items (per-item tile extents: adell_item.tile_dim) */
  int32_t smem_bytes;  /* dynamic shared memory = largest staged source box among the items     */
  int32_t n_staged;    /* items eligible for the TMA-staged path                                */
  int64_t first_copy_tile; /* tiles [0, first_copy_tile) belong to resampled / generic items, the rest to
                              identity (box-copy) items: adell_aug_prepare moves the copy items behind the
                              others so that the kernel can run a memory-bound copy tile and a compute-bound
                              resampled tile side by side on every SM                                   */
} adell_launch_info;

/* Host-only, no GPU work: validates the items, decides per item whether its source footprint
 * can be staged through shared memory by TMA (fp32 source, unit inner stride, 16-byte aligned
 * rows, footprint box <= 100 KiB) and if so encodes the CUtensorMap into items_host[i].tmap
 * (+ tmap_off/tmap_sign/tmap_box, ADELL_F_TMAP) via the driver entry point
 * cuTensorMapEncodeTiled (ADELL_ERR_NO_DRIVER if libcuda is unavailable), and fills
 * tile_start_host[0..n_items] with the exclusive prefix of per-item tile counts, followed by four
 * zero words (tile_start_host must hold n_items + 5 entries): the launch's two chunk queues
 * (resampled tiles, copy tiles) from which the SMs draw their tiles at run time.  The items may be
 * REORDERED in place (identity items last); they are independent, so the result does not change.  The caller then uploads items + prefix + queue words and
 * calls adell_aug_gather; the kernel leaves the queue words zero again, so a buffer can be
 * launched repeatedly (not concurrently with itself).  items_host must start on a 64-byte boundary like every
 * adell_item array (host or device): ADELL_ERR_ALIGN otherwise, also from adell_aug_plan, adell_chain_compose, the
 * *_prepare_steps entry points (through their item offsets) and adell_seq_prepare_steps (buf_host). */
int adell_aug_prepare(adell_item* items_host, int n_items, int32_t* tile_start_host,
                      adell_launch_info* info);
/* Host-only, needs neither a GPU nor a driver: the same decisions as adell_aug_prepare (item order,
 * per-item path, tile shape, column-group shear, staged box, tile prefix) without encoding any tensor
 * map.  For inspection and for tests of the host policy; the result is NOT launchable (info->n_staged
 * comes back as -1 - count and adell_aug_gather refuses it). */
int adell_aug_plan(adell_item* items_host, int n_items, int32_t* tile_start_host, adell_launch_info* info);
/* Host-only convenience for several launches packed in one buffer (one upload for many steps):
 * step k has n_items[k] items at buf_host + item_off[k] (64-byte aligned) and its tile prefix
 * (n_items[k] + 5 int32, see adell_aug_prepare) at buf_host + tile_off[k]; runs adell_aug_prepare on each, filling infos[k]. */
int adell_aug_prepare_steps(void* buf_host, int n_steps, const int32_t* n_items, const int64_t* item_off,
                            const int64_t* tile_off, adell_launch_info* infos);
/* Host-only chain composer: the integer index algebra that collapses one volume's transform chain
 *   parent -> SpatialCrop(crop0) -> flips(flip0) -> [RandAffined A] -> flips(flip1) -> CenterSpatialCrop(crop1) -> intensity
 * (get_augmentations_unet / _class / _ssl single-resample chains,
 *  adell_mri/transform_factory/augmentations.py:98-176,255-320,427-515) into the canonical adell_item,
 * without any Python / numpy work per volume.  Semantics are those of adell_mri_b200/plan.py
 * (BatchPlan.crop / flip / affine / center_crop + _fill_items), which the tests compare byte for byte. */
#define ADELL_CHAIN_AFFINE 0x01 /* the resample fired: A / interp / padding are used            */
#define ADELL_CHAIN_STRICT 0x02 /* ADELL_F_STRICT on the item                                      */
typedef struct adell_chain {
  const void* src;         /* parent volume, element (0,0,0)                                       */
  float* dst;              /* fp32 destination, element (0,0,0) of the output                      */
  const float* pre_dev;    /* optional device {scale, offset} (sets ADELL_F_PRE_DEV) or NULL       */
  const int32_t* win_dev;  /* optional device int32[3]: START of the first crop, chosen on the device (adell_posneg_starts;
                              sets ADELL_F_WIN_DEV): crop0_start must then be 0 and crop0_size the window's extents  */
  int64_t src_stride[3];   /* parent strides in elements                                           */
  int64_t dst_stride[3];
  int32_t src_shape[3];    /* parent extents                                                       */
  int32_t crop0_start[3];  /* window of the first crop; crop0_size all <= 0: no crop                */
  int32_t crop0_size[3];
  int32_t crop1_size[3];   /* CenterSpatialCrop roi after the flips (<= 0 per axis: keep)           */
  float A[12];             /* rows 0..2 of the fp32 MONAI affine (ADELL_CHAIN_AFFINE)               */
  float pre_scale, pre_offset, post_scale, post_offset;
  uint8_t src_dtype, interp, padding, flags;
  uint8_t flip0, flip1;    /* bit a set: torch.flip over axis a before / after the resample         */
  uint8_t reserved_[2];
} adell_chain;
int adell_chain_size(void);
int adell_chain_compose(const adell_chain* chains, int n, adell_item* items_host);
/* Compose + adell_aug_prepare for several steps packed in one (typically pinned) host buffer: step k
 * takes the next n_items[k] chains, writes its items at buf_host + item_off[k] (64-byte aligned) and
 * its tile prefix (n_items[k] + 5 int32) at buf_host + tile_off[k].  plan_only != 0: adell_aug_plan
 * instead (no driver, not launchable; for tests). */
int adell_chain_prepare_steps(const adell_chain* chains, void* buf_host, int n_steps, const int32_t* n_items,
                              const int64_t* item_off, const int64_t* tile_off, adell_launch_info* infos,
                              int plan_only);
/* Host-only SEQUENCE composer: volumes whose chain is an ordered list of up to ADELL_SEQ_MAX_OPS members after one
 * crop — the two-view stream of get_augmentations_ssl, where every sample applies its own drawn ORDER of the fused
 * AugmentationWorkhorsed members (adell_mri/transform_factory/augmentations.py:391-516,
 * adell_mri/modules/augmentations.py:165-256).  A second resample, or anything after a noise member, closes the
 * volume's pass: it is materialised into a scratch fp32 volume by an earlier launch (the reference's sequential
 * resamples).  Closed passes are grouped by LEVEL (a volume's k-th closed pass goes into launch k), the final launch
 * holds every volume.  Semantics — pass closing, pre / post folding of the intensity maps, scratch layout, item order —
 * are those of adell_mri_b200/plan.py (BatchPlan.crop / affine / intensity / add_philox_noise + build_launches) applied
 * slot by slot (slot s = the s-th op of every volume; within a slot: the affines, then the intensity maps, then the
 * noise members), which the tests compare byte for byte.  ADELL_SEQ_FAST composes consecutive affines into one matrix
 * instead of closing (documented deviation of fast mode). */
#define ADELL_OP_NONE 0
#define ADELL_OP_AFFINE 1    /* A (rows 0..2 of the fp32 MONAI matrix), interp, padding; output grid = current size */
#define ADELL_OP_INTENSITY 2 /* v * scale + offset                                                                 */
#define ADELL_OP_PHILOX 3    /* + philox_std * N(0,1) from the device generator (seed, offset)                      */
#define ADELL_SEQ_MAX_OPS 4
#define ADELL_SEQ_FAST 0x01
#define ADELL_SEQ_STRICT 0x02
typedef struct adell_seq_op {
  float A[12];
  double scale, offset;
  uint64_t philox_seed, philox_offset;
  float philox_std;
  uint8_t kind, interp, padding, reserved_;
} adell_seq_op;
typedef struct adell_seq {
  const void* src;        /* parent volume, element (0,0,0)                              */
  float* dst;             /* fp32 destination of the LAST pass                           */
  int64_t src_stride[3];  /* parent strides in elements                                  */
  int64_t dst_stride[3];
  int32_t src_shape[3];
  int32_t crop0_start[3]; /* SpatialCrop before the members; crop0_size all <= 0: none   */
  int32_t crop0_size[3];
  uint8_t src_dtype, n_ops, flags, reserved_;
  adell_seq_op ops[ADELL_SEQ_MAX_OPS];
} adell_seq;
/* One launch of a composed step: its items start at buf_host + item_off (n_items x 768 B, followed by the int32 tile
 * prefix of n_items + 5 words); launches of one step are listed in execution order. */
typedef struct adell_seq_launch {
  int64_t item_off;
  int32_t n_items;
  int32_t step;
  adell_launch_info info;
} adell_seq_launch;
int adell_seq_size(void);
/* Composes n_steps consecutive steps (step k = the next n_vols[k] sequences) into buf_host (buf_bytes available;
 * typically pinned) and prepares every launch (adell_aug_prepare; prepare_mode 1: adell_aug_plan, 2: composed items
 * only — tests).  Scratch volumes are placed in [scratch_dev, scratch_dev + 4 * scratch_elems) — every step starts
 * again at scratch_dev: launches of one stream are ordered.  *n_launches, *bytes_used and *scratch_used are always
 * reported; ADELL_ERR_NO_SPACE when buf_bytes, max_launches or scratch_elems is too small (nothing launchable then). */
int adell_seq_prepare_steps(const adell_seq* seqs, int n_steps, const int32_t* n_vols, uint64_t scratch_dev,
                            int64_t scratch_elems, void* buf_host, int64_t buf_bytes, adell_seq_launch* launches,
                            int max_launches, int32_t* n_launches, int64_t* bytes_used, int64_t* scratch_used,
                            int prepare_mode);
/* Enqueue the fused gather over all items: ONE kernel launch per call. */
int adell_aug_gather(const adell_item* items_dev, const int32_t* tile_start_dev, int n_items,
                     const adell_launch_info* info, void* stream);
/* Number of kernel launches the last-compiled policy issues per adell_aug_gather call. */
int adell_aug_gather_launches(void);

/* -- statistics for intensity normalisation -------------------------------------------- */
/* One descriptor per volume for the batched statistics kernels. */
typedef struct adell_vol {
  const void* data; /* contiguous volume                              */
  int64_t n;        /* number of elements                             */
  int32_t dtype;    /* ADELL_F32 / ADELL_I16 / ADELL_U8               */
  int32_t _pad;
} adell_vol;

/* out_dev[2*v+0] = min, out_dev[2*v+1] = max over volume v (as fp32).  out_dev must be
 * pre-initialised by the call itself (it is: +inf/-inf written by a first tiny kernel). */
int adell_minmax(const adell_vol* vols_dev, int n_vols, int64_t max_n, float* out_dev, void* stream);

/* out_dev[2*v+0] = mean, out_dev[2*v+1] = population standard deviation of volume v (1 when it
 * is 0), as monai NormalizeIntensityd computes them (named by BASELINE.json north_star; the
 * reference itself never calls it).  fp64 accumulation in acc_dev (3 doubles per volume, scratch,
 * zeroed by the call); nonzero != 0 restricts the statistics to elements != 0. */
int adell_meanstd(const adell_vol* vols_dev, int n_vols, int64_t max_n, int nonzero, double* acc_dev,
                  float* out_dev, void* stream);
#define ADELL_MEANSTD_NONZERO 1 /* bit 0 of `nonzero`: statistics over elements != 0                  */
#define ADELL_MEANSTD_RAW_STD 2 /* bit 1 of `nonzero`: report a zero std as 0 (RandStdShiftIntensityd:
                                   offset = factor * std), not as NormalizeIntensityd's 1             */

/* Label construction of the cached stage (CombineBinaryLabelsd + LabelOperatorSegmentationd,
 * /root/reference/adell_mri/utils/monai_transforms/labels.py:123-220, called from
 * transform_factory/transforms.py:181-194): n_src (<= 8) label maps of n elements each
 * (src_dev / dtypes are HOST arrays of device pointers / ADELL_* dtypes) are combined voxel-wise, then
 * mapped through `table` (host array, <= 16 values).  dst is fp32. */
#define ADELL_LABEL_COMBINE_NONE 0     /* one map, as is                                       */
#define ADELL_LABEL_COMBINE_ANY 1      /* float(sum over the maps > 0)                         */
#define ADELL_LABEL_COMBINE_MAJORITY 2 /* float(mean over the maps > 0.5)                      */
#define ADELL_LABEL_OP_NONE 0          /* values unchanged                                     */
#define ADELL_LABEL_OP_BINARY 1        /* 1 where the value is in table (positive labels)      */
#define ADELL_LABEL_OP_CATEGORICAL 2   /* index of the value in table (possible labels), else 0 */
int adell_label_map(const void* const* src_dev, const int32_t* dtypes, int n_src, int combine, int op,
                    const float* table, int n_table, float* dst_dev, int64_t n, void* stream);

/* Bounding box of the non-zero voxels of each [S0,S1,S2] volume (shapes_dev: 3 int32 per volume):
 * out_dev[6*v ..] = {lo0, hi0, lo1, hi1, lo2, hi2}, hi exclusive; an empty mask yields lo = INT32_MAX,
 * hi = 0.  The reduction behind CropFromMaskd
 * (/root/reference/adell_mri/utils/monai_transforms/labels.py:412-522: `torch.where(mask)` + min/max). */
int adell_mask_bbox(const adell_vol* vols_dev, const int32_t* shapes_dev, int n_vols, int64_t max_n,
                    int32_t* out_dev, void* stream);

/* monai AdjustContrast (RandAdjustContrastd, --augment intensity;
 * /root/reference/adell_mri/transform_factory/augmentations.py:66-76,219-232):
 *   y = pow((x - min) / (range + 1e-7f), gamma) * range + min,  range = max - min,
 * every op rounded to fp32 in that order; {min, max} per volume are read from minmax_dev[2*v..]
 * (adell_minmax), the exponent from gamma_dev[v].  dst is fp32, contiguous. */
int adell_gamma_map(const adell_vol* vols_dev, float* const* dst_dev, const float* minmax_dev,
                    const float* gamma_dev, int n_vols, int64_t max_n, void* stream);

/* -- K5: Resized ------------------------------------------------------------------------- */
#define ADELL_RESIZE_AREA 0     /* F.interpolate(mode="area") = ATen adaptive_avg_pool3d, op for op (bit-identical) */
#define ADELL_RESIZE_NEAREST 1  /* F.interpolate(mode="nearest") (legacy index rule)                               */
/* monai.transforms.Resized of the reference's scaled crop
 * (/root/reference/adell_mri/transform_factory/augmentations.py:427-444) and of its cached stage
 * (/root/reference/adell_mri/transform_factory/transforms.py:157-167,455-462): volume v, contiguous fp32
 * [I0,I1,I2] = in_shapes_dev[3v..] at src_dev[v], is resized to the contiguous fp32 [O0,O1,O2] =
 * out_shape (HOST array of 3) at dst_dev[v].  src_dev / dst_dev are DEVICE arrays of device pointers. */
int adell_resize(const float* const* src_dev, const int32_t* in_shapes_dev, float* const* dst_dev, int n_vols,
                 const int32_t* out_shape, int mode, void* stream);

/* monai RandRicianNoise (the SSL workhorse's `rician_noise` member,
 * /root/reference/adell_mri/modules/augmentations.py:53,86,117; RandRicianNoised of --augment noise,
 * /root/reference/adell_mri/transform_factory/augmentations.py:81-91):
 *   y = sqrt((x + n1)^2 + n2^2)
 * with the two host-drawn normal volumes n1, n2 (RandomState.normal, float64 -> fp32) resident on the
 * device; add, squares, add and square root each rounded to fp32 in that order.  All four buffers
 * are contiguous fp32 of n elements; dst may alias x. */
int adell_rician_map(const float* x_dev, const float* noise1_dev, const float* noise2_dev, float* dst_dev,
                     int64_t n, void* stream);

/* Exact elementwise intensity program  y = ((x*m0 - a)/d)*m1*m2 + b  with every op rounded to
 * fp32 in that order and no-op steps skipped bit-exactly (m0=1, a=0, d=1, m1=1, m2=1, b=0); the
 * six coefficients are read from coef_dev[6*v ..] so they can come from device statistics.
 * Optional clamp to [clip_lo, clip_hi] when clip != 0.  dst is fp32, contiguous. */
int adell_intensity_map(const adell_vol* vols_dev, float* const* dst_dev, const float* coef_dev,
                        int n_vols, int64_t max_n, int clip, float clip_lo, float clip_hi,
                        void* stream);
/* Fills coef_dev from device min/max per the reference's scalers (see ADELL_SCALER_*). */
#define ADELL_SCALER_MINMAX 0    /* ScaleIntensityd(minv,maxv): ((x-min)/(max-min))*(maxv-minv)+minv */
#define ADELL_SCALER_ADC_SEG 1   /* ConditionalRescalingd(500,.001) -> ScaleIntensityd(factor=-2/3)  */
#define ADELL_SCALER_ADC_CLASS 2 /* ConditionalRescalingd -> Offsetd(None) -> ScaleIntensityd(factor) */
#define ADELL_SCALER_RANGE 3     /* ScaleIntensityRange(a_min=lo,a_max=hi,b_min=p0,b_max=p1) as used by
                                    ScaleIntensityRangePercentilesd; stats = the two percentiles     */
#define ADELL_SCALER_ZSCORE 4     /* NormalizeIntensityd: (x - mean) / std; stats = {mean, std}           */
/* stats_dev holds {lo, hi} per volume (min/max from adell_minmax, or two percentiles).
 * MINMAX: p0=minv, p1=maxv.  ADC_*: p0=max_value (500), p1=scale (0.001); factor fixed at -2/3
 * (ADC_FACTOR, transforms.py:25). */
int adell_scaler_coefs(const float* stats_dev, int n_vols, int scaler, double p0, double p1,
                       float* coef_dev, void* stream);
/* Collapses the 6-coefficient program into the fused-mode pair {scale, offset} that
 * adell_item.pre_dev points at (x*scale+offset; <=1e-4 contract instead of bit-exact). */
int adell_coefs_to_affine(const float* coef_dev, int n_vols, float* pre_dev_out, void* stream);

/* Radix histogram of the order-preserving 32-bit key of every element (fp32: sign-flip
 * transform; int16/uint8: biased value).  A pass looks at key bits [shift, shift+bits) of the
 * elements whose bits above shift+bits equal those of prefix_dev[h*n_sel+s] (first pass,
 * shift+bits==32: every element, one histogram shared by all selections) and counts into
 *   bins_dev[((h*n_sel_eff)+s) << bits],  h = shared ? 0 : v,  n_sel_eff = first pass ? 1 : n_sel.
 * shared!=0 accumulates every volume into histogram 0 (dataset-wide statistics: all-reduce
 * bins_dev across ranks between adell_hist_pass and adell_hist_select).  bins must be zeroed
 * by the caller (torch.zero_ / cudaMemsetAsync) before each pass. */
int adell_hist_pass(const adell_vol* vols_dev, int n_vols, int64_t max_n, int n_sel, int shared,
                    const uint32_t* prefix_dev, int pass_shift, int pass_bits, uint64_t* bins_dev,
                    void* stream);
/* For every (histogram h, selection s): walk the bins, find the bin holding rank_dev[h*n_sel+s]
 * (0-based order statistic among the elements counted there), OR its index << shift into
 * prefix_dev[h*n_sel+s] and subtract the count of the preceding bins from the rank.  After the
 * last pass (shift==0) prefix_dev holds the full key of the order statistic. */
int adell_hist_select(const uint64_t* bins_dev, int n_hist, int n_sel, int pass_shift, int pass_bits,
                      uint32_t* prefix_dev, uint64_t* rank_dev, void* stream);
/* The same order statistics in ONE full read of every volume (per-volume statistics only; fp32 reads a volume three
 * times through the passes above): a strided sample of <= 32 Ki keys brackets each quantile's (lo, hi) ranks between
 * two sample order statistics 6 standard deviations apart, one streaming pass counts the keys beyond the bracket and
 * equal to its ends and lists the keys strictly inside (a few 0.1 % of the volume), the ranks are then selected
 * exactly among those.  A bracket that misses, or overflows its list, un-gates the three radix passes inside the same
 * call, so keys_dev always holds the exact keys ([n_vols][n_q][2], like the prefix array adell_hist_select leaves after
 * the last pass: feed it to adell_percentile_finalize).  rank_dev: [n_vols][n_q][2] 0-based (lo, hi) ranks, not
 * modified.  workspace_dev: 256-byte aligned scratch of adell_quantile_workspace(...) bytes.  n_q <= 4; all volumes of
 * element type `dtype`. */
/* pooled_n > 0: POOLED statistics over all n_vols volumes (pooled_n = their total element count; dataset-wide
 * percentiles of one rank): rank_dev / keys_dev are then [1][n_q][2]; at most 64 volumes of >= 128 Ki elements in
 * total (else ADELL_ERR_UNSUPPORTED: use the radix passes). */
int64_t adell_quantile_workspace(int n_vols, int n_q, int64_t max_n, int64_t pooled_n);
/* flags: ADELL_QUANTILE_REUSE_BRACKETS keeps the brackets the previous call with the same arguments left in this
 * workspace instead of sampling again (the same cached volumes come back every epoch; a bracket is a hint: one that
 * misses — data changed in place — un-gates the fallback, so the keys stay exact). */
#define ADELL_QUANTILE_REUSE_BRACKETS 0x01
int adell_quantile_keys(const adell_vol* vols_dev, int n_vols, int64_t max_n, int64_t pooled_n, int dtype, int n_q,
                        const uint64_t* rank_dev, uint32_t* keys_dev, void* workspace_dev, int64_t workspace_bytes, int flags,
                        void* stream);
/* Test aid (synchronises): *out_host = 1 when the last adell_quantile_keys on this workspace took the radix fallback. */
int adell_quantile_fell_back(const void* workspace_dev, int n_vols, int n_q, int64_t max_n, int64_t pooled_n, int* out_host);
/* Converts selected keys back to values and evaluates numpy's 'linear' percentile
 *   v = a + (b-a)*t  (t<0.5)   |   v = b - (b-a)*(1-t)  (t>=0.5)      in float64, cast to fp32
 * for n_q quantiles per volume from the (lo, hi) order statistics; keys_dev is
 * [n_vols][n_q][2], frac_dev [n_vols][n_q] float64, out_dev [n_vols][n_q] fp32. */
int adell_percentile_finalize(const uint32_t* keys_dev, const double* frac_dev, int n_vols, int n_q,
                              int dtype, float* out_dev, void* stream);

/* -- RandCropByPosNegLabeld: crop windows selected on the device ----------------------------- */
/* One crop: the host made the two draws of monai generate_pos_neg_label_crop_centers
 * (`R.rand() < pos_ratio` picks the foreground or the background list, `R.randint(len(list))` the entry:
 * /root/reference/adell_mri/transform_factory/augmentations.py:147-158 -> RandCropByPosNegLabeld); the index
 * lists themselves (FgBgToIndicesd, transform_factory/transforms.py:196-203) stay on the device. */
typedef struct adell_posneg {
  const int64_t* indices; /* device: the chosen list (flat voxel indices of the label volume)   */
  int64_t pick;           /* host draw: entry of that list                                        */
  int32_t shape[3];       /* label volume extents                                                 */
  int32_t size[3];        /* crop size (clipped to shape)                                         */
} adell_posneg;
/* starts_dev[3*i..] = start of crop i (centre corrected like monai correct_crop_centers, then
 * SpatialCrop(roi_center, roi_size)): what an ADELL_F_WIN_DEV item's win_dev points at. */
int adell_posneg_starts(const adell_posneg* crops_dev, int n_crops, int32_t* starts_dev, void* stream);

/* -- K4: batch-level mixing after collation ------------------------------------------------ */
/* (partial) mixup of a collated fp32 batch [batch, per_sample] in one pass
 * (/root/reference/adell_mri/utils/batch_preprocessing.py:31-118):
 *   out[b] = x[b] * factor[b] + x[perm[b]] * (1 - factor[b])     (fp32: mul, mul, add, as torch)
 * for the samples with sel[b] != 0 (sel_dev == NULL: all of them); the others are copied.
 * out_dev must not alias x_dev.  HBM-bound: 8 B per element algorithmically (12 B when the partner
 * sample is no longer in L2). */
int adell_mixup(const float* x_dev, float* out_dev, const float* factor_dev, const int32_t* perm_dev,
                const uint8_t* sel_dev, int batch, int64_t per_sample, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADELL_B200_H_ */
