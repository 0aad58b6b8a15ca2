// K5 — `Resized` of the reference's optional scaled crop and of its cached stage:
// torch.nn.functional.interpolate(mode="area") = ATen adaptive_avg_pool3d, and mode="nearest".
//
// Replaces monai.transforms.Resized in
//   /root/reference/adell_mri/transform_factory/augmentations.py:427-444  (--scaled_crop_size: SpatialPadd ->
//     RandSpatialCropd(random_size=True) -> Resized(scaled_crop_size), MONAI default mode "area")
//   /root/reference/adell_mri/transform_factory/transforms.py:157-167,455-462 (resize_size, "area" for images,
//     "nearest" for label maps).
// Area: output voxel (od, oh, ow) averages the input window [floor(o*I/O), ceil((o+1)*I/O)) per axis
// (bounds in float arithmetic like ATen's start_index / end_index), summed in (d, h, w) order in fp32
// and divided by kd, kh, kw one after the other — ATen's CPU kernel op for op, so the result is
// bit-identical to the reference's.  Nearest (ATen "nearest", legacy): source index
// min(floor(o * (float)I / O), I - 1) per axis.
// One thread per output voxel, lanes along the contiguous axis; HBM-bound (4 B per input voxel of the
// windows, which tile the input exactly when I >= O, + 4 B per output voxel); the overlapping reads of
// neighbouring windows are served by L1 / L2.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;

__device__ __forceinline__ int rs_start(int o, int O, int I) {
  return static_cast<int>(floorf(__fdiv_rn(static_cast<float>(o * I), static_cast<float>(O))));
}
__device__ __forceinline__ int rs_end(int o, int O, int I) {
  return static_cast<int>(ceilf(__fdiv_rn(static_cast<float>((o + 1) * I), static_cast<float>(O))));
}
__device__ __forceinline__ int rs_nearest(int o, int O, int I) {
  if (O == I) return o;
  if (O == 2 * I) return o >> 1;
  const float scale = __fdiv_rn(static_cast<float>(I), static_cast<float>(O));
  return min(static_cast<int>(floorf(__fmul_rn(static_cast<float>(o), scale))), I - 1);
}

// src: contiguous [I0, I1, I2] fp32 per volume (in_shapes: 3 int32 per volume), dst: contiguous [O0, O1, O2].
// Every block first tabulates the window bounds (area) or source indices (nearest) of the three output
// axes for its volume in shared memory — O0 + O1 + O2 entries instead of six IEEE divisions per voxel —
// then walks its voxels with 32-bit index arithmetic.
template <bool AREA>
__global__ void __launch_bounds__(RS_THREADS)
rs_resize(const float* const* __restrict__ srcs, const int32_t* __restrict__ in_shapes, float* const* __restrict__ dsts,
          int O0, int O1, int O2) {
  extern __shared__ int2 rs_tab[];   // [O0 | O1 | O2] -> {start, end} (area) or {index, -} (nearest)
  const int v = blockIdx.y;
  const float* __restrict__ src = srcs[v];
  float* __restrict__ dst = dsts[v];
  const int I0 = __ldg(in_shapes + 3 * v), I1 = __ldg(in_shapes + 3 * v + 1), I2 = __ldg(in_shapes + 3 * v + 2);
  for (int t = threadIdx.x; t < O0 + O1 + O2; t += RS_THREADS) {
    const int ax = t < O0 ? 0 : (t < O0 + O1 ? 1 : 2);
    const int o = ax == 0 ? t : (ax == 1 ? t - O0 : t - O0 - O1);
    const int O = ax == 0 ? O0 : (ax == 1 ? O1 : O2), I = ax == 0 ? I0 : (ax == 1 ? I1 : I2);
    rs_tab[t] = AREA ? make_int2(rs_start(o, O, I), rs_end(o, O, I)) : make_int2(rs_nearest(o, O, I), 0);
  }
  __syncthreads();
  const int2* __restrict__ t0 = rs_tab, *t1 = rs_tab + O0, *t2 = rs_tab + O0 + O1;
  const unsigned n = static_cast<unsigned>(O0) * O1 * O2;   // < 2^31 (checked by the host)
  const unsigned nthr = gridDim.x * blockDim.x;
  const unsigned uO2 = O2, uO1 = O1;
  // (od, oh, ow) of the thread's first voxel and of the grid stride, then carried additions: no
  // division inside the loop
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = i / uO2, ow = i - r * uO2, od = r / uO1, oh = r - od * uO1;
  const unsigned sr = nthr / uO2, sw = nthr - sr * uO2, sd = sr / uO1, sh = sr - sd * uO1;
  for (; i < n; i += nthr, ow += sw, oh += sh, od += sd) {
    if (ow >= uO2) { ow -= uO2; ++oh; }
    if (oh >= uO1) { oh -= uO1; ++od; }
    if (AREA) {
      const int2 d = t0[od], h = t1[oh], w = t2[ow];
      const int kd = d.y - d.x, kh = h.y - h.x, kw = w.y - w.x;
      float sum = 0.0f;
      if (kd <= 2 && kh <= 2 && kw <= 2) {
        // the common case (scale factors between 1/2 and 2): all taps in flight at once, then ATen's
        // summation order (d, h, w); absent taps are not added
        const float* __restrict__ p = src + (static_cast<int64_t>(d.x) * I1 + h.x) * I2 + w.x;
        const int64_t sd = kd > 1 ? static_cast<int64_t>(I1) * I2 : 0, sh = kh > 1 ? I2 : 0;
        const int sw = kw > 1 ? 1 : 0;
        const float v000 = __ldg(p), v001 = __ldg(p + sw), v010 = __ldg(p + sh), v011 = __ldg(p + sh + sw);
        const float v100 = __ldg(p + sd), v101 = __ldg(p + sd + sw), v110 = __ldg(p + sd + sh), v111 = __ldg(p + sd + sh + sw);
        sum = __fadd_rn(sum, v000);
        if (kw > 1) sum = __fadd_rn(sum, v001);
        if (kh > 1) { sum = __fadd_rn(sum, v010); if (kw > 1) sum = __fadd_rn(sum, v011); }
        if (kd > 1) {
          sum = __fadd_rn(sum, v100);
          if (kw > 1) sum = __fadd_rn(sum, v101);
          if (kh > 1) { sum = __fadd_rn(sum, v110); if (kw > 1) sum = __fadd_rn(sum, v111); }
        }
      } else {
        for (int a = d.x; a < d.y; ++a)
          for (int b = h.x; b < h.y; ++b) {
            const float* __restrict__ row = src + (static_cast<int64_t>(a) * I1 + b) * I2;
            for (int c = w.x; c < w.y; ++c) sum = __fadd_rn(sum, __ldg(row + c));
          }
      }
      // sum / kd / kh / kw, one IEEE division after the other (by 1: nothing, by 2: an exact halving)
      float q = sum;
      q = kd == 1 ? q : (kd == 2 ? __fmul_rn(q, 0.5f) : __fdiv_rn(q, static_cast<float>(kd)));
      q = kh == 1 ? q : (kh == 2 ? __fmul_rn(q, 0.5f) : __fdiv_rn(q, static_cast<float>(kh)));
      q = kw == 1 ? q : (kw == 2 ? __fmul_rn(q, 0.5f) : __fdiv_rn(q, static_cast<float>(kw)));
      __stcs(dst + i, q);
    } else {
      __stcs(dst + i, __ldg(src + (static_cast<int64_t>(t0[od].x) * I1 + t1[oh].x) * I2 + t2[ow].x));
    }
  }
}

// ATen's window average of one output voxel, any window size (sequential sum in (d, h, w) order).
__device__ __forceinline__ float rs_area_one(const float* __restrict__ src, int I1, int I2, int2 d, int2 h, int2 w) {
  float sum = 0.0f;
  for (int a = d.x; a < d.y; ++a)
    for (int b = h.x; b < h.y; ++b) {
      const float* __restrict__ row = src + (static_cast<int64_t>(a) * I1 + b) * I2;
      for (int c = w.x; c < w.y; ++c) sum = __fadd_rn(sum, __ldg(row + c));
    }
  return sum;
}
__device__ __forceinline__ float rs_div(float q, int k) {  // q / k: by 1 nothing, by 2 an exact halving
  return k == 1 ? q : (k == 2 ? __fmul_rn(q, 0.5f) : __fdiv_rn(q, static_cast<float>(k)));
}

// Area, O2 a multiple of four: one thread produces four consecutive voxels of an output row.  The
// (d, h) windows and the (at most four) source row pointers are shared by the four, the taps of a voxel
// are in flight together, the result leaves as one 128-bit store.
__global__ void __launch_bounds__(RS_THREADS)
rs_area_quad(const float* const* __restrict__ srcs, const int32_t* __restrict__ in_shapes, float* const* __restrict__ dsts,
             int O0, int O1, int O2) {
  extern __shared__ int2 rs_tab[];
  const int v = blockIdx.y;
  const float* __restrict__ src = srcs[v];
  float* __restrict__ dst = dsts[v];
  const int I0 = __ldg(in_shapes + 3 * v), I1 = __ldg(in_shapes + 3 * v + 1), I2 = __ldg(in_shapes + 3 * v + 2);
  for (int t = threadIdx.x; t < O0 + O1 + O2; t += RS_THREADS) {
    const int ax = t < O0 ? 0 : (t < O0 + O1 ? 1 : 2);
    const int o = ax == 0 ? t : (ax == 1 ? t - O0 : t - O0 - O1);
    const int O = ax == 0 ? O0 : (ax == 1 ? O1 : O2), I = ax == 0 ? I0 : (ax == 1 ? I1 : I2);
    rs_tab[t] = make_int2(rs_start(o, O, I), rs_end(o, O, I));
  }
  __syncthreads();
  const int2* __restrict__ t0 = rs_tab, *t1 = rs_tab + O0, *t2 = rs_tab + O0 + O1;
  const unsigned Q = static_cast<unsigned>(O2) >> 2, uO1 = O1;
  const unsigned n = static_cast<unsigned>(O0) * O1 * Q;
  const unsigned nthr = gridDim.x * blockDim.x;
  const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
  const int64_t plane = static_cast<int64_t>(I1) * I2;
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = i / Q, oq = i - r * Q, od = r / uO1, oh = r - od * uO1;
  const unsigned sr = nthr / Q, sq = nthr - sr * Q, sd = sr / uO1, sh = sr - sd * uO1;
  for (; i < n; i += nthr, oq += sq, oh += sh, od += sd) {
    if (oq >= Q) { oq -= Q; ++oh; }
    if (oh >= uO1) { oh -= uO1; ++od; }
    const int2 d = t0[od], h = t1[oh];
    const int kd = d.y - d.x, kh = h.y - h.x;
    const float* __restrict__ r00 = src + (static_cast<int64_t>(d.x) * I1 + h.x) * I2;
    const float* __restrict__ r01 = r00 + (kh > 1 ? I2 : 0);
    const float* __restrict__ r10 = r00 + (kd > 1 ? plane : 0);
    const float* __restrict__ r11 = r10 + (kh > 1 ? I2 : 0);
    const bool small = kd <= 2 && kh <= 2;
    float q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int2 w = t2[4 * oq + j];
      const int kw = w.y - w.x;
      float sum;
      if (small && kw <= 2) {
        const int c0 = w.x, c1 = w.x + (kw > 1 ? 1 : 0);
        const float v000 = __ldg(r00 + c0), v001 = __ldg(r00 + c1), v010 = __ldg(r01 + c0), v011 = __ldg(r01 + c1);
        const float v100 = __ldg(r10 + c0), v101 = __ldg(r10 + c1), v110 = __ldg(r11 + c0), v111 = __ldg(r11 + c1);
        sum = __fadd_rn(0.0f, v000);
        if (kw > 1) sum = __fadd_rn(sum, v001);
        if (kh > 1) { sum = __fadd_rn(sum, v010); if (kw > 1) sum = __fadd_rn(sum, v011); }
        if (kd > 1) {
          sum = __fadd_rn(sum, v100);
          if (kw > 1) sum = __fadd_rn(sum, v101);
          if (kh > 1) { sum = __fadd_rn(sum, v110); if (kw > 1) sum = __fadd_rn(sum, v111); }
        }
      } else {
        sum = rs_area_one(src, I1, I2, d, h, w);
      }
      q[j] = rs_div(rs_div(rs_div(sum, kd), kh), kw);
    }
    float* __restrict__ o = dst + (static_cast<int64_t>(od) * O1 + oh) * O2 + 4 * oq;
    if (vec) __stcs(reinterpret_cast<float4*>(o), make_float4(q[0], q[1], q[2], q[3]));
    else { o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; }
  }
}

}  // namespace

extern "C" int adell_resize(const float* const* src_dev, const int32_t* in_shapes_dev, float* const* dst_dev, int n_vols,
                            const int32_t* out_shape, int mode, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (src_dev == nullptr || in_shapes_dev == nullptr || dst_dev == nullptr || out_shape == nullptr || n_vols < 0) return ADELL_ERR_BAD_ARG;
  if (mode != ADELL_RESIZE_AREA && mode != ADELL_RESIZE_NEAREST) return ADELL_ERR_UNSUPPORTED;
  const int O0 = out_shape[0], O1 = out_shape[1], O2 = out_shape[2];
  if (O0 <= 0 || O1 <= 0 || O2 <= 0 || n_vols > 65535) return ADELL_ERR_BAD_ARG;
  const int64_t n = static_cast<int64_t>(O0) * O1 * O2;
  if (n > 0x7fffffffLL) return ADELL_ERR_BAD_ARG;
  const size_t tab = sizeof(int2) * (static_cast<size_t>(O0) + O1 + O2);
  if (tab > 48 * 1024) return ADELL_ERR_BAD_ARG;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // a block amortises its tables over >= 16 voxels per thread; whole waves of the SM count
  int64_t blocks = (n + RS_THREADS * 16 - 1) / (RS_THREADS * 16);
  const int64_t wave = (static_cast<int64_t>(sms) * 8 + n_vols - 1) / n_vols;
  if (blocks > wave) blocks = (blocks + wave - 1) / wave * wave;
  if (blocks > 65535) blocks = 65535;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(n_vols));
  if (mode == ADELL_RESIZE_AREA && (O2 & 3) == 0) {
    int64_t qb = (n / 4 + RS_THREADS * 4 - 1) / (RS_THREADS * 4);   // four quads per thread
    if (qb > wave) qb = (qb + wave - 1) / wave * wave;
    if (qb > 65535) qb = 65535;
    if (qb < 1) qb = 1;
    rs_area_quad<<<dim3(static_cast<unsigned>(qb), static_cast<unsigned>(n_vols)), RS_THREADS, tab, static_cast<cudaStream_t>(stream)>>>(
        src_dev, in_shapes_dev, dst_dev, O0, O1, O2);
  } else if (mode == ADELL_RESIZE_AREA)
    rs_resize<true><<<grid, RS_THREADS, tab, static_cast<cudaStream_t>(stream)>>>(src_dev, in_shapes_dev, dst_dev, O0, O1, O2);
  else
    rs_resize<false><<<grid, RS_THREADS, tab, static_cast<cudaStream_t>(stream)>>>(src_dev, in_shapes_dev, dst_dev, O0, O1, O2);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}
