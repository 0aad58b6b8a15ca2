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

// src: contiguous [I0, I1, I2] fp32 per volume (in_shapes: 3 int32 per volume), dst: contiguous [O0, O1, O2]
template <bool AREA>
__global__ void __launch_bounds__(RS_THREADS)
rs_resize(const float* const* __restrict__ srcs, const int32_t* __restrict__ in_shapes, float* const* __restrict__ dsts,
          int O0, int O1, int O2) {
  const int v = blockIdx.y;
  const float* __restrict__ src = srcs[v];
  float* __restrict__ dst = dsts[v];
  const int I0 = __ldg(in_shapes + 3 * v), I1 = __ldg(in_shapes + 3 * v + 1), I2 = __ldg(in_shapes + 3 * v + 2);
  const int64_t n = static_cast<int64_t>(O0) * O1 * O2;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += nthr) {
    const int ow = static_cast<int>(i % O2);
    const int64_t r = i / O2;
    const int oh = static_cast<int>(r % O1), od = static_cast<int>(r / O1);
    if (AREA) {
      const int d0 = rs_start(od, O0, I0), d1 = rs_end(od, O0, I0);
      const int h0 = rs_start(oh, O1, I1), h1 = rs_end(oh, O1, I1);
      const int w0 = rs_start(ow, O2, I2), w1 = rs_end(ow, O2, I2);
      float sum = 0.0f;
      for (int a = d0; a < d1; ++a)
        for (int b = h0; b < h1; ++b) {
          const float* __restrict__ row = src + (static_cast<int64_t>(a) * I1 + b) * I2;
          for (int c = w0; c < w1; ++c) sum = __fadd_rn(sum, __ldg(row + c));
        }
      const float q = __fdiv_rn(__fdiv_rn(__fdiv_rn(sum, static_cast<float>(d1 - d0)), static_cast<float>(h1 - h0)),
                                static_cast<float>(w1 - w0));
      __stcs(dst + i, q);
    } else {
      const int a = rs_nearest(od, O0, I0), b = rs_nearest(oh, O1, I1), c = rs_nearest(ow, O2, I2);
      __stcs(dst + i, __ldg(src + (static_cast<int64_t>(a) * I1 + b) * I2 + c));
    }
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
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // whole waves of the SM count, about eight voxels per thread
  int64_t blocks = (n + RS_THREADS * 8 - 1) / (RS_THREADS * 8);
  const int64_t wave = (static_cast<int64_t>(sms) * 8 + n_vols - 1) / n_vols;
  if (blocks > wave) blocks = (blocks + wave - 1) / wave * wave;
  if (blocks > 65535 * 16) blocks = 65535 * 16;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(n_vols));
  if (mode == ADELL_RESIZE_AREA)
    rs_resize<true><<<grid, RS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(src_dev, in_shapes_dev, dst_dev, O0, O1, O2);
  else
    rs_resize<false><<<grid, RS_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(src_dev, in_shapes_dev, dst_dev, O0, O1, O2);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}
