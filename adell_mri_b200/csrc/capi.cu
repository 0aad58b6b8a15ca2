// C-ABI housekeeping entry points (include/adell_b200.h).
#include "common.cuh"

extern "C" int adell_abi_version(void) { return ADELL_ABI_VERSION; }

extern "C" int adell_item_size(void) { return static_cast<int>(sizeof(adell_item)); }

extern "C" const char* adell_status_string(int status) {
  switch (status) {
    case ADELL_OK: return "ok";
    case ADELL_ERR_BAD_ARG: return "bad argument (shape/pointer/count)";
    case ADELL_ERR_DTYPE: return "unsupported source dtype";
    case ADELL_ERR_ALIGN: return "misaligned pointer";
    case ADELL_ERR_LAUNCH: return "CUDA launch failure";
    case ADELL_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
    case ADELL_ERR_NO_DRIVER: return "CUDA driver entry point unavailable";
    case ADELL_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

extern "C" int adell_device_sm_count(int* out) {
  if (out == nullptr) return ADELL_ERR_BAD_ARG;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return ADELL_ERR_NO_DEVICE; }
  e = cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return ADELL_ERR_NO_DEVICE; }
  return ADELL_OK;
}

// Host-only: out[b] = mats[b][0] @ mats[b][1] @ ... @ mats[b][K-1] for 4x4 fp32 matrices, each
// product evaluated as the fp32 FMA chain in k order that torch's CPU mm (MKL sgemm) produces
// for MONAI's `affine @ create_rotate(...)` etc. — so a whole batch of MONAI AffineGrid
// matrices is composed bit-identically to the reference without a Python loop.
extern "C" int adell_mat4_chain(const float* mats, int batch, int k, float* out) {
  if (mats == nullptr || out == nullptr || batch < 0 || k < 1) return ADELL_ERR_BAD_ARG;
  for (int b = 0; b < batch; ++b) {
    float acc[16];
    const float* m0 = mats + (static_cast<size_t>(b) * k) * 16;
    for (int i = 0; i < 16; ++i) acc[i] = m0[i];
    for (int j = 1; j < k; ++j) {
      const float* m = mats + (static_cast<size_t>(b) * k + j) * 16;
      float nxt[16];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          float s = acc[4 * r + 0] * m[0 * 4 + c];
          s = fmaf(acc[4 * r + 1], m[1 * 4 + c], s);
          s = fmaf(acc[4 * r + 2], m[2 * 4 + c], s);
          s = fmaf(acc[4 * r + 3], m[3 * 4 + c], s);
          nxt[4 * r + c] = s;
        }
      for (int i = 0; i < 16; ++i) acc[i] = nxt[i];
    }
    for (int i = 0; i < 16; ++i) out[static_cast<size_t>(b) * 16 + i] = acc[i];
  }
  return ADELL_OK;
}
