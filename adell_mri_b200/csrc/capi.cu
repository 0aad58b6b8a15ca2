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
    case ADELL_ERR_NO_SPACE: return "caller-provided buffer too small";
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

// Host-only: the whole MONAI AffineGrid matrix  eye @ Rx @ Ry @ Rz @ shear @ translate @ scale  for a batch, from
// the raw fp32 parameters (sines / cosines of the rotation angles are passed in: they are evaluated by the caller
// with the same fp32 routine MONAI uses, torch.sin / torch.cos), every 4x4 product the fp32 FMA chain of
// adell_mat4_chain.  Replaces six numpy matrix fills + a stack per call of geometry.compose_affine.
extern "C" int adell_affine_compose(const float* sin_r, const float* cos_r, int k_rot, const float* shear, int k_shear,
                                    const float* translate, int k_trans, const float* scale, int k_scale, int batch,
                                    float* out) {
  if (out == nullptr || batch < 0 || k_rot < 0 || k_rot > 3 || k_shear < 0 || k_trans < 0 || k_scale < 0) return ADELL_ERR_BAD_ARG;
  if ((k_rot && (sin_r == nullptr || cos_r == nullptr)) || (k_shear && shear == nullptr) || (k_trans && translate == nullptr) ||
      (k_scale && scale == nullptr))
    return ADELL_ERR_BAD_ARG;
  auto eye = [](float* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; };
  auto mul = [](float* acc, const float* m) {
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
  };
  for (int b = 0; b < batch; ++b) {
    float acc[16], m[16];
    eye(acc);
    for (int j = 0; j < k_rot; ++j) {
      const float s = sin_r[static_cast<size_t>(b) * k_rot + j], c = cos_r[static_cast<size_t>(b) * k_rot + j];
      eye(m);
      if (j == 0) { m[5] = c; m[6] = -s; m[9] = s; m[10] = c; }        // Rx
      else if (j == 1) { m[0] = c; m[2] = s; m[8] = -s; m[10] = c; }   // Ry
      else { m[0] = c; m[1] = -s; m[4] = s; m[5] = c; }                // Rz
      mul(acc, m);
    }
    if (k_shear) {
      float c6[6] = {0, 0, 0, 0, 0, 0};
      for (int j = 0; j < k_shear && j < 6; ++j) c6[j] = shear[static_cast<size_t>(b) * k_shear + j];
      eye(m);
      m[1] = c6[0]; m[2] = c6[1]; m[4] = c6[2]; m[6] = c6[3]; m[8] = c6[4]; m[9] = c6[5];
      mul(acc, m);
    }
    if (k_trans) {
      eye(m);
      for (int j = 0; j < k_trans && j < 3; ++j) m[4 * j + 3] = translate[static_cast<size_t>(b) * k_trans + j];
      mul(acc, m);
    }
    if (k_scale) {
      eye(m);
      for (int j = 0; j < k_scale && j < 3; ++j) m[5 * j] = scale[static_cast<size_t>(b) * k_scale + j];
      mul(acc, m);
    }
    for (int i = 0; i < 16; ++i) out[static_cast<size_t>(b) * 16 + i] = acc[i];
  }
  return ADELL_OK;
}
