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

// Staged (TMA) path descriptor encoding: implemented in gather_staged.cu once that path
// lands; until then no item is eligible and the direct path serves every item.
extern "C" int adell_item_encode_tensormap(adell_item* item_host) {
  if (item_host == nullptr) return ADELL_ERR_BAD_ARG;
  item_host->flags &= static_cast<uint8_t>(~ADELL_F_TMAP);
  return ADELL_ERR_UNSUPPORTED;
}
