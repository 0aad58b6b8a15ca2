// Shared device helpers for the adell_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/adell_b200.h"

static_assert(sizeof(adell_item) == 768, "adell_item must stay 768 bytes (since ABI v6)");

#define ADELL_CUDA_CHECK_LAUNCH()                         \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return adell_map_cuda_error(e__); \
  } while (0)

static inline int adell_map_cuda_error(cudaError_t e) {
  if (e == cudaSuccess) return ADELL_OK;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return ADELL_ERR_NO_DEVICE;
  return ADELL_ERR_LAUNCH;
}

// ---- typed source loads (read-only path) ------------------------------------------------
__device__ __forceinline__ float adell_load_src(const void* __restrict__ base, int64_t idx, int dtype) {
  if (dtype == ADELL_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (dtype == ADELL_I16) return static_cast<float>(__ldg(reinterpret_cast<const short*>(base) + idx));
  return static_cast<float>(__ldg(reinterpret_cast<const unsigned char*>(base) + idx));
}

template <int DT>
__device__ __forceinline__ float adell_load_src_t(const void* __restrict__ base, int64_t idx) {
  if (DT == ADELL_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  if (DT == ADELL_I16) return static_cast<float>(__ldg(reinterpret_cast<const short*>(base) + idx));
  return static_cast<float>(__ldg(reinterpret_cast<const unsigned char*>(base) + idx));
}

// ---- order-preserving keys for the radix statistics ------------------------------------
// fp32: flip all bits of negatives, flip the sign bit of non-negatives => unsigned order ==
// numeric order (-0 < +0 adjacent; NaNs sort above +inf like numpy's sort-last convention).
__device__ __forceinline__ uint32_t adell_key_f32(float v) {
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float adell_unkey_f32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}
#define ADELL_KEY32 3   /* internal: the elements ARE order-preserving keys (candidate lists of adell_quantile_keys) */
__device__ __forceinline__ uint32_t adell_key(const void* __restrict__ base, int64_t idx, int dtype) {
  if (dtype == ADELL_F32) return adell_key_f32(__ldg(reinterpret_cast<const float*>(base) + idx));
  if (dtype == ADELL_KEY32) return __ldg(reinterpret_cast<const uint32_t*>(base) + idx);
  if (dtype == ADELL_I16)
    return (static_cast<uint32_t>(static_cast<int>(__ldg(reinterpret_cast<const short*>(base) + idx)) + 32768)) << 16;
  return static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned char*>(base) + idx)) << 24;
}
__device__ __forceinline__ float adell_unkey(uint32_t k, int dtype) {
  if (dtype == ADELL_F32) return adell_unkey_f32(k);
  if (dtype == ADELL_I16) return static_cast<float>(static_cast<int>(k >> 16) - 32768);
  return static_cast<float>(k >> 24);
}

// ---- Philox4x32-10 (counter-based; Salmon et al. 2011) -----------------------------------
struct adell_philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ adell_philox4 adell_philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = static_cast<uint32_t>(ctr_lo), c1 = static_cast<uint32_t>(ctr_lo >> 32), c2 = ctr_hi, c3 = 0u;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  adell_philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}
// one N(0,1) per counter (Box-Muller on two of the four words)
__device__ __forceinline__ float adell_philox_normal(uint64_t seed, uint64_t ctr) {
  adell_philox4 r = adell_philox4x32_10(seed, ctr, 0u);
  float u1 = (static_cast<float>(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float u2 = (static_cast<float>(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}
