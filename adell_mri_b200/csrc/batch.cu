// K4 — batch-level mixing right after collation: (partial) mixup of the collated [B, ...] batch.
//
// Replaces `mixup` / `partial_mixup` of the reference
// (/root/reference/adell_mri/utils/batch_preprocessing.py:31-118, wired in
//  /root/reference/adell_mri/utils/network_factories.py:201-212):
//   x = x * f + x[perm] * (1 - f)            per sample b, f = factor[b]  (fp32: mul, mul, add)
// which torch evaluates as three full passes over the batch plus a gathered copy x[perm].  Here it
// is one pass, HBM-bound: algorithmically 8 B per element (the batch read once, written once; the
// partner x[perm] is another sample of the same batch, served from L2 when the batch fits, else a
// second read: 12 B).  Unselected samples (partial mixup) are copied through.  Out of place: `out` must not
// alias `x` (a permuted partner may itself be a mixed sample).
#include "common.cuh"

namespace {

constexpr int MX_THREADS = 256;

__global__ void __launch_bounds__(MX_THREADS)
mx_mixup(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ factor,
         const int32_t* __restrict__ perm, const uint8_t* __restrict__ sel, int64_t per_sample, int vec) {
  const int b = blockIdx.y;
  const float f = __ldg(factor + b);
  const float g = __fsub_rn(1.0f, f);
  const bool on = sel == nullptr || sel[b] != 0;
  const float* __restrict__ xa = x + static_cast<int64_t>(b) * per_sample;
  const float* __restrict__ xb = x + static_cast<int64_t>(__ldg(perm + b)) * per_sample;
  float* __restrict__ o = out + static_cast<int64_t>(b) * per_sample;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (vec) {
    const float4* a4 = reinterpret_cast<const float4*>(xa);
    const float4* b4 = reinterpret_cast<const float4*>(xb);
    float4* o4 = reinterpret_cast<float4*>(o);
    const int64_t n4 = per_sample >> 2;
    if (on) {
      for (int64_t i = tid; i < n4; i += nthr) {
        const float4 p = __ldcs(a4 + i), q = __ldg(b4 + i);
        float4 r;
        r.x = __fadd_rn(__fmul_rn(p.x, f), __fmul_rn(q.x, g)); r.y = __fadd_rn(__fmul_rn(p.y, f), __fmul_rn(q.y, g));
        r.z = __fadd_rn(__fmul_rn(p.z, f), __fmul_rn(q.z, g)); r.w = __fadd_rn(__fmul_rn(p.w, f), __fmul_rn(q.w, g));
        __stcs(o4 + i, r);
      }
    } else {
      for (int64_t i = tid; i < n4; i += nthr) __stcs(o4 + i, __ldcs(a4 + i));
    }
  } else {
    for (int64_t i = tid; i < per_sample; i += nthr)
      o[i] = on ? __fadd_rn(__fmul_rn(xa[i], f), __fmul_rn(xb[i], g)) : xa[i];
  }
}

}  // namespace

extern "C" int adell_mixup(const float* x_dev, float* out_dev, const float* factor_dev, const int32_t* perm_dev,
                           const uint8_t* sel_dev, int batch, int64_t per_sample, void* stream) {
  if (batch == 0 || per_sample == 0) return ADELL_OK;
  if (x_dev == nullptr || out_dev == nullptr || factor_dev == nullptr || perm_dev == nullptr || batch < 0 || per_sample < 0 ||
      batch > 65535)
    return ADELL_ERR_BAD_ARG;
  if (x_dev == out_dev) return ADELL_ERR_BAD_ARG;
  const int vec = ((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15u) == 0 && (per_sample & 3) == 0;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  // ~8 resident blocks per SM over the whole batch, at least one block per sample
  const int64_t work = vec ? per_sample >> 2 : per_sample;
  int64_t bx = (static_cast<int64_t>(sms) * 8 + batch - 1) / batch;
  const int64_t need = (work + MX_THREADS - 1) / MX_THREADS;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(batch));
  mx_mixup<<<grid, MX_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, out_dev, factor_dev, perm_dev, sel_dev, per_sample, vec);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}
