// K2/K3 — intensity statistics on device: min/max, exact radix-histogram order statistics
// (percentiles), and the elementwise intensity program of the reference's scalers.
//
// Replaces the CPU reductions inside monai ScaleIntensityd / ScaleIntensityRangePercentilesd and
// the reference's ConditionalRescalingd / Offsetd
// (/root/reference/adell_mri/transform_factory/transforms.py:143-155,430-443,772-786;
//  /root/reference/adell_mri/utils/monai_transforms/image_intensity_ops.py:71-74,119-121).
// All kernels are HBM-bound streaming passes: 128-bit coalesced loads, warp-shuffle /
// shared-memory privatised reductions, one global atomic per block per result.
#include "common.cuh"

namespace {

constexpr int ST_THREADS = 256;
constexpr int HIST_MAX_BITS = 11;
constexpr int HIST_REPL = 8;  // lane-interleaved copies of the first-pass histogram

inline int st_blocks_per_vol(int64_t max_n, int n_vols, int elems_per_thread) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (max_n + static_cast<int64_t>(ST_THREADS) * elems_per_thread - 1) /
                 (static_cast<int64_t>(ST_THREADS) * elems_per_thread);
  int64_t cap = (static_cast<int64_t>(sms) * 8 + n_vols - 1) / n_vols;  // ~8 resident CTAs per SM overall
  if (cap < 1) cap = 1;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

// Visit every element of a volume as float, 128-bit loads for aligned fp32.
template <typename F>
__device__ __forceinline__ void st_for_each(const adell_vol& v, F&& f) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (v.dtype == ADELL_F32) {
    const float* p = reinterpret_cast<const float*>(v.data);
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
      const int64_t n4 = v.n >> 2;
      const float4* p4 = reinterpret_cast<const float4*>(p);
      for (int64_t i = tid; i < n4; i += nthr) {
        float4 q = __ldg(p4 + i);
        f(q.x, 4 * i); f(q.y, 4 * i + 1); f(q.z, 4 * i + 2); f(q.w, 4 * i + 3);
      }
      for (int64_t i = (n4 << 2) + tid; i < v.n; i += nthr) f(__ldg(p + i), i);
    } else {
      for (int64_t i = tid; i < v.n; i += nthr) f(__ldg(p + i), i);
    }
  } else if (v.dtype == ADELL_I16) {
    const short* p = reinterpret_cast<const short*>(v.data);
    for (int64_t i = tid; i < v.n; i += nthr) f(static_cast<float>(__ldg(p + i)), i);
  } else {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data);
    for (int64_t i = tid; i < v.n; i += nthr) f(static_cast<float>(__ldg(p + i)), i);
  }
}

// ---------------------------------------------------------------- min / max --------------
__global__ void st_minmax_init(uint32_t* out, int n_vols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_vols) {
    out[2 * i + 0] = 0xffffffffu;  // key-space +max
    out[2 * i + 1] = 0u;           // key-space min
  }
}

__global__ void __launch_bounds__(ST_THREADS) st_minmax(const adell_vol* __restrict__ vols, uint32_t* out) {
  const adell_vol v = vols[blockIdx.y];
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  st_for_each(v, [&](float x, int64_t) {
    uint32_t k = adell_key_f32(x);
    kmin = min(kmin, k);
    kmax = max(kmax, k);
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __shared__ uint32_t smin[ST_THREADS / 32], smax[ST_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { smin[warp] = kmin; smax[warp] = kmax; }
  __syncthreads();
  if (warp == 0) {
    kmin = lane < ST_THREADS / 32 ? smin[lane] : 0xffffffffu;
    kmax = lane < ST_THREADS / 32 ? smax[lane] : 0u;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
      kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0 && v.n > 0) {
      atomicMin(out + 2 * blockIdx.y + 0, kmin);
      atomicMax(out + 2 * blockIdx.y + 1, kmax);
    }
  }
}

__global__ void st_minmax_fin(uint32_t* out, int n_vols) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n_vols) reinterpret_cast<float*>(out)[i] = adell_unkey_f32(out[i]);
}

// ---------------------------------------------------------------- mean / std --------------
// NormalizeIntensityd: per-volume mean and population standard deviation.  Accumulated in fp64
// (sum, sum of squares, count) so that the result is the correctly rounded fp32 statistic; with
// nonzero != 0 only elements != 0 take part (MONAI `nonzero=True`).  acc_dev: 3 doubles / volume.
__global__ void __launch_bounds__(ST_THREADS) st_meanstd(const adell_vol* __restrict__ vols, int nonzero, double* acc) {
  const adell_vol v = vols[blockIdx.y];
  double s = 0.0, q = 0.0, c = 0.0;
  st_for_each(v, [&](float x, int64_t) {
    if (!nonzero || x != 0.0f) {
      const double d = static_cast<double>(x);
      s += d; q = fma(d, d, q); c += 1.0;
    }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __shared__ double ss[ST_THREADS / 32], sq[ST_THREADS / 32], sc[ST_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { ss[warp] = s; sq[warp] = q; sc[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < ST_THREADS / 32; ++w) { s += ss[w]; q += sq[w]; c += sc[w]; }
    atomicAdd(acc + 3 * blockIdx.y + 0, s);
    atomicAdd(acc + 3 * blockIdx.y + 1, q);
    atomicAdd(acc + 3 * blockIdx.y + 2, c);
  }
}

// out[2v] = mean, out[2v+1] = std (population; 1 when it is 0, as MONAI substitutes)
__global__ void st_meanstd_fin(const double* __restrict__ acc, int n_vols, float* out, int raw_std) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vols) return;
  const double s = acc[3 * i], q = acc[3 * i + 1], c = acc[3 * i + 2];
  double mean = 0.0, var = 0.0;
  if (c > 0.0) { mean = s / c; var = q / c - mean * mean; }
  if (var < 0.0) var = 0.0;
  float sd = static_cast<float>(sqrt(var));
  if (sd == 0.0f && !raw_std) sd = 1.0f;
  out[2 * i] = static_cast<float>(mean);
  out[2 * i + 1] = sd;
}

// ---------------------------------------------------------------- intensity program -------
// y = ((x*m0 - a)/d)*m1*m2 + b, each op rounded to fp32 (IEEE division).
__device__ __forceinline__ float st_program(float x, const float* __restrict__ c) {
  float y = __fmul_rn(x, c[0]);
  y = __fsub_rn(y, c[1]);
  y = __fdiv_rn(y, c[2]);
  y = __fmul_rn(y, c[3]);
  y = __fmul_rn(y, c[4]);
  y = __fadd_rn(y, c[5]);
  return y;
}

__global__ void __launch_bounds__(ST_THREADS)
st_intensity_map(const adell_vol* __restrict__ vols, float* const* __restrict__ dsts,
                 const float* __restrict__ coefs, int clip, float lo, float hi) {
  const adell_vol v = vols[blockIdx.y];
  float* __restrict__ dst = dsts[blockIdx.y];
  float c[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) c[i] = __ldg(coefs + 6 * blockIdx.y + i);
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool vec = v.dtype == ADELL_F32 && ((reinterpret_cast<uintptr_t>(v.data) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
  if (vec) {
    const float4* p4 = reinterpret_cast<const float4*>(v.data);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int64_t n4 = v.n >> 2;
    for (int64_t i = tid; i < n4; i += nthr) {
      float4 q = __ldg(p4 + i);
      q.x = st_program(q.x, c); q.y = st_program(q.y, c); q.z = st_program(q.z, c); q.w = st_program(q.w, c);
      if (clip) {
        q.x = fminf(hi, fmaxf(q.x, lo)); q.y = fminf(hi, fmaxf(q.y, lo));
        q.z = fminf(hi, fmaxf(q.z, lo)); q.w = fminf(hi, fmaxf(q.w, lo));
      }
      d4[i] = q;
    }
    for (int64_t i = (n4 << 2) + tid; i < v.n; i += nthr) {
      float y = st_program(__ldg(reinterpret_cast<const float*>(v.data) + i), c);
      dst[i] = clip ? fminf(hi, fmaxf(y, lo)) : y;
    }
  } else {
    for (int64_t i = tid; i < v.n; i += nthr) {
      float y = st_program(adell_load_src(v.data, i, v.dtype), c);
      dst[i] = clip ? fminf(hi, fmaxf(y, lo)) : y;
    }
  }
}

// Bounding box of the non-zero voxels of a [S0,S1,S2] volume: out = {lo0, hi0, lo1, hi1, lo2, hi2}
// (hi exclusive; empty mask: lo = INT_MAX, hi = 0).  One read of the mask, warp-shuffle reductions,
// six atomics per warp that saw a non-zero voxel.
__global__ void st_bbox_init(int32_t* out, int n_vols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * n_vols) out[i] = (i & 1) ? 0 : 0x7fffffff;
}
__global__ void __launch_bounds__(ST_THREADS)
st_mask_bbox(const adell_vol* __restrict__ vols, const int32_t* __restrict__ shapes, int32_t* __restrict__ out) {
  const adell_vol v = vols[blockIdx.y];
  const int S1 = shapes[3 * blockIdx.y + 1], S2 = shapes[3 * blockIdx.y + 2];
  const int64_t plane = static_cast<int64_t>(S1) * S2;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {0, 0, 0};
  for (int64_t i = tid; i < v.n; i += nthr) {
    if (adell_load_src(v.data, i, v.dtype) != 0.0f) {
      const int c0 = static_cast<int>(i / plane);
      const int64_t r = i - c0 * plane;
      const int c1 = static_cast<int>(r / S2), c2 = static_cast<int>(r - static_cast<int64_t>(c1) * S2);
      lo[0] = min(lo[0], c0); hi[0] = max(hi[0], c0 + 1);
      lo[1] = min(lo[1], c1); hi[1] = max(hi[1], c1 + 1);
      lo[2] = min(lo[2], c2); hi[2] = max(hi[2], c2 + 1);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0 && hi[0] > 0) {
    int32_t* o = out + 6 * blockIdx.y;
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(o + 2 * a, lo[a]); atomicMax(o + 2 * a + 1, hi[a]); }
  }
}

// Label construction of the cached stage: combine K label maps voxel-wise (any: sum > 0, majority:
// mean > 0.5, none: the single map as is) and map the values (binary: 1 where the value is one of
// `table`, categorical: index of the value in `table`, 0 when absent, none: unchanged).
struct st_label_args {
  const void* src[8];
  int32_t dtype[8];
  float table[16];
  int32_t n_src, combine, op, n_table;
};
__global__ void __launch_bounds__(ST_THREADS)
st_label_map(const st_label_args a, float* __restrict__ dst, int64_t n) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = tid; i < n; i += nthr) {
    float x;
    if (a.combine == ADELL_LABEL_COMBINE_NONE) {
      x = adell_load_src(a.src[0], i, a.dtype[0]);
    } else {
      float sum = 0.0f;  // torch.stack(...).sum(-1) in fp32, key order
      for (int k = 0; k < a.n_src; ++k) sum = __fadd_rn(sum, adell_load_src(a.src[k], i, a.dtype[k]));
      if (a.combine == ADELL_LABEL_COMBINE_ANY) x = sum > 0.0f ? 1.0f : 0.0f;
      else x = __fdiv_rn(sum, static_cast<float>(a.n_src)) > 0.5f ? 1.0f : 0.0f;
    }
    if (a.op != ADELL_LABEL_OP_NONE) {
      float y = 0.0f;
      for (int t = 0; t < a.n_table; ++t)
        if (x == a.table[t]) { y = a.op == ADELL_LABEL_OP_BINARY ? 1.0f : static_cast<float>(t); break; }
      x = y;
    }
    dst[i] = x;
  }
}

// monai AdjustContrast: ((x - min) / (range + eps)) ** gamma * range + min.  Subtraction, division,
// multiplication and addition are fp32 op by op; the power is exp2(gamma * log2(t)) on the SFU
// (t in [0, 1], gamma in (0.5, 4.5]: the absolute error stays below 1e-6 of the range, far inside the
// 1e-4 the intensity maps are held to; powf's ~50 instructions per voxel kept the pass at 0.2 of the
// HBM peak).  t = 0 gives log2 = -inf and exp2 = 0 like pow.
__device__ __forceinline__ float st_gamma_one(float x, float lo, float den, float range, float gamma) {
  const float t = __fdiv_rn(__fsub_rn(x, lo), den);
  const float p = exp2f(__fmul_rn(gamma, __log2f(t)));
  return __fadd_rn(__fmul_rn(p, range), lo);
}
__global__ void __launch_bounds__(ST_THREADS)
st_gamma_map(const adell_vol* __restrict__ vols, float* const* __restrict__ dsts, const float* __restrict__ minmax,
             const float* __restrict__ gammas) {
  const adell_vol v = vols[blockIdx.y];
  float* __restrict__ dst = dsts[blockIdx.y];
  const float lo = __ldg(minmax + 2 * blockIdx.y), hi = __ldg(minmax + 2 * blockIdx.y + 1);
  const float range = __fsub_rn(hi, lo), den = __fadd_rn(range, 1e-7f), gamma = __ldg(gammas + blockIdx.y);
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t done = 0;
  if (v.dtype == ADELL_F32 && ((reinterpret_cast<uintptr_t>(v.data) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0) {
    const float4* __restrict__ src4 = reinterpret_cast<const float4*>(v.data);
    const int64_t n4 = v.n >> 2;
    for (int64_t i = tid; i < n4; i += nthr) {
      const float4 x = __ldcs(src4 + i);
      float4 y;
      y.x = st_gamma_one(x.x, lo, den, range, gamma); y.y = st_gamma_one(x.y, lo, den, range, gamma);
      y.z = st_gamma_one(x.z, lo, den, range, gamma); y.w = st_gamma_one(x.w, lo, den, range, gamma);
      __stcs(reinterpret_cast<float4*>(dst) + i, y);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < v.n; i += nthr) dst[i] = st_gamma_one(adell_load_src(v.data, i, v.dtype), lo, den, range, gamma);
}

// monai RandRicianNoise: sqrt((x + n1)^2 + n2^2), fp32 op by op (torch: add, pow(2) = mul, add, sqrt).
// Streaming: 16 B per voxel; 128-bit accesses when the four pointers allow it.
__global__ void __launch_bounds__(ST_THREADS)
st_rician_map(const float* __restrict__ x, const float* __restrict__ n1, const float* __restrict__ n2,
              float* __restrict__ dst, int64_t n, int vec) {
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto f = [](float v, float a, float b) {
    const float t = __fadd_rn(v, a);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b)));
  };
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthr) {
      const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
      const float4 a = __ldcs(reinterpret_cast<const float4*>(n1) + i);
      const float4 b = __ldcs(reinterpret_cast<const float4*>(n2) + i);
      float4 y;
      y.x = f(v.x, a.x, b.x); y.y = f(v.y, a.y, b.y); y.z = f(v.z, a.z, b.z); y.w = f(v.w, a.w, b.w);
      __stcs(reinterpret_cast<float4*>(dst) + i, y);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthr) dst[i] = f(x[i], n1[i], n2[i]);
}

__global__ void st_scaler_coefs(const float* __restrict__ stats, int n_vols, int scaler, double p0, double p1,
                                float* __restrict__ coefs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vols) return;
  const float lo = stats[2 * i], hi = stats[2 * i + 1];
  float m0 = 1.f, a = 0.f, d = 1.f, m1 = 1.f, m2 = 1.f, b = 0.f;
  const float adc_mult = static_cast<float>(1.0 + (-2.0 / 3.0));  // ScaleIntensity(factor): x*(1+factor)
  if (scaler == ADELL_SCALER_MINMAX) {
    if (lo == hi) {
      m0 = static_cast<float>(p0);  // rescale_array: mina == maxa -> arr * minv
    } else {
      a = lo;
      d = __fsub_rn(hi, lo);
      m1 = static_cast<float>(p1 - p0);
      b = static_cast<float>(p0);
    }
  } else if (scaler == ADELL_SCALER_ADC_SEG) {
    if (hi > static_cast<float>(p0)) m0 = static_cast<float>(p1);
    m1 = adc_mult;
  } else if (scaler == ADELL_SCALER_ADC_CLASS) {
    if (hi > static_cast<float>(p0)) m0 = static_cast<float>(p1);
    a = __fmul_rn(lo, m0);  // Offsetd(None): minus the minimum of the (rescaled) array
    m1 = adc_mult;
  } else if (scaler == ADELL_SCALER_ZSCORE) {
    a = lo;   // mean
    d = hi;   // std (already 1 when the volume is constant)
  } else {  // ADELL_SCALER_RANGE
    a = lo;
    if (__fsub_rn(hi, lo) != 0.0f) {
      d = __fsub_rn(hi, lo);
      m1 = static_cast<float>(p1 - p0);
    }
    b = static_cast<float>(p0);
  }
  float* c = coefs + 6 * i;
  c[0] = m0; c[1] = a; c[2] = d; c[3] = m1; c[4] = m2; c[5] = b;
}

__global__ void st_coefs_to_affine(const float* __restrict__ coefs, int n_vols, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vols) return;
  const float* c = coefs + 6 * i;
  double g = static_cast<double>(c[3]) * c[4] / c[2];
  out[2 * i + 0] = static_cast<float>(c[0] * g);
  out[2 * i + 1] = static_cast<float>(c[5] - c[1] * g);
}

// ---------------------------------------------------------------- radix histogram ---------
// First pass (shift+bits == 32): one histogram per volume, HIST_REPL lane-interleaved copies
// in shared memory so that a warp full of equal keys (background voxels) collides 4-way
// instead of 32-way.  Later passes: only elements whose high bits match a selection's prefix
// count; matching lanes are warp-aggregated with __match_any_sync before the atomic.
__global__ void __launch_bounds__(ST_THREADS)
st_hist_first(const adell_vol* __restrict__ vols, int shared_bins, int bits, unsigned long long* __restrict__ bins,
              const int* __restrict__ gate = nullptr) {
  extern __shared__ uint32_t sh[];  // [nb][HIST_REPL]
  if (gate != nullptr && *gate == 0) return;   // fallback of adell_quantile_keys: runs only when a bracket failed
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < nb * HIST_REPL; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const adell_vol v = vols[blockIdx.y];
  const int shift = 32 - bits;
  const int copy = threadIdx.x & (HIST_REPL - 1);
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  if (v.dtype == ADELL_F32 && (reinterpret_cast<uintptr_t>(v.data) & 15u) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(v.data);
    const int64_t n4 = v.n >> 2;
    auto count4 = [&](const float4 q) {
      uint32_t b0 = adell_key_f32(q.x) >> shift, b1 = adell_key_f32(q.y) >> shift;
      uint32_t b2 = adell_key_f32(q.z) >> shift, b3 = adell_key_f32(q.w) >> shift;
      if (b0 == b1 && b1 == b2 && b2 == b3) {
        atomicAdd(&sh[b0 * HIST_REPL + copy], 4u);
      } else {
        atomicAdd(&sh[b0 * HIST_REPL + copy], 1u); atomicAdd(&sh[b1 * HIST_REPL + copy], 1u);
        atomicAdd(&sh[b2 * HIST_REPL + copy], 1u); atomicAdd(&sh[b3 * HIST_REPL + copy], 1u);
      }
    };
    int64_t i = tid;
    for (; i + 3 * nthr < n4; i += 4 * nthr) {  // four independent 128-bit loads in flight per thread
      const float4 q0 = __ldcs(p4 + i), q1 = __ldcs(p4 + i + nthr), q2 = __ldcs(p4 + i + 2 * nthr), q3 = __ldcs(p4 + i + 3 * nthr);
      count4(q0); count4(q1); count4(q2); count4(q3);
    }
    for (; i < n4; i += nthr) count4(__ldcs(p4 + i));
    const float* p = reinterpret_cast<const float*>(v.data);
    for (int64_t i = (n4 << 2) + tid; i < v.n; i += nthr)
      atomicAdd(&sh[(adell_key_f32(__ldg(p + i)) >> shift) * HIST_REPL + copy], 1u);
  } else {
    for (int64_t i = tid; i < v.n; i += nthr)
      atomicAdd(&sh[(adell_key(v.data, i, v.dtype) >> shift) * HIST_REPL + copy], 1u);
  }
  __syncthreads();
  unsigned long long* out = bins + (shared_bins ? 0 : static_cast<size_t>(blockIdx.y) << bits);
  for (int b = threadIdx.x; b < nb; b += blockDim.x) {
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < HIST_REPL; ++r) s += sh[b * HIST_REPL + r];
    if (s) atomicAdd(out + b, static_cast<unsigned long long>(s));
  }
}

constexpr int NEXT_REPL = 2;  // shared-memory copies of the later passes' histograms

__global__ void __launch_bounds__(ST_THREADS)
st_hist_next(const adell_vol* __restrict__ vols, int n_sel, int shared_bins, const uint32_t* __restrict__ prefix,
             int shift, int bits, unsigned long long* __restrict__ bins, const int* __restrict__ gate = nullptr) {
  extern __shared__ uint32_t sh[];  // [n_uniq][nb]: one histogram per DISTINCT selected prefix
  if (gate != nullptr && *gate == 0) return;
  __shared__ uint32_t spre[8];      // distinct prefixes (high bits)
  __shared__ int s_uniq[8];         // selection -> index of its prefix in spre
  __shared__ int s_nu;
  const int nb = 1 << bits;
  const int h = shared_bins ? 0 : blockIdx.y;
  const int hi_shift = shift + bits;  // < 32 here
  if (threadIdx.x == 0) {
    // the lo / hi order statistics of one percentile (and neighbouring percentiles of flat data)
    // usually share their prefix: such selections see the very same elements, count them once
    int nu = 0;
    for (int s = 0; s < n_sel; ++s) {
      const uint32_t p = __ldg(prefix + h * n_sel + s) >> hi_shift;
      int u = 0;
      while (u < nu && spre[u] != p) ++u;
      if (u == nu) spre[nu++] = p;
      s_uniq[s] = u;
    }
    s_nu = nu;
  }
  __syncthreads();
  const int nu = s_nu;
  for (int i = threadIdx.x; i < nb * nu * NEXT_REPL; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const adell_vol v = vols[blockIdx.y];
  const uint32_t mask = static_cast<uint32_t>(nb - 1);
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int copy = threadIdx.x & (NEXT_REPL - 1);   // lane-interleaved copies: equal keys of a warp collide 32/REPL-way
  // `w` elements with this key under distinct prefix `u`: plain shared-memory atomic (a prefix holds
  // 2^-11 .. 2^-22 of the keys, hits are rare and spread over the bins — no warp votes)
  auto add = [&](int u, uint32_t key, uint32_t w) {
    atomicAdd(&sh[((u << bits) + ((key >> shift) & mask)) * NEXT_REPL + copy], w);
  };
  auto count = [&](uint32_t key, uint32_t w) {   // generic: any dtype, any number of distinct prefixes
    const uint32_t khi = key >> hi_shift;
#pragma unroll 1
    for (int u = 0; u < nu; ++u)
      if (khi == spre[u]) add(u, key, w);
  };
  // fp32 fast test on the RAW bits (no key transform per element): the order-preserving key is
  // u | 0x80000000 for non-negative floats and ~u for negative ones, so "key >> hi_shift == p" is
  // "u >> hi_shift == p ^ top" resp. "u >> hi_shift == ~p" — two registers for the common case of
  // two distinct prefixes (lo / hi statistics of a percentile share theirs), the generic loop otherwise
  const uint32_t top = 1u << (31 - hi_shift), himask = 0xffffffffu >> hi_shift;
  auto raw_of = [&](uint32_t p) { return (p & top) ? (p ^ top) : (~p & himask); };
  const uint32_t r0 = raw_of(spre[0]), r1 = nu > 1 ? raw_of(spre[1]) : r0;
  const bool two = nu <= 2;
  auto count_raw = [&](uint32_t u, uint32_t w) {
    const uint32_t uh = u >> hi_shift;
    if (two) {
      if (uh == r0) add(0, adell_key_f32(__uint_as_float(u)), w);
      else if (uh == r1 && nu > 1) add(1, adell_key_f32(__uint_as_float(u)), w);
    } else {
      count(adell_key_f32(__uint_as_float(u)), w);
    }
  };
  auto count4 = [&](const float4 q) {
    const uint32_t u0 = __float_as_uint(q.x), u1 = __float_as_uint(q.y), u2 = __float_as_uint(q.z), u3 = __float_as_uint(q.w);
    if (u0 == u1 && u1 == u2 && u2 == u3) {  // runs of equal voxels (background): one atomic for the four
      count_raw(u0, 4u);
    } else {
      count_raw(u0, 1u); count_raw(u1, 1u); count_raw(u2, 1u); count_raw(u3, 1u);
    }
  };
  if (v.dtype == ADELL_F32 && (reinterpret_cast<uintptr_t>(v.data) & 15u) == 0) {
    // fp32: four independent 128-bit loads in flight per thread, like the first pass
    const float4* p4 = reinterpret_cast<const float4*>(v.data);
    const int64_t n4 = v.n >> 2;
    int64_t i = tid;
    for (; i + 3 * nthr < n4; i += 4 * nthr) {
      const float4 q0 = __ldcs(p4 + i), q1 = __ldcs(p4 + i + nthr), q2 = __ldcs(p4 + i + 2 * nthr), q3 = __ldcs(p4 + i + 3 * nthr);
      count4(q0); count4(q1); count4(q2); count4(q3);
    }
    for (; i < n4; i += nthr) count4(__ldcs(p4 + i));
    const float* p = reinterpret_cast<const float*>(v.data);
    for (int64_t j = (n4 << 2) + tid; j < v.n; j += nthr) count(adell_key_f32(__ldg(p + j)), 1u);
  } else {
    for (int64_t i = tid; i < v.n; i += nthr) count(adell_key(v.data, i, v.dtype), 1u);
  }
  __syncthreads();
  unsigned long long* out = bins + ((static_cast<size_t>(h) * n_sel) << bits);
  for (int b = threadIdx.x; b < nb * n_sel; b += blockDim.x) {
    const int s_ = b >> bits, bin = b & (nb - 1);
    uint32_t c = 0;
#pragma unroll
    for (int r = 0; r < NEXT_REPL; ++r) c += sh[((s_uniq[s_] << bits) + bin) * NEXT_REPL + r];
    if (c) atomicAdd(out + b, static_cast<unsigned long long>(c));
  }
}

// One block per (histogram, selection): find the bin that holds the requested rank.
__global__ void __launch_bounds__(ST_THREADS)
st_hist_select(const unsigned long long* __restrict__ bins, int n_sel, int first_pass, int shift, int bits,
               uint32_t* __restrict__ prefix, unsigned long long* __restrict__ rank, const int* __restrict__ gate = nullptr) {
  if (gate != nullptr && *gate == 0) return;
  const int h = blockIdx.x / n_sel, s = blockIdx.x % n_sel;
  const int nb = 1 << bits;
  const unsigned long long* hb = bins + ((first_pass ? static_cast<size_t>(h) : static_cast<size_t>(h) * n_sel + s) << bits);
  const int per = (nb + ST_THREADS - 1) / ST_THREADS;  // bins per thread (contiguous)
  const int b_lo = threadIdx.x * per;
  unsigned long long local = 0;
  for (int b = b_lo; b < min(nb, b_lo + per); ++b) local += hb[b];
  __shared__ unsigned long long scan[ST_THREADS];
  scan[threadIdx.x] = local;
  // every thread reads the requested rank (and the prefix so far) BEFORE the barriers of the scan: the one
  // selecting thread overwrites both at the end, and a late warp must not see the already-reduced rank
  const unsigned long long r = rank[blockIdx.x];
  const uint32_t p_in = first_pass ? 0u : prefix[blockIdx.x];
  __syncthreads();
  // inclusive Hillis-Steele scan over 256 partial sums
  for (int o = 1; o < ST_THREADS; o <<= 1) {
    unsigned long long t = threadIdx.x >= o ? scan[threadIdx.x - o] : 0ull;
    __syncthreads();
    scan[threadIdx.x] += t;
    __syncthreads();
  }
  const unsigned long long before = scan[threadIdx.x] - local;
  if (r >= before && r < scan[threadIdx.x]) {  // exactly one thread
    unsigned long long cum = before;
    for (int b = b_lo; b < min(nb, b_lo + per); ++b) {
      unsigned long long c = hb[b];
      if (r < cum + c) {
        prefix[blockIdx.x] = p_in | (static_cast<uint32_t>(b) << shift);
        rank[blockIdx.x] = r - cum;
        break;
      }
      cum += c;
    }
  }
}

__global__ void st_percentile_finalize(const uint32_t* __restrict__ keys, const double* __restrict__ frac, int n,
                                       int dtype, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = adell_unkey(keys[2 * i], dtype), b = adell_unkey(keys[2 * i + 1], dtype);
  const double t = frac[i];
  // numpy _lerp on float32 samples with a float64 weight
  const float diff = __fsub_rn(b, a);
  double r = __dadd_rn(static_cast<double>(a), __dmul_rn(static_cast<double>(diff), t));
  if (t >= 0.5) r = __dsub_rn(static_cast<double>(b), __dmul_rn(static_cast<double>(diff), __dsub_rn(1.0, t)));
  out[i] = static_cast<float>(r);
}

// ---------------------------------------------------------------- one-read order statistics -------
// adell_quantile_keys: the radix passes above read every volume three times (fp32).  Here a strided SAMPLE of the
// volume (<= 32 Ki keys, ~3 % of its sectors) brackets each requested order statistic between two sample order
// statistics 6 sigma apart; ONE full read then counts the keys outside the bracket, the keys equal to its two ends
// (so that the heavy ties of a background level never have to be stored) and appends the few keys strictly inside
// (~0.5 % of the volume at the 0.5th / 99.5th percentile) to a candidate list; the order statistic is selected
// exactly among those.  A bracket that missed (or overflowed its list) raises a flag that un-gates the three
// radix passes: the result is exact either way.
constexpr int QS_THREADS = 1024;
constexpr int QS_SAMPLE = 32768;   // sample keys per volume (128 KiB of shared memory)
constexpr int QMAX = 4;            // quantiles per call
constexpr int QS_POOL = 64;        // volumes a pooled (dataset-wide) call may hold

struct QBracket {
  uint32_t lo, hi;                    // sample order statistics bracketing the (lo, hi) ranks of one quantile
  uint32_t n_cand, pad_;              // keys strictly inside (may exceed the list's capacity: overflow)
  unsigned long long out, ne_lo, ne_hi;   // keys beyond the bracket on the counted side, keys equal to its ends
};

__device__ __forceinline__ uint32_t st_key_at(const adell_vol& v, int64_t i) { return adell_key(v.data, i, v.dtype); }

// Exact radix selection of n_sel ranks among `m` keys that all lie in [klo, khi], by one block.  Only the bits below
// the common prefix of klo and khi vary: digits of 8 bits are taken from there downwards, so the keys spread over the
// 256 bins of every pass (selecting on fixed byte positions sent every key of a narrow bracket to ONE bin for the first
// passes: its shared-memory atomics serialised).  hist: [n_sel][256] shared words; on return sel_key[s] holds the key
// of rank sel_rank[s] (ranks are consumed).
__device__ void st_block_select(const uint32_t* __restrict__ keys, int m, int n_sel, uint32_t klo, uint32_t khi,
                                unsigned long long* sel_rank, uint32_t* sel_key, uint32_t* hist) {
  const int vary = 32 - __clz(static_cast<int>(klo ^ khi));   // number of low bits that differ inside the bracket (0..32)
  const uint32_t common = vary >= 32 ? 0u : (klo >> vary) << vary;
  for (int s = threadIdx.x; s < n_sel; s += blockDim.x) sel_key[s] = common;
  __syncthreads();
  for (int top = vary; top > 0; top -= 8) {
    const int shift = top > 8 ? top - 8 : 0, bits = top - shift;
    const uint32_t mask = (1u << bits) - 1u;
    const bool first = top == vary;   // every key shares the bits above `vary`: one histogram serves all selections
    for (int i = threadIdx.x; i < (first ? 1 : n_sel) * 256; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
      const uint32_t k = keys[i];
      if (first) { atomicAdd(&hist[(k >> shift) & mask], 1u); continue; }
      for (int s = 0; s < n_sel; ++s)
        if ((k >> top) == (sel_key[s] >> top)) atomicAdd(&hist[s * 256 + ((k >> shift) & mask)], 1u);
    }
    __syncthreads();
    if ((threadIdx.x >> 5) < n_sel) {   // warp s places selection s: 8 bins per lane, warp scan, then the lane's own bins
      const int s = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const uint32_t* hs = hist + (first ? 0 : s * 256);
      uint32_t part = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) part += hs[8 * lane + b];   // (bins beyond the digit's width are zero)
      uint32_t incl = part;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const unsigned long long r = sel_rank[s];
      const unsigned hit = __ballot_sync(0xffffffffu, r < incl);   // lanes whose inclusive count exceeds the rank
      const int owner = hit ? __ffs(hit) - 1 : 31;
      if (lane == owner) {
        unsigned long long cum = incl - part;
        uint32_t b = 8 * lane;
        const uint32_t last = min(8u * lane + 7u, mask);
        for (; b < last; ++b) {
          const unsigned long long c = hs[b];
          if (r < cum + c) break;
          cum += c;
        }
        sel_key[s] |= b << shift;
        sel_rank[s] = r - cum;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(QS_THREADS)
st_quantile_sample(const adell_vol* __restrict__ vols, int n_vols, int shared, long long total_n, int n_q,
                   const unsigned long long* __restrict__ rank, QBracket* __restrict__ br, int* __restrict__ any_fail) {
  extern __shared__ uint32_t qsm[];           // [QS_SAMPLE] keys, then [8][256] histogram words
  __shared__ unsigned long long s_rank[8];
  __shared__ uint32_t s_key[8], s_mn[32], s_mx[32];
  __shared__ int s_open[8];
  __shared__ int s_first[QS_POOL + 1];        // pooled mode: first sample group of every volume
  uint32_t* hist = qsm + QS_SAMPLE;
  if (blockIdx.x == 0 && threadIdx.x == 0) *any_fail = 0;
  int m;
  long long n_all;
  auto load_groups = [&](const adell_vol& v, int g0, int ng) {
    // ng groups of four consecutive elements (one 16-byte load of fp32), evenly strided over the volume
    const int64_t groups = v.n >> 2, sg = groups / ng;
    const bool vec = v.dtype == ADELL_F32 && (reinterpret_cast<uintptr_t>(v.data) & 15u) == 0;
    for (int g = threadIdx.x; g < ng; g += blockDim.x) {
      const int64_t e = static_cast<int64_t>(g) * sg;
      if (vec) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(v.data) + e);
        qsm[4 * (g0 + g)] = adell_key_f32(q.x); qsm[4 * (g0 + g) + 1] = adell_key_f32(q.y);
        qsm[4 * (g0 + g) + 2] = adell_key_f32(q.z); qsm[4 * (g0 + g) + 3] = adell_key_f32(q.w);
      } else {
        for (int c = 0; c < 4; ++c) qsm[4 * (g0 + g) + c] = st_key_at(v, 4 * e + c);
      }
    }
  };
  if (!shared) {
    const adell_vol v = vols[blockIdx.x];
    n_all = v.n;
    if (v.n <= QS_SAMPLE) {
      m = static_cast<int>(v.n);
      for (int i = threadIdx.x; i < m; i += blockDim.x) qsm[i] = st_key_at(v, i);
    } else {
      m = QS_SAMPLE;
      load_groups(v, 0, QS_SAMPLE / 4);
    }
  } else {
    // one sample of the POOLED data: every volume contributes groups in proportion to its size (the host sends only
    // pools of at most QS_POOL volumes, each of at least 4 * QS_SAMPLE elements, this way)
    n_all = total_n;
    if (threadIdx.x == 0) {
      long long acc = 0;
      int g = 0;
      for (int v = 0; v < n_vols; ++v) {
        s_first[v] = g;
        acc += vols[v].n;
        g = static_cast<int>((static_cast<long long>(QS_SAMPLE / 4) * acc) / total_n);
      }
      s_first[n_vols] = QS_SAMPLE / 4;
    }
    __syncthreads();
    m = QS_SAMPLE;
    for (int v = 0; v < n_vols; ++v) {
      const int ng = s_first[v + 1] - s_first[v];
      if (ng > 0) load_groups(vols[v], s_first[v], ng);
    }
  }
  __syncthreads();
  // range of the sample (the digits of the selection are taken below the common prefix of its extremes)
  uint32_t mn = 0xffffffffu, mx = 0u;
  for (int i = threadIdx.x; i < m; i += blockDim.x) { const uint32_t k = qsm[i]; mn = min(mn, k); mx = max(mx, k); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x < 32) {
    mn = s_mn[threadIdx.x]; mx = s_mx[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (threadIdx.x == 0) { s_mn[0] = mn; s_mx[0] = mx; }
  }
  if (threadIdx.x < n_q) {
    // bracket of quantile j in sample ranks: position of its lo rank, 6 standard deviations of the binomial either way
    const int j = threadIdx.x;
    const double R = static_cast<double>(rank[(static_cast<size_t>(blockIdx.x) * n_q + j) * 2]);
    const double p = n_all > 0 ? R * m / static_cast<double>(n_all) : 0.0;
    const double var = p * (1.0 - p / static_cast<double>(m > 0 ? m : 1));
    const double sd = var > 0.0 ? sqrt(var) : 0.0;
    const double d = 6.0 * sd + 8.0;
    const double lo = floor(p - d), hi = ceil(p + d) + 1.0;
    s_open[2 * j] = lo < 0.0; s_open[2 * j + 1] = hi >= m;
    s_rank[2 * j] = lo < 0.0 ? 0ull : static_cast<unsigned long long>(lo);
    s_rank[2 * j + 1] = hi >= m ? static_cast<unsigned long long>(m > 0 ? m - 1 : 0) : static_cast<unsigned long long>(hi);
  }
  __syncthreads();
  if (m > 0) st_block_select(qsm, m, 2 * n_q, s_mn[0], s_mx[0], s_rank, s_key, hist);
  __syncthreads();
  if (threadIdx.x < n_q) {
    const int j = threadIdx.x;
    QBracket b;
    b.lo = (s_open[2 * j] || m == 0) ? 0u : s_key[2 * j];
    b.hi = (s_open[2 * j + 1] || m == 0) ? 0xffffffffu : s_key[2 * j + 1];
    b.n_cand = 0u; b.pad_ = 0u; b.out = 0ull; b.ne_lo = 0ull; b.ne_hi = 0ull;
    br[static_cast<size_t>(blockIdx.x) * n_q + j] = b;
  }
}

// A caller that keeps its workspace may reuse the brackets of an earlier call on the same volumes (a device-resident
// cache hands the same volumes back every epoch, and a bracket is only a hint: one that misses un-gates the fallback):
// this replaces the sample kernel by clearing the counters.
__global__ void st_quantile_reset(QBracket* __restrict__ br, int n, int* __restrict__ any_fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *any_fail = 0;
  if (i < n) { br[i].n_cand = 0u; br[i].out = 0ull; br[i].ne_lo = 0ull; br[i].ne_hi = 0ull; }
}

// the counted side of a bracket: keys BELOW it for quantiles in the lower half, keys ABOVE it otherwise (the common
// element, far from every bracket, then touches no counter at all)
__device__ __forceinline__ bool st_low_side(unsigned long long rank_lo, int64_t n) { return 2ull * rank_lo < static_cast<unsigned long long>(n); }

constexpr int QSTAGE = 128;   // candidate keys a WARP stages per quantile before one global append
constexpr int QQUEUE = 512;   // keys of one iteration (16 per lane) a warp can defer to its slow loop

template <int NQ>
__global__ void __launch_bounds__(ST_THREADS)
st_quantile_main(const adell_vol* __restrict__ vols, int shared, long long total_n, const unsigned long long* __restrict__ rank,
                 QBracket* __restrict__ br_all, uint32_t* __restrict__ cand_all, int64_t cap) {
  // The streaming loop is bound by instruction issue, not by HBM (ncu: 28 thread instructions per element and 21 of
  // 32 lanes active when every key was classified where it was loaded).  So the loop only TESTS each key (two
  // instructions for the order-preserving key, one unsigned range test per quantile, one warp vote) and defers the
  // ~1.5 % that concern a bracket to a per-warp queue in shared memory, written without divergence (ballot +
  // prefix popcount); the queue is drained once per iteration with all lanes busy.  Candidates are staged per warp
  // and appended to the volume's list with ONE global atomic per flush (an atomic per candidate serialises on the
  // list's counter: 2 ms per call for a 512x512x128 volume).
  __shared__ uint32_t s_buf[ST_THREADS / 32][NQ][QSTAGE];
  __shared__ uint32_t s_queue[ST_THREADS / 32][QQUEUE];
  __shared__ uint32_t s_cnt[ST_THREADS / 32][NQ], s_qn[ST_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const adell_vol v = vols[blockIdx.y];
  const size_t h = shared ? 0 : blockIdx.y;            // pooled statistics: every volume counts into bracket set 0
  QBracket* __restrict__ br = br_all + h * NQ;
  uint32_t* __restrict__ cand = cand_all + h * NQ * cap;
  rank += h * NQ * 2;
  uint32_t lo[NQ], hi[NQ], c_out[NQ], c_lo[NQ], c_hi[NQ], ia[NQ], iw[NQ];
  bool low[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const QBracket& b = br[j];
    lo[j] = b.lo; hi[j] = b.hi;
    low[j] = st_low_side(rank[j * 2], shared ? total_n : v.n);
    c_out[j] = c_lo[j] = c_hi[j] = 0u;
    // the keys that concern bracket j: its counted side up to the far end of the bracket
    ia[j] = low[j] ? 0u : lo[j]; iw[j] = low[j] ? hi[j] : 0xffffffffu - lo[j];
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {   // keep the two test constants in registers (the compiler otherwise re-derives them
    asm volatile("" : "+r"(ia[j]), "+r"(iw[j]));   // from the 64-bit rank / size comparison at every use)
  }
  if (lane < NQ) s_cnt[warp][lane] = 0u;
  if (lane == 0) s_qn[warp] = 0u;
  __syncwarp();
  uint32_t* queue = s_queue[warp];
  auto visit = [&](uint32_t k) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < NQ; ++j) any |= (k - ia[j]) <= iw[j];
    if (any) queue[atomicAdd(&s_qn[warp], 1u)] = k;   // rare (~1.5 % of the keys)
  };
  auto classify = [&](uint32_t k) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      if (low[j] ? (k > hi[j]) : (k < lo[j])) continue;      // the uncounted side: nothing to do
      if (k < lo[j] || k > hi[j]) { ++c_out[j]; continue; }   // beyond the bracket on the counted side
      if (k == lo[j]) { ++c_lo[j]; continue; }                // (lo == hi: every tie is counted here)
      if (k == hi[j]) { ++c_hi[j]; continue; }
      const uint32_t pos = atomicAdd(&s_cnt[warp][j], 1u);
      if (pos < QSTAGE) {
        s_buf[warp][j][pos] = k;
      } else {   // staging area full before the next flush point (a wide bracket): straight to the list
        const uint32_t g = atomicAdd(&br[j].n_cand, 1u);
        if (g < cap) cand[j * cap + g] = k;
      }
    }
  };
  // called by every lane of the warp, once per iteration: drain the queue, then append full staging areas
  auto drain = [&](bool force) {
    __syncwarp();
    const uint32_t qn = s_qn[warp];
    for (uint32_t i = lane; i < qn; i += 32) classify(queue[i]);
    __syncwarp();
    if (lane == 0) s_qn[warp] = 0u;
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const uint32_t c = min(s_cnt[warp][j], static_cast<uint32_t>(QSTAGE));
      if (!(force ? c > 0u : c > QSTAGE / 2)) continue;      // warp-uniform
      uint32_t base = 0u;
      if (lane == 0) base = atomicAdd(&br[j].n_cand, c);
      base = __shfl_sync(0xffffffffu, base, 0);
      for (uint32_t i = lane; i < c; i += 32)
        if (base + i < cap) cand[j * cap + base + i] = s_buf[warp][j][i];
      __syncwarp();
      if (lane == 0) s_cnt[warp][j] = 0u;
    }
    __syncwarp();
  };
  // order-preserving key of raw fp32 bits in two instructions: u ^ (sign ? 0xffffffff : 0x80000000)
  auto fkey = [](float f) { const uint32_t u = __float_as_uint(f); return u ^ (static_cast<uint32_t>(static_cast<int32_t>(u) >> 31) | 0x80000000u); };
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // every lane of a warp runs the same number of iterations and votes in each visit (lanes past the end vote "no")
  if (v.dtype == ADELL_F32 && (reinterpret_cast<uintptr_t>(v.data) & 15u) == 0) {
    const float4* p4 = reinterpret_cast<const float4*>(v.data);
    const int64_t n4 = v.n >> 2;
    const int64_t full = n4 / (4 * nthr);   // iterations in which every thread of the grid has four loads
    for (int64_t it = 0; it < full; ++it) {
      const float4* p = p4 + tid + it * 4 * nthr;
      const float4 q0 = __ldcs(p), q1 = __ldcs(p + nthr), q2 = __ldcs(p + 2 * nthr), q3 = __ldcs(p + 3 * nthr);   // four independent 128-bit loads in flight
      visit(fkey(q0.x)); visit(fkey(q0.y)); visit(fkey(q0.z)); visit(fkey(q0.w));
      visit(fkey(q1.x)); visit(fkey(q1.y)); visit(fkey(q1.z)); visit(fkey(q1.w));
      visit(fkey(q2.x)); visit(fkey(q2.y)); visit(fkey(q2.z)); visit(fkey(q2.w));
      visit(fkey(q3.x)); visit(fkey(q3.y)); visit(fkey(q3.z)); visit(fkey(q3.w));
      drain(false);
    }
    for (int64_t i = full * 4 * nthr + tid; i < n4; i += nthr) {   // the ragged rest (< 4 loads per thread): classified directly
      const float4 q = __ldcs(p4 + i);
      classify(fkey(q.x)); classify(fkey(q.y)); classify(fkey(q.z)); classify(fkey(q.w));
    }
    for (int64_t j0 = (n4 << 2); j0 < v.n; j0 += nthr) {   // (at most three elements of the whole volume)
      const int64_t j = j0 + tid;
      if (j < v.n) classify(st_key_at(v, j));
    }
  } else {
    const int64_t full = v.n / (8 * nthr);
    for (int64_t it = 0; it < full; ++it) {
      const int64_t i0 = tid + it * 8 * nthr;
#pragma unroll
      for (int u = 0; u < 8; ++u) visit(st_key_at(v, i0 + u * nthr));
      drain(false);
    }
    for (int64_t i = full * 8 * nthr + tid; i < v.n; i += nthr) classify(st_key_at(v, i));
  }
  drain(true);
  // block totals: warp shuffles, then one atomic per counter per warp that saw anything
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    uint32_t a = c_out[j], b = c_lo[j], c = c_hi[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
      QBracket& q = br[j];
      if (a) atomicAdd(&q.out, static_cast<unsigned long long>(a));
      if (b) atomicAdd(&q.ne_lo, static_cast<unsigned long long>(b));
      if (c) atomicAdd(&q.ne_hi, static_cast<unsigned long long>(c));
    }
  }
}

// One block per (volume, quantile): place the lo / hi ranks among {below, == lo, candidates, == hi} and select.
__global__ void __launch_bounds__(QS_THREADS)
st_quantile_finish(const adell_vol* __restrict__ vols, int shared, long long total_n, int n_q, const unsigned long long* __restrict__ rank,
                   const QBracket* __restrict__ br, const uint32_t* __restrict__ cand, int64_t cap,
                   uint32_t* __restrict__ keys_out, int* __restrict__ any_fail) {
  __shared__ uint32_t hist[2 * 256];
  __shared__ unsigned long long s_rank[2];
  __shared__ uint32_t s_key[2];
  __shared__ int s_kind[2];   // 0 = lo key, 1 = candidate, 2 = hi key, 3 = failed
  const int v = blockIdx.x / n_q;
  const int64_t n = shared ? total_n : vols[v].n;
  const QBracket b = br[blockIdx.x];
  if (threadIdx.x < 2) {
    const unsigned long long r0 = rank[static_cast<size_t>(blockIdx.x) * 2 + threadIdx.x];
    const bool low = st_low_side(rank[static_cast<size_t>(blockIdx.x) * 2], n);
    const unsigned long long nc = b.n_cand;
    const unsigned long long below = low ? b.out : static_cast<unsigned long long>(n) - b.out - b.ne_hi - nc - b.ne_lo;
    int kind = 3;
    unsigned long long r = r0;
    if (n > 0 && nc <= static_cast<unsigned long long>(cap) && r >= below) {
      r -= below;
      if (r < b.ne_lo) kind = 0;
      else {
        r -= b.ne_lo;
        if (r < nc) kind = 1;
        else { r -= nc; if (r < b.ne_hi) kind = 2; }
      }
    }
    s_kind[threadIdx.x] = kind;
    s_rank[threadIdx.x] = kind == 1 ? r : 0ull;
  }
  __syncthreads();
  const int m = static_cast<int>(b.n_cand < static_cast<unsigned long long>(cap) ? b.n_cand : cap);
  if (s_kind[0] == 1 || s_kind[1] == 1) st_block_select(cand + static_cast<size_t>(blockIdx.x) * cap, m, 2, b.lo, b.hi, s_rank, s_key, hist);
  __syncthreads();
  if (threadIdx.x < 2) {
    const int kind = s_kind[threadIdx.x];
    uint32_t k = 0u;
    if (kind == 0) k = b.lo;
    else if (kind == 1) k = s_key[threadIdx.x];
    else if (kind == 2) k = b.hi;
    else if (n > 0) atomicExch(any_fail, 1);
    keys_out[static_cast<size_t>(blockIdx.x) * 2 + threadIdx.x] = k;
  }
}

// Long candidate lists (volumes of tens of millions of elements list hundreds of thousands of keys per bracket) are
// not selected by ONE block: st_quantile_place turns every list into a "volume" of raw keys and the two residual
// ranks, the radix kernels above select among them with the whole grid, st_quantile_merge assembles the keys.
__global__ void st_quantile_place(const adell_vol* __restrict__ vols, int shared, long long total_n, int n_q, int n_sets,
                                  const unsigned long long* __restrict__ rank, const QBracket* __restrict__ br,
                                  uint32_t* __restrict__ cand, int64_t cap, adell_vol* __restrict__ cvol,
                                  unsigned long long* __restrict__ crank, int* __restrict__ kinds, int* __restrict__ any_fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (set, lo / hi rank)
  if (i >= n_sets * 2) return;
  const int set = i >> 1, v = set / n_q;
  const int64_t n = shared ? total_n : vols[v].n;
  const QBracket b = br[set];
  const unsigned long long r0 = rank[i];
  const bool low = st_low_side(rank[static_cast<size_t>(set) * 2], n);
  const unsigned long long nc = b.n_cand;
  const unsigned long long below = low ? b.out : static_cast<unsigned long long>(n) - b.out - b.ne_hi - nc - b.ne_lo;
  int kind = 3;
  unsigned long long r = r0;
  if (n > 0 && nc <= static_cast<unsigned long long>(cap) && r >= below) {
    r -= below;
    if (r < b.ne_lo) kind = 0;
    else {
      r -= b.ne_lo;
      if (r < nc) kind = 1;
      else { r -= nc; if (r < b.ne_hi) kind = 2; }
    }
  }
  kinds[i] = kind;
  crank[i] = kind == 1 ? r : 0ull;
  if (kind == 3 && n > 0) atomicExch(any_fail, 1);
  if ((i & 1) == 0) {
    adell_vol c;
    c.data = cand + static_cast<size_t>(set) * cap;
    c.n = static_cast<int64_t>(nc < static_cast<unsigned long long>(cap) ? nc : static_cast<unsigned long long>(cap));
    c.dtype = ADELL_KEY32; c._pad = 0;
    cvol[set] = c;
  }
}

__global__ void st_quantile_merge(const QBracket* __restrict__ br, const int* __restrict__ kinds, const uint32_t* __restrict__ ckeys,
                                  int n, uint32_t* __restrict__ keys_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kind = kinds[i];
  const QBracket b = br[i >> 1];
  keys_out[i] = kind == 0 ? b.lo : (kind == 1 ? ckeys[i] : (kind == 2 ? b.hi : 0u));
}

// One thread per crop: centre = the voxel the host's draw selected from the sample's foreground / background index
// list, corrected like monai correct_crop_centers (allow_smaller semantics: centre clamped into
// [size/2, shape + 1 - size/2 (as uint16) - 1], the upper bound bumped when it equals the lower one), then the
// start of SpatialCrop(roi_center, roi_size): max(centre - size/2, 0).
__global__ void st_posneg_starts(const adell_posneg* __restrict__ crops, int n, int32_t* __restrict__ starts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const adell_posneg c = crops[i];
  int64_t idx = c.indices[c.pick];
  int centre[3];
  centre[2] = static_cast<int>(idx % c.shape[2]); idx /= c.shape[2];
  centre[1] = static_cast<int>(idx % c.shape[1]); idx /= c.shape[1];
  centre[0] = static_cast<int>(idx);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int sz = c.size[a] < c.shape[a] ? c.size[a] : c.shape[a];
    const int vs = sz / 2;
    // np.subtract(shape + 1, size / 2).astype(uint16): float subtraction, truncation
    int ve = static_cast<int>(static_cast<unsigned short>(static_cast<int>(static_cast<double>(c.shape[a] + 1) - static_cast<double>(sz) / 2.0)));
    if (vs == ve) ve += 1;
    int ctr = centre[a] > vs ? centre[a] : vs;
    ctr = ctr < ve - 1 ? ctr : ve - 1;
    const int st = ctr - sz / 2;
    starts[3 * i + a] = st > 0 ? st : 0;
  }
}

}  // namespace

extern "C" int adell_posneg_starts(const adell_posneg* crops_dev, int n_crops, int32_t* starts_dev, void* stream) {
  if (n_crops == 0) return ADELL_OK;
  if (crops_dev == nullptr || starts_dev == nullptr || n_crops < 0) return ADELL_ERR_BAD_ARG;
  st_posneg_starts<<<(n_crops + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(crops_dev, n_crops, starts_dev);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_minmax(const adell_vol* vols_dev, int n_vols, int64_t max_n, float* out_dev, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || out_dev == nullptr || n_vols < 0 || max_n < 0) return ADELL_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint32_t* o = reinterpret_cast<uint32_t*>(out_dev);
  st_minmax_init<<<(n_vols + 127) / 128, 128, 0, st>>>(o, n_vols);
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 32), n_vols);
  st_minmax<<<grid, ST_THREADS, 0, st>>>(vols_dev, o);
  st_minmax_fin<<<(2 * n_vols + 127) / 128, 128, 0, st>>>(o, n_vols);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_meanstd(const adell_vol* vols_dev, int n_vols, int64_t max_n, int nonzero, double* acc_dev,
                             float* out_dev, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || out_dev == nullptr || acc_dev == nullptr || n_vols < 0 || max_n < 0) return ADELL_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(acc_dev, 0, sizeof(double) * 3 * static_cast<size_t>(n_vols), st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 32), n_vols);
  st_meanstd<<<grid, ST_THREADS, 0, st>>>(vols_dev, nonzero & ADELL_MEANSTD_NONZERO, acc_dev);
  st_meanstd_fin<<<(n_vols + 127) / 128, 128, 0, st>>>(acc_dev, n_vols, out_dev, (nonzero & ADELL_MEANSTD_RAW_STD) != 0);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_intensity_map(const adell_vol* vols_dev, float* const* dst_dev, const float* coef_dev,
                                   int n_vols, int64_t max_n, int clip, float clip_lo, float clip_hi,
                                   void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || dst_dev == nullptr || coef_dev == nullptr || n_vols < 0 || max_n < 0)
    return ADELL_ERR_BAD_ARG;
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 16), n_vols);
  st_intensity_map<<<grid, ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(vols_dev, dst_dev, coef_dev, clip,
                                                                                clip_lo, clip_hi);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_label_map(const void* const* src_dev, const int32_t* dtypes, int n_src, int combine, int op,
                               const float* table, int n_table, float* dst_dev, int64_t n, void* stream) {
  if (n == 0) return ADELL_OK;
  if (src_dev == nullptr || dtypes == nullptr || dst_dev == nullptr || n < 0 || n_src < 1 || n_src > 8 || n_table < 0 ||
      n_table > 16 || combine < 0 || combine > ADELL_LABEL_COMBINE_MAJORITY || op < 0 || op > ADELL_LABEL_OP_CATEGORICAL ||
      (n_table > 0 && table == nullptr) || (combine == ADELL_LABEL_COMBINE_NONE && n_src != 1))
    return ADELL_ERR_BAD_ARG;
  st_label_args a;
  for (int k = 0; k < 8; ++k) {
    a.src[k] = k < n_src ? src_dev[k] : nullptr;
    a.dtype[k] = k < n_src ? dtypes[k] : 0;
    if (k < n_src && (src_dev[k] == nullptr || dtypes[k] < 0 || dtypes[k] > ADELL_U8)) return ADELL_ERR_BAD_ARG;
  }
  for (int t = 0; t < 16; ++t) a.table[t] = t < n_table ? table[t] : 0.0f;
  a.n_src = n_src; a.combine = combine; a.op = op; a.n_table = n_table;
  const int blocks = st_blocks_per_vol(n, 1, 16);
  st_label_map<<<blocks, ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a, dst_dev, n);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_mask_bbox(const adell_vol* vols_dev, const int32_t* shapes_dev, int n_vols, int64_t max_n,
                               int32_t* out_dev, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || shapes_dev == nullptr || out_dev == nullptr || n_vols < 0 || max_n < 0) return ADELL_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  st_bbox_init<<<(6 * n_vols + 127) / 128, 128, 0, st>>>(out_dev, n_vols);
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 32), n_vols);
  st_mask_bbox<<<grid, ST_THREADS, 0, st>>>(vols_dev, shapes_dev, out_dev);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_gamma_map(const adell_vol* vols_dev, float* const* dst_dev, const float* minmax_dev,
                               const float* gamma_dev, int n_vols, int64_t max_n, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || dst_dev == nullptr || minmax_dev == nullptr || gamma_dev == nullptr || n_vols < 0 || max_n < 0)
    return ADELL_ERR_BAD_ARG;
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 16), n_vols);
  st_gamma_map<<<grid, ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(vols_dev, dst_dev, minmax_dev, gamma_dev);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_rician_map(const float* x_dev, const float* noise1_dev, const float* noise2_dev, float* dst_dev,
                                int64_t n, void* stream) {
  if (n == 0) return ADELL_OK;
  if (x_dev == nullptr || noise1_dev == nullptr || noise2_dev == nullptr || dst_dev == nullptr || n < 0) return ADELL_ERR_BAD_ARG;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(noise1_dev) |
                         reinterpret_cast<uintptr_t>(noise2_dev) | reinterpret_cast<uintptr_t>(dst_dev);
  if (bits & 3u) return ADELL_ERR_ALIGN;
  const int vec = (bits & 15u) == 0;
  st_rician_map<<<st_blocks_per_vol(n, 1, vec ? 16 : 4), ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      x_dev, noise1_dev, noise2_dev, dst_dev, n, vec);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_scaler_coefs(const float* stats_dev, int n_vols, int scaler, double p0, double p1,
                                  float* coef_dev, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (stats_dev == nullptr || coef_dev == nullptr || n_vols < 0 || scaler < 0 || scaler > ADELL_SCALER_ZSCORE)
    return ADELL_ERR_BAD_ARG;
  st_scaler_coefs<<<(n_vols + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats_dev, n_vols, scaler, p0,
                                                                                       p1, coef_dev);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_coefs_to_affine(const float* coef_dev, int n_vols, float* pre_dev_out, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (coef_dev == nullptr || pre_dev_out == nullptr || n_vols < 0) return ADELL_ERR_BAD_ARG;
  st_coefs_to_affine<<<(n_vols + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(coef_dev, n_vols,
                                                                                          pre_dev_out);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_hist_pass(const adell_vol* vols_dev, int n_vols, int64_t max_n, int n_sel, int shared,
                               const uint32_t* prefix_dev, int pass_shift, int pass_bits, uint64_t* bins_dev,
                               void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || bins_dev == nullptr || n_vols < 0 || max_n < 0 || pass_bits < 1 ||
      pass_bits > HIST_MAX_BITS || pass_shift < 0 || pass_shift + pass_bits > 32 || n_sel < 1 || n_sel > 8)
    return ADELL_ERR_BAD_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(bins_dev);
  const bool first = pass_shift + pass_bits == 32;
  if (first) {
    dim3 grid(st_blocks_per_vol(max_n, n_vols, 64), n_vols);
    size_t smem = (static_cast<size_t>(1) << pass_bits) * HIST_REPL * sizeof(uint32_t);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(st_hist_first, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    st_hist_first<<<grid, ST_THREADS, smem, st>>>(vols_dev, shared, pass_bits, bins);
  } else {
    if (prefix_dev == nullptr) return ADELL_ERR_BAD_ARG;
    dim3 grid(st_blocks_per_vol(max_n, n_vols, 64), n_vols);
    size_t smem = (static_cast<size_t>(1) << pass_bits) * n_sel * NEXT_REPL * sizeof(uint32_t);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(st_hist_next, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    st_hist_next<<<grid, ST_THREADS, smem, st>>>(vols_dev, n_sel, shared, prefix_dev, pass_shift, pass_bits, bins);
  }
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_hist_select(const uint64_t* bins_dev, int n_hist, int n_sel, int pass_shift, int pass_bits,
                                 uint32_t* prefix_dev, uint64_t* rank_dev, void* stream) {
  if (n_hist == 0) return ADELL_OK;
  if (bins_dev == nullptr || prefix_dev == nullptr || rank_dev == nullptr || n_hist < 0 || n_sel < 1 || n_sel > 8 ||
      pass_bits < 1 || pass_bits > HIST_MAX_BITS || pass_shift < 0 || pass_shift + pass_bits > 32)
    return ADELL_ERR_BAD_ARG;
  const int first = pass_shift + pass_bits == 32;
  st_hist_select<<<n_hist * n_sel, ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(bins_dev), n_sel, first, pass_shift, pass_bits, prefix_dev,
      reinterpret_cast<unsigned long long*>(rank_dev));
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

extern "C" int adell_percentile_finalize(const uint32_t* keys_dev, const double* frac_dev, int n_vols, int n_q,
                                         int dtype, float* out_dev, void* stream) {
  const int n = n_vols * n_q;
  if (n == 0) return ADELL_OK;
  if (keys_dev == nullptr || frac_dev == nullptr || out_dev == nullptr || n < 0 || dtype < 0 || dtype > ADELL_U8)
    return ADELL_ERR_BAD_ARG;
  st_percentile_finalize<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(keys_dev, frac_dev, n, dtype,
                                                                                         out_dev);
  ADELL_CUDA_CHECK_LAUNCH();
  return ADELL_OK;
}

namespace {
struct QLayout {
  int64_t cap, off_br, off_flag, off_cand, off_rank, off_bins, off_cvol, off_crank, off_ckeys, off_kinds, total;
};
// n_hist: sets of statistics (one per volume, or ONE for pooled statistics); span: elements behind one set
QLayout st_quantile_layout(int n_hist, int n_q, int64_t span) {
  QLayout L;
  L.cap = span / 24 > 4096 ? span / 24 : 4096;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  int64_t o = 0;
  L.off_br = o; o += up(static_cast<int64_t>(sizeof(QBracket)) * n_hist * n_q);
  L.off_flag = o; o += 256;
  L.off_cand = o; o += up(4 * L.cap * n_hist * n_q);
  L.off_rank = o; o += up(8ll * n_hist * n_q * 2);
  L.off_bins = o; o += up((8ll * n_hist * n_q * 2 * 2) << HIST_MAX_BITS);   // (fallback: n_hist x 2 n_q; candidate lists: n_hist n_q x 2)
  L.off_cvol = o; o += up(static_cast<int64_t>(sizeof(adell_vol)) * n_hist * n_q);
  L.off_crank = o; o += up(8ll * n_hist * n_q * 2);
  L.off_ckeys = o; o += up(4ll * n_hist * n_q * 2);
  L.off_kinds = o; o += up(4ll * n_hist * n_q * 2);
  L.total = o;
  return L;
}
}  // namespace

extern "C" int64_t adell_quantile_workspace(int n_vols, int n_q, int64_t max_n, int64_t pooled_n) {
  if (n_vols < 0 || n_q < 1 || n_q > QMAX || max_n < 0 || pooled_n < 0) return -1;
  return pooled_n > 0 ? st_quantile_layout(1, n_q, pooled_n).total : st_quantile_layout(n_vols, n_q, max_n).total;
}

extern "C" int adell_quantile_keys(const adell_vol* vols_dev, int n_vols, int64_t max_n, int64_t pooled_n, int dtype, int n_q,
                                   const uint64_t* rank_dev, uint32_t* keys_dev, void* workspace_dev, int64_t workspace_bytes,
                                   int flags, void* stream) {
  if (n_vols == 0) return ADELL_OK;
  if (vols_dev == nullptr || rank_dev == nullptr || keys_dev == nullptr || workspace_dev == nullptr || n_vols < 0 || max_n < 0 ||
      pooled_n < 0 || n_q < 1 || n_q > QMAX || dtype < 0 || dtype > ADELL_U8)
    return ADELL_ERR_BAD_ARG;
  const int shared = pooled_n > 0 ? 1 : 0;
  // pooled statistics: the sample takes groups from every volume in proportion to its size
  if (shared && (n_vols > QS_POOL || pooled_n < 4ll * QS_SAMPLE)) return ADELL_ERR_UNSUPPORTED;
  const int n_hist = shared ? 1 : n_vols;
  const QLayout L = st_quantile_layout(n_hist, n_q, shared ? pooled_n : max_n);
  if (workspace_bytes < L.total) return ADELL_ERR_NO_SPACE;
  if (reinterpret_cast<uintptr_t>(workspace_dev) & 255u) return ADELL_ERR_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  QBracket* br = reinterpret_cast<QBracket*>(ws + L.off_br);
  int* flag = reinterpret_cast<int*>(ws + L.off_flag);
  uint32_t* cand = reinterpret_cast<uint32_t*>(ws + L.off_cand);
  unsigned long long* rank_copy = reinterpret_cast<unsigned long long*>(ws + L.off_rank);
  unsigned long long* bins = reinterpret_cast<unsigned long long*>(ws + L.off_bins);
  const unsigned long long* rank = reinterpret_cast<const unsigned long long*>(rank_dev);
  const long long total_n = static_cast<long long>(pooled_n);
  const size_t smem = (QS_SAMPLE + 8 * 256) * sizeof(uint32_t);
  cudaError_t e = cudaFuncSetAttribute(st_quantile_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  if (flags & ADELL_QUANTILE_REUSE_BRACKETS)
    st_quantile_reset<<<(n_hist * n_q + 127) / 128, 128, 0, st>>>(br, n_hist * n_q, flag);
  else
    st_quantile_sample<<<n_hist, QS_THREADS, smem, st>>>(vols_dev, n_vols, shared, total_n, n_q, rank, br, flag);
  ADELL_CUDA_CHECK_LAUNCH();
  dim3 grid(st_blocks_per_vol(max_n, n_vols, 64), n_vols);
  {
    // one wave: as many blocks as are resident at once (the generic estimate assumes 8 per SM; this kernel's registers
    // allow fewer, and a second, partial wave leaves most SMs idle at the end)
    int dev = 0, sms = 148, occ = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t oe = cudaSuccess;
    switch (n_q) {
      case 1: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, st_quantile_main<1>, ST_THREADS, 0); break;
      case 2: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, st_quantile_main<2>, ST_THREADS, 0); break;
      case 3: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, st_quantile_main<3>, ST_THREADS, 0); break;
      default: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, st_quantile_main<4>, ST_THREADS, 0); break;
    }
    if (oe != cudaSuccess) { (void)cudaGetLastError(); occ = 0; }
    if (occ > 0) {
      const int per_vol = (sms * occ) / n_vols;
      if (per_vol >= 1 && per_vol < static_cast<int>(grid.x)) grid.x = static_cast<unsigned>(per_vol);
    }
  }
  switch (n_q) {
    case 1: st_quantile_main<1><<<grid, ST_THREADS, 0, st>>>(vols_dev, shared, total_n, rank, br, cand, L.cap); break;
    case 2: st_quantile_main<2><<<grid, ST_THREADS, 0, st>>>(vols_dev, shared, total_n, rank, br, cand, L.cap); break;
    case 3: st_quantile_main<3><<<grid, ST_THREADS, 0, st>>>(vols_dev, shared, total_n, rank, br, cand, L.cap); break;
    default: st_quantile_main<4><<<grid, ST_THREADS, 0, st>>>(vols_dev, shared, total_n, rank, br, cand, L.cap); break;
  }
  ADELL_CUDA_CHECK_LAUNCH();
  const int64_t span = shared ? static_cast<int64_t>(pooled_n) : max_n;
  if (span < (1ll << 23)) {
    // short candidate lists (a few ten thousand keys): one block per (set, quantile) selects both ranks
    st_quantile_finish<<<n_hist * n_q, QS_THREADS, 0, st>>>(vols_dev, shared, total_n, n_q, rank, br, cand, L.cap, keys_dev, flag);
    ADELL_CUDA_CHECK_LAUNCH();
  } else {
    // long lists: the radix kernels select among the candidates with the whole grid (three passes over a few MB)
    const int n_sets = n_hist * n_q;
    adell_vol* cvol = reinterpret_cast<adell_vol*>(ws + L.off_cvol);
    unsigned long long* crank = reinterpret_cast<unsigned long long*>(ws + L.off_crank);
    uint32_t* ckeys = reinterpret_cast<uint32_t*>(ws + L.off_ckeys);
    int* kinds = reinterpret_cast<int*>(ws + L.off_kinds);
    st_quantile_place<<<(2 * n_sets + 63) / 64, 64, 0, st>>>(vols_dev, shared, total_n, n_q, n_sets, rank, br, cand, L.cap, cvol, crank,
                                                              kinds, flag);
    ADELL_CUDA_CHECK_LAUNCH();
    static const int kKeySched[3][2] = {{21, 11}, {10, 11}, {0, 10}};
    dim3 cgrid(st_blocks_per_vol(L.cap / 8 + 1, n_sets, 64), n_sets);   // (a list holds a fraction of its capacity)
    for (int p = 0; p < 3; ++p) {
      const int shift = kKeySched[p][0], bits = kKeySched[p][1];
      const bool first = p == 0;
      const size_t nbins = (static_cast<size_t>(n_sets) * (first ? 1 : 2)) << bits;
      e = cudaMemsetAsync(bins, 0, nbins * sizeof(unsigned long long), st);
      if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
      if (first) {
        const size_t sm = (static_cast<size_t>(1) << bits) * HIST_REPL * sizeof(uint32_t);
        if (sm > 48 * 1024) cudaFuncSetAttribute(st_hist_first, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
        st_hist_first<<<cgrid, ST_THREADS, sm, st>>>(cvol, 0, bits, bins, nullptr);
      } else {
        const size_t sm = (static_cast<size_t>(1) << bits) * 2 * NEXT_REPL * sizeof(uint32_t);
        if (sm > 48 * 1024) cudaFuncSetAttribute(st_hist_next, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
        st_hist_next<<<cgrid, ST_THREADS, sm, st>>>(cvol, 2, 0, ckeys, shift, bits, bins, nullptr);
      }
      ADELL_CUDA_CHECK_LAUNCH();
      st_hist_select<<<n_sets * 2, ST_THREADS, 0, st>>>(bins, 2, first ? 1 : 0, shift, bits, ckeys, crank, nullptr);
      ADELL_CUDA_CHECK_LAUNCH();
    }
    st_quantile_merge<<<(2 * n_sets + 63) / 64, 64, 0, st>>>(br, kinds, ckeys, 2 * n_sets, keys_dev);
    ADELL_CUDA_CHECK_LAUNCH();
  }
  // Fallback, gated on the flag (every kernel returns at once when no bracket failed): the three radix passes over
  // all volumes, writing the same keys.
  const int n_sel = 2 * n_q;
  e = cudaMemcpyAsync(rank_copy, rank, sizeof(unsigned long long) * static_cast<size_t>(n_hist) * n_sel, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  static const int kSched[3][3][2] = {{{21, 11}, {10, 11}, {0, 10}}, {{21, 11}, {16, 5}, {-1, 0}}, {{24, 8}, {-1, 0}, {-1, 0}}};
  dim3 hgrid(st_blocks_per_vol(max_n, n_vols, 64), n_vols);
  for (int p = 0; p < 3; ++p) {
    const int shift = kSched[dtype][p][0], bits = kSched[dtype][p][1];
    if (shift < 0) break;
    const bool first = p == 0;
    const size_t nbins = (static_cast<size_t>(n_hist) * (first ? 1 : n_sel)) << bits;
    e = cudaMemsetAsync(bins, 0, nbins * sizeof(unsigned long long), st);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
    if (first) {
      const size_t sm = (static_cast<size_t>(1) << bits) * HIST_REPL * sizeof(uint32_t);
      if (sm > 48 * 1024) cudaFuncSetAttribute(st_hist_first, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
      st_hist_first<<<hgrid, ST_THREADS, sm, st>>>(vols_dev, shared, bits, bins, flag);
    } else {
      const size_t sm = (static_cast<size_t>(1) << bits) * n_sel * NEXT_REPL * sizeof(uint32_t);
      if (sm > 48 * 1024) cudaFuncSetAttribute(st_hist_next, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
      st_hist_next<<<hgrid, ST_THREADS, sm, st>>>(vols_dev, n_sel, shared, keys_dev, shift, bits, bins, flag);
    }
    ADELL_CUDA_CHECK_LAUNCH();
    st_hist_select<<<n_hist * n_sel, ST_THREADS, 0, st>>>(bins, n_sel, first ? 1 : 0, shift, bits, keys_dev, rank_copy, flag);
    ADELL_CUDA_CHECK_LAUNCH();
  }
  return ADELL_OK;
}

// debugging / test aid: 1 when the last adell_quantile_keys on this workspace took the gated fallback (read after a sync)
extern "C" int adell_quantile_fell_back(const void* workspace_dev, int n_vols, int n_q, int64_t max_n, int64_t pooled_n, int* out_host) {
  if (workspace_dev == nullptr || out_host == nullptr) return ADELL_ERR_BAD_ARG;
  const QLayout L = pooled_n > 0 ? st_quantile_layout(1, n_q, pooled_n) : st_quantile_layout(n_vols, n_q, max_n);
  cudaError_t e = cudaMemcpy(out_host, static_cast<const uint8_t*>(workspace_dev) + L.off_flag, sizeof(int), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return adell_map_cuda_error(e); }
  return ADELL_OK;
}
