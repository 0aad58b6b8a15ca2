// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tex_probe tools/tex_probe.cu
// Stand-alone probe (measurement tooling, not product): can the texture units of a B200 feed a trilinear affine
// resample straight from PITCH-LINEAR global memory?  A volume [D0][D1][D2] is bound as a 2-D texture of height D0 and
// width D1*D2 (cudaResourceTypePitch2D, no copy into a CUDA array); one tex2Dgather returns the 2 x 2 footprint over
// (axis 0, axis 2), so a voxel needs TWO texture instructions instead of eight shared-memory loads.  The probe
// (1) prints the component order of the gather, (2) checks a rotated resample against a host computation and
// (3) times it (CUDA events) next to an 8 x tex2D point-fetch variant and an 8 x __ldg variant.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Map { float m[9]; float t[3]; };

__global__ void order_probe(cudaTextureObject_t tex, float x, float y, float4* out) { *out = tex2Dgather<float4>(tex, x, y, 0); }

template <int MODE>  // 0 gather, 1 eight point fetches, 2 eight __ldg
__global__ void __launch_bounds__(256) resample(const cudaTextureObject_t* texs, const float* __restrict__ src, float* __restrict__ dst,
                                                int D0, int D1, int D2, Map M) {
  const int v = blockIdx.z;
  const cudaTextureObject_t tex = texs[v];
  const float* s = src + (size_t)v * D0 * D1 * D2;
  float* d = dst + (size_t)v * D0 * D1 * D2;
  const int o2 = threadIdx.x, o1 = blockIdx.x * 8 + threadIdx.y;
  const float fD2 = (float)D2;
#pragma unroll 2
  for (int k = 0; k < 8; ++k) {
    const int o0 = blockIdx.y * 8 + k;
    const float a = (float)o0, b = (float)o1, c = (float)o2;
    const float v0 = fmaf(M.m[0], a, fmaf(M.m[1], b, fmaf(M.m[2], c, M.t[0])));
    const float v1 = fmaf(M.m[3], a, fmaf(M.m[4], b, fmaf(M.m[5], c, M.t[1])));
    const float v2 = fmaf(M.m[6], a, fmaf(M.m[7], b, fmaf(M.m[8], c, M.t[2])));
    const float f0 = floorf(v0), f1 = floorf(v1), f2 = floorf(v2);
    const float r0 = v0 - f0, r1 = v1 - f1, r2 = v2 - f2;
    float t[8];  // t[4*a0 + 2*a1 + a2]
    if (MODE == 0) {
      // gather footprint over (axis 0 = y, axis 2 = x); coordinates on the texel corner shared by the four texels
      const float x = fmaf(f1, fD2, f2) + 1.0f, y = f0 + 1.0f;
      const float4 g = tex2Dgather<float4>(tex, x, y, 0);        // row f1
      const float4 h = tex2Dgather<float4>(tex, x + fD2, y, 0);  // row f1 + 1
      // order (checked by order_probe): w = (x0,y0), z = (x1,y0), x = (x0,y1), y = (x1,y1)
      const bool c0ok = f2 >= 0.0f && f2 < fD2, c1ok = f2 >= -1.0f && f2 < fD2 - 1.0f;
      const bool r0ok = f1 >= 0.0f && f1 < (float)D1, r1ok = f1 >= -1.0f && f1 < (float)D1 - 1.0f;
      t[0] = (c0ok && r0ok) ? g.w : 0.0f; t[1] = (c1ok && r0ok) ? g.z : 0.0f;
      t[4] = (c0ok && r0ok) ? g.x : 0.0f; t[5] = (c1ok && r0ok) ? g.y : 0.0f;
      t[2] = (c0ok && r1ok) ? h.w : 0.0f; t[3] = (c1ok && r1ok) ? h.z : 0.0f;
      t[6] = (c0ok && r1ok) ? h.x : 0.0f; t[7] = (c1ok && r1ok) ? h.y : 0.0f;
    } else {
      const int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int j0 = i0 + (q >> 2), j1 = i1 + ((q >> 1) & 1), j2 = i2 + (q & 1);
        const bool ok = j0 >= 0 && j0 < D0 && j1 >= 0 && j1 < D1 && j2 >= 0 && j2 < D2;
        if (MODE == 1) t[q] = ok ? tex2D<float>(tex, (float)(j1 * D2 + j2) + 0.5f, (float)j0 + 0.5f) : 0.0f;
        else t[q] = ok ? __ldg(s + ((size_t)j0 * D1 + j1) * D2 + j2) : 0.0f;
      }
    }
    const float x00 = fmaf(r2, t[1] - t[0], t[0]), x01 = fmaf(r2, t[3] - t[2], t[2]);
    const float x10 = fmaf(r2, t[5] - t[4], t[4]), x11 = fmaf(r2, t[7] - t[6], t[6]);
    const float y0 = fmaf(r1, x01 - x00, x00), y1 = fmaf(r1, x11 - x10, x10);
    d[((size_t)o0 * D1 + o1) * D2 + o2] = fmaf(r0, y1 - y0, y0);
  }
}

static float host_voxel(const std::vector<float>& s, int D0, int D1, int D2, const Map& M, int o0, int o1, int o2) {
  const float a = (float)o0, b = (float)o1, c = (float)o2;
  const float v0 = fmaf(M.m[0], a, fmaf(M.m[1], b, fmaf(M.m[2], c, M.t[0])));
  const float v1 = fmaf(M.m[3], a, fmaf(M.m[4], b, fmaf(M.m[5], c, M.t[1])));
  const float v2 = fmaf(M.m[6], a, fmaf(M.m[7], b, fmaf(M.m[8], c, M.t[2])));
  const float f0 = floorf(v0), f1 = floorf(v1), f2 = floorf(v2);
  const float r0 = v0 - f0, r1 = v1 - f1, r2 = v2 - f2;
  float t[8];
  for (int q = 0; q < 8; ++q) {
    const int j0 = (int)f0 + (q >> 2), j1 = (int)f1 + ((q >> 1) & 1), j2 = (int)f2 + (q & 1);
    const bool ok = j0 >= 0 && j0 < D0 && j1 >= 0 && j1 < D1 && j2 >= 0 && j2 < D2;
    t[q] = ok ? s[((size_t)j0 * D1 + j1) * D2 + j2] : 0.0f;
  }
  const float x00 = fmaf(r2, t[1] - t[0], t[0]), x01 = fmaf(r2, t[3] - t[2], t[2]);
  const float x10 = fmaf(r2, t[5] - t[4], t[4]), x11 = fmaf(r2, t[7] - t[6], t[6]);
  const float y0 = fmaf(r1, x01 - x00, x00), y1 = fmaf(r1, x11 - x10, x10);
  return fmaf(r0, y1 - y0, y0);
}

int main(int argc, char** argv) {
  const int D0 = argc > 1 ? atoi(argv[1]) : 256, D1 = argc > 2 ? atoi(argv[2]) : 256, D2 = argc > 3 ? atoi(argv[3]) : 32;
  const int NV = argc > 4 ? atoi(argv[4]) : 32;
  const double ang[3] = {argc > 5 ? atof(argv[5]) : 0.27, argc > 6 ? atof(argv[6]) : -0.2, argc > 7 ? atof(argv[7]) : 0.15};
  const size_t n = (size_t)D0 * D1 * D2;
  std::vector<float> h(n * NV);
  unsigned s = 12345u;
  for (size_t i = 0; i < h.size(); ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)(s >> 8) * (1.0f / 16777216.0f); }
  float *src, *dst;
  CK(cudaMalloc(&src, h.size() * 4));
  CK(cudaMalloc(&dst, h.size() * 4));
  CK(cudaMemcpy(src, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, SMs %d, texturePitchAlignment %zu, maxTexture2DLinear %d x %d (pitch %d), maxTexture2DGather %d x %d\n", prop.name,
         prop.multiProcessorCount, prop.texturePitchAlignment, prop.maxTexture2DLinear[0], prop.maxTexture2DLinear[1], prop.maxTexture2DLinear[2],
         prop.maxTexture2DGather[0], prop.maxTexture2DGather[1]);
  std::vector<cudaTextureObject_t> texs(NV);
  for (int v = 0; v < NV; ++v) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = src + v * n;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = (size_t)D1 * D2;
    rd.res.pitch2D.height = D0;
    rd.res.pitch2D.pitchInBytes = (size_t)D1 * D2 * 4;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    CK(cudaCreateTextureObject(&texs[v], &rd, &td, nullptr));
    CK(cudaEventRecord(b));
    (void)a; (void)b;
  }
  cudaTextureObject_t* dtex;
  CK(cudaMalloc(&dtex, NV * sizeof(cudaTextureObject_t)));
  CK(cudaMemcpy(dtex, texs.data(), NV * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));

  {  // component order
    float4* o;
    CK(cudaMalloc(&o, 16));
    const int i0 = 7, i1 = 3, i2 = 5;
    order_probe<<<1, 1>>>(texs[0], (float)(i1 * D2 + i2) + 1.0f, (float)i0 + 1.0f, o);
    CK(cudaDeviceSynchronize());
    float4 g;
    CK(cudaMemcpy(&g, o, 16, cudaMemcpyDeviceToHost));
    auto at = [&](int a0, int a2) { return h[((size_t)(i0 + a0) * D1 + i1) * D2 + i2 + a2]; };
    printf("gather order: x=%g y=%g z=%g w=%g | (y0,x0)=%g (y0,x1)=%g (y1,x0)=%g (y1,x1)=%g\n", g.x, g.y, g.z, g.w, at(0, 0), at(0, 1), at(1, 0), at(1, 1));
    const bool ok = g.w == at(0, 0) && g.z == at(0, 1) && g.x == at(1, 0) && g.y == at(1, 1);
    printf("gather on a pitch-2D texture: %s\n", ok ? "WORKS (order w,z,x,y = (y0x0),(y0x1),(y1x0),(y1x1))" : "MISMATCH");
  }

  // rotation about the centre (Rx Ry Rz), output -> source
  Map M;
  {
    const double cx = cos(ang[0]), sx = sin(ang[0]), cy = cos(ang[1]), sy = sin(ang[1]), cz = cos(ang[2]), sz = sin(ang[2]);
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    double T[9], R[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T[3 * i + j] = 0; for (int k = 0; k < 3; ++k) T[3 * i + j] += Rx[3 * i + k] * Ry[3 * k + j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { R[3 * i + j] = 0; for (int k = 0; k < 3; ++k) R[3 * i + j] += T[3 * i + k] * Rz[3 * k + j]; }
    const double c[3] = {(D0 - 1) / 2.0, (D1 - 1) / 2.0, (D2 - 1) / 2.0};
    for (int i = 0; i < 3; ++i) {
      double t = c[i];
      for (int j = 0; j < 3; ++j) { M.m[3 * i + j] = (float)R[3 * i + j]; t -= R[3 * i + j] * c[j]; }
      M.t[i] = (float)t;
    }
  }
  const dim3 grid(D1 / 8, D0 / 8, NV), block(32, 8);
  if (D2 != 32) { printf("the probe maps 32 lanes to axis 2: D2 must be 32\n"); return 1; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> out(n);
  const char* names[3] = {"2 x tex2Dgather", "8 x tex2D point", "8 x __ldg"};
  for (int mode = 0; mode < 3; ++mode) {
    CK(cudaMemset(dst, 0, h.size() * 4));
    for (int rep = 0; rep < 3; ++rep) {
      if (mode == 0) resample<0><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
      if (mode == 1) resample<1><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
      if (mode == 2) resample<2><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
    }
    CK(cudaDeviceSynchronize());
    const int reps = 10;
    CK(cudaEventRecord(e0));
    for (int rep = 0; rep < reps; ++rep) {
      if (mode == 0) resample<0><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
      if (mode == 1) resample<1><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
      if (mode == 2) resample<2><<<grid, block>>>(dtex, src, dst, D0, D1, D2, M);
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    CK(cudaMemcpy(out.data(), dst + (NV - 1) * n, n * 4, cudaMemcpyDeviceToHost));
    std::vector<float> hv(h.begin() + (NV - 1) * n, h.end());
    double maxerr = 0;
    size_t bad = 0, checked = 0;
    for (size_t i = 0; i < n; i += 97) {
      const int o2 = i % D2, o1 = (i / D2) % D1, o0 = i / ((size_t)D1 * D2);
      const float r = host_voxel(hv, D0, D1, D2, M, o0, o1, o2);
      const double e = fabs((double)r - out[i]);
      if (e > maxerr) maxerr = e;
      if (e > 1e-6) ++bad;
      ++checked;
    }
    const double vox = (double)n * NV;
    printf("%-18s %8.4f ms  %7.1f Gvox/s  %7.1f GB/s at 8 B per voxel (%.3f of 6541)   max |err| %.3g  (%zu of %zu sampled voxels off by > 1e-6)\n",
           names[mode], ms, vox / ms / 1e6, 8.0 * vox / ms / 1e6, 8.0 * vox / ms / 1e6 / 6541.1, maxerr, bad, checked);
  }
  return 0;
}
