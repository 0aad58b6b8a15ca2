// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tma_timing tools/tma_timing.cu -lcuda
// Micro-timing on one SM: cost of the tensormap proxy fence (sys / gpu scope) and of one 3-D TMA
// box load with short inner rows, in SM cycles.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void timing(const CUtensorMap* gmap, long long* out, int bytes, int reps) {
  extern __shared__ __align__(128) float box[];
  __shared__ __align__(8) uint64_t mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"((uint64_t)gmap) : "memory");
    long long t1 = clock64();
    for (int i = 0; i < reps; ++i) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"((uint64_t)gmap) : "memory");
    long long t2 = clock64();
    out[0] = (t1 - t0) / reps; out[1] = (t2 - t1) / reps;
    long long tt = 0;
    for (int i = 0; i < reps; ++i) {
      long long a = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(box)), "l"((uint64_t)gmap), "r"(smem_u32(&mbar)), "r"(4 * (i % 5)), "r"(3 + i % 7), "r"(5 + i % 3) : "memory");
      asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&mbar)), "r"(i & 1) : "memory");
      tt += clock64() - a;
    }
    out[2] = tt / reps;
  }
}
int main(int argc, char** argv) {
  const int X = 256, Y = 256, Z = argc > 4 ? atoi(argv[4]) : 32;
  const int bx = argc > 1 ? atoi(argv[1]) : 25, by = argc > 2 ? atoi(argv[2]) : 25, bz = argc > 3 ? atoi(argv[3]) : 36;
  float* d; cudaMalloc(&d, (size_t)X * Y * Z * 4); cudaMemset(d, 0, (size_t)X * Y * Z * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  alignas(64) CUtensorMap tm;
  cuuint64_t gdim[3] = {(cuuint64_t)Z, Y, X}, gstr[2] = {(cuuint64_t)Z * 4, (cuuint64_t)Y * Z * 4};
  cuuint32_t bdim[3] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx}, estr[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUtensorMap* gtm; cudaMalloc(&gtm, 512); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  long long* o; cudaMalloc(&o, 64);
  const int bytes = bx * by * bz * 4;
  cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  timing<<<1, 32, bytes>>>(gtm, o, bytes, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[3]; cudaMemcpy(h, o, 24, cudaMemcpyDeviceToHost);
  printf("encode rc=%d err=%s box %dx%dx%d (%d B): fence.sys %lld cyc, fence.gpu %lld cyc, TMA box load %lld cyc (%.1f B/cyc)\n", (int)r,
         cudaGetErrorString(e), bx, by, bz, bytes, h[0], h[1], h[2], (double)bytes / h[2]);
  return 0;
}
