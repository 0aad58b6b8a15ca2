// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tma_probe tools/tma_probe.cu -lcuda
// Stand-alone probe: 3-D TMA tile load (fp32, no swizzle, OOB zero fill) with the descriptor
// passed (a) as a __grid_constant__ parameter and (b) through global memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool FROM_GLOBAL>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, float* out, int bx, int by, int bz,
                      int c0, int c1, int c2) {
  extern __shared__ __align__(128) float box[];
  __shared__ __align__(8) uint64_t mbar;
  const void* tm = FROM_GLOBAL ? (const void*)gmap : (const void*)&pmap;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (FROM_GLOBAL)
      asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"((uint64_t)tm) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bx * by * bz * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(box)), "l"((uint64_t)tm), "r"(smem_u32(&mbar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&mbar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bx * by * bz; i += blockDim.x) out[i] = box[i];
}

int main(int argc, char** argv) {
  const int v_global = argc > 1 ? atoi(argv[1]) : 0;
  const int a0 = argc > 2 ? atoi(argv[2]) : 0, a1 = argc > 3 ? atoi(argv[3]) : 0, a2 = argc > 4 ? atoi(argv[4]) : 0;
  const int abz = argc > 5 ? atoi(argv[5]) : 8;
  const int promo = argc > 6 ? atoi(argv[6]) : 0;
  const int X = 20, Y = 24, Z = 16;  // volume [X][Y][Z]
  const int bx = 6, by = 7, bz = abz;
  std::vector<float> h(X * Y * Z);
  for (int i = 0; i < X * Y * Z; ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, bx * by * bz * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (!fn) { printf("no driver entry point\n"); return 1; }
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  alignas(64) CUtensorMap tm;
  cuuint64_t gdim[3] = {Z, Y, X}, gstr[2] = {Z * 4, (cuuint64_t)Y * Z * 4};
  cuuint32_t bdim[3] = {bz, by, bx}, estr[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  CUtensorMap* gtm;
  cudaMalloc(&gtm, 512);
  cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  for (int variant = v_global; variant <= v_global; ++variant) {
    const int c0 = a0, c1 = a1, c2 = a2;  // z starts out of bounds, y runs out of bounds
    cudaMemset(o, 0xff, bx * by * bz * 4);
    if (variant == 0) probe<false><<<1, 128, bx * by * bz * 4>>>(tm, gtm, o, bx, by, bz, c0, c1, c2);
    else probe<true><<<1, 128, bx * by * bz * 4>>>(tm, gtm, o, bx, by, bz, c0, c1, c2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d (%s): %s\n", variant, variant ? "global" : "grid_constant", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> ho(bx * by * bz);
    cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int x = 0; x < bx; ++x) for (int y = 0; y < by; ++y) for (int z = 0; z < bz; ++z) {
      int gx = c2 + x, gy = c1 + y, gz = c0 + z;
      float want = (gx < 0 || gx >= X || gy < 0 || gy >= Y || gz < 0 || gz >= Z) ? 0.f : h[(gx * Y + gy) * Z + gz];
      if (ho[(x * by + y) * bz + z] != want) ++bad;
    }
    printf("  mismatches: %d\n", bad);
  }
  return 0;
}
