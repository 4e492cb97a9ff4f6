// Streaming-rate probe (diagnostics, not part of the library): how fast can 148 SMs pull a [rows, d] bf16 matrix
// through TMA into shared memory, as a function of the box shape and ring depth?  No MMA, no selection.
//   mode 0: 2-D tensor-map boxes {64 elements, R rows} with 128-byte swizzle (what the search kernels issue)
//   mode 1: 1-D bulk copies of `chunk` contiguous bytes (what a pre-tiled gallery would allow)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}

struct Args { const void* base; long long rows; int d; int box_rows; int stages; int mode; int chunk; long long rows_per_cta; };

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tmap, Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[32], empty[32];
  const int nkb = a.d / 64;
  const uint32_t stage_bytes = a.mode == 0 ? (uint32_t)a.box_rows * 128u : (uint32_t)a.chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * a.rows_per_cta;
  long long r1 = r0 + a.rows_per_cta; if (r1 > a.rows) r1 = a.rows;
  long long nslots;
  if (a.mode == 0) nslots = ((r1 - r0 + a.box_rows - 1) / a.box_rows) * nkb;
  else nslots = ((r1 - r0) * (long long)a.d * 2) / a.chunk;
  if (threadIdx.x == 0) {  // producer
    int stage = 0; uint32_t phase = 0;
    for (long long i = 0; i < nslots; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_expect(&full[stage], stage_bytes);
      uint8_t* dst = smem + (size_t)stage * stage_bytes;
      if (a.mode == 0) {
        const int kb = (int)(i % nkb); const long long t = i / nkb;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(smem_u32(dst)), "l"((uint64_t)&tmap), "r"(smem_u32(&full[stage])), "r"(kb * 64), "r"((int)(r0 + t * a.box_rows)) : "memory");
      } else {
        const uint8_t* src = (const uint8_t*)a.base + r0 * (long long)a.d * 2 + i * (long long)a.chunk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(dst)), "l"(src), "r"(stage_bytes), "r"(smem_u32(&full[stage])) : "memory");
      }
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {  // consumer
    int stage = 0; uint32_t phase = 0;
    for (long long i = 0; i < nslots; ++i) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == a.stages) { stage = 0; phase ^= 1; }
    }
  }
}

int main(int argc, char** argv) {
  const long long rows = argc > 1 ? atoll(argv[1]) : 10000000;
  const int d = argc > 2 ? atoi(argv[2]) : 768;
  void* g; CK(cudaMalloc(&g, (size_t)rows * d * 2)); CK(cudaMemset(g, 1, (size_t)rows * d * 2));
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn; cudaDriverEntryPointQueryResult q; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  Enc enc = (Enc)fn;
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int mode, box_rows, stages, chunk, waves, promo; };
  Cfg cfgs[] = {
    {0, 256, 4, 0, 2, 2}, {0, 256, 6, 0, 2, 2}, {0, 256, 6, 0, 1, 2}, {0, 256, 6, 0, 2, 1}, {0, 256, 6, 0, 2, 0},
    {0, 128, 8, 0, 2, 2}, {0, 128, 13, 0, 2, 2}, {0, 64, 16, 0, 2, 2}, {0, 64, 26, 0, 2, 2},
    {1, 0, 6, 32768, 2, 0}, {1, 0, 13, 16384, 2, 0}, {1, 0, 26, 8192, 2, 0}, {1, 0, 3, 65536, 2, 0}, {1, 0, 6, 32768, 1, 0}, {1, 0, 6, 32768, 4, 0},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm; memset(&tm, 0, sizeof(tm));
    if (c.mode == 0) {
      cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows}; cuuint64_t gstr[1] = {(cuuint64_t)d * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows}; cuuint32_t es[2] = {1, 1};
      CUtensorMapL2promotion pr = c.promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : c.promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    }
    const int grid = sms * c.waves;
    Args a; a.base = g; a.rows = rows; a.d = d; a.box_rows = c.box_rows; a.stages = c.stages; a.mode = c.mode; a.chunk = c.chunk;
    long long rpc = (rows + grid - 1) / grid; rpc = (rpc + 255) / 256 * 256; a.rows_per_cta = rpc;
    const size_t smem = (size_t)c.stages * (c.mode == 0 ? c.box_rows * 128 : c.chunk);
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) {
      cudaEventRecord(e0);
      probe<<<grid, 64, smem>>>(tm, a);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("mode=%d box_rows=%3d chunk=%6d stages=%2d waves=%d promo=%d smem=%3zu KB : %.3f ms  %.0f GB/s\n", c.mode, c.box_rows, c.chunk, c.stages, c.waves, c.promo, smem / 1024, best, (double)rows * d * 2 / best / 1e6);
  }
  return 0;
}
