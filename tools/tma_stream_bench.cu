// tma_stream_bench.cu -- development microbenchmark: how fast can one persistent CTA per SM stream a
// row-major fp32 [N][128] matrix through TMA into shared memory, as a function of box rows, pipeline
// depth and the number of TMA ops per stage?  (No compute: a consumer warp only releases the stages.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench tools/tma_stream_bench.cu
//   ./tma_stream_bench <box_rows> <stages> <split> [n_rows] [consumers]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "l"(hint) : "memory");
}

__global__ void __launch_bounds__(256, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm, int64_t n_rows, int box_rows, int stages, int split,
              int consumers, int hint_mode, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int tile_rows = box_rows * split;
  const uint32_t tile_bytes = (uint32_t)tile_rows * 512u;
  const uint32_t bar_full = sbase + stages * tile_bytes;
  const uint32_t bar_empty = bar_full + 8 * stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, consumers); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t total = (n_rows + tile_rows - 1) / tile_rows;
  const int64_t t0 = (int64_t)blockIdx.x * total / gridDim.x, t1 = (int64_t)(blockIdx.x + 1) * total / gridDim.x;
  const uint64_t hint = hint_mode == 0 ? 0x12F0000000000000ull : (hint_mode == 1 ? 0x1000000000000000ull : 0x14F0000000000000ull);
  float acc = 0.f;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int64_t t = t0; t < t1; ++t) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        mbar_expect_tx(bar_full + 8 * s, tile_bytes);
        for (int k = 0; k < split; ++k)
          tma_load_2d(sbase + s * tile_bytes + k * box_rows * 512, &tm, 0, (int)(t * tile_rows + k * box_rows), bar_full + 8 * s, hint);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp <= consumers) {
    int s = 0; uint32_t ph = 0;
    for (int64_t t = t0; t < t1; ++t) {
      mbar_wait(bar_full + 8 * s, ph);
      if (lane == 0) {
        float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + s * tile_bytes + warp * 4));
        acc += v;
        mbar_arrive(bar_empty + 8 * s);
      }
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 64;
  const int stages = argc > 2 ? atoi(argv[2]) : 4;
  const int split = argc > 3 ? atoi(argv[3]) : 1;
  const int64_t n = argc > 4 ? atoll(argv[4]) : 10000000;
  const int consumers = argc > 5 ? atoi(argv[5]) : 1;
  const int hint_mode = argc > 6 ? atoi(argv[6]) : 0;
  const int promo = argc > 7 ? atoi(argv[7]) : 2;
  float *X, *sink;
  CK(cudaMalloc(&X, (size_t)n * 512));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(X, 0, (size_t)n * 512));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap tm;
  cuuint64_t dims[2] = {128, (cuuint64_t)n}; cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, X, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const size_t smem = (size_t)stages * box_rows * split * 512 + 16 * stages + 1024;
  if (smem > 227 * 1024) { printf("box_rows=%d stages=%d split=%d: smem %zu too large\n", box_rows, stages, split, smem); return 0; }
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    CK(cudaEventRecord(e0));
    stream_kernel<<<sms, 256, smem>>>(tm, n, box_rows, stages, split, consumers, hint_mode, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (it > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  printf("box_rows=%3d stages=%2d split=%d consumers=%d hint=%d promo=%d smem=%3zuKB: %.3f ms  %.1f GB/s (%.3f of 6575)\n", box_rows, stages, split,
         consumers, hint_mode, promo, smem >> 10, best, n * 512.0 / best / 1e6, n * 512.0 / best / 1e6 / 6575.1);
  return 0;
}
