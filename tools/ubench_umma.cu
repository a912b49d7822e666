// ubench_umma.cu -- cost of one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as the Gram kernels issue it:
// A from tensor memory or from shared memory, N = 16 .. 256, one or two accumulators, K-major no-swizzle operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_umma tools/ubench_umma.cu && ./ubench_umma
// Prints cycles per MMA (clock64 around `reps` back-to-back issues + the commit's mbarrier wait), 1 CTA and 148 CTAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// mode bit 0: A from smem (else TMEM); bit 1: alternate two accumulators; bit 2: rotate the B stage (3 stages)
__global__ void __launch_bounds__(128, 1) k(int n_cols, int mode, int reps, long long* out, int commit_every) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_ptr;
  const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
  for (uint32_t o = threadIdx.x * 16; o < 3 * 40960; o += blockDim.x * 16)
    *reinterpret_cast<uint4*>(raw + (sbase - smem_u32(raw)) + o) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_ptr;
  if (threadIdx.x < 32) {
    const uint32_t lbo = 2560;                       // K-group stride of the operand stage ([E | hi | E] groups)
    const uint32_t id = idesc(n_cols);
    long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {            // pass 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
          const uint32_t st = (mode & 4) ? (uint32_t)(r % 3) * 40960u : 0u;
#pragma unroll
          for (int k2 = 0; k2 < 4; ++k2) {
            const uint32_t d = tb + (((mode & 2) && (k2 & 1)) ? 256u : 0u);
            const uint64_t b = make_desc(sbase + st + k2 * 2 * lbo, lbo);
            if (mode & 1) mma_ss(d, make_desc(sbase + st + k2 * 2 * lbo + 20480u, lbo), b, id, 1u);
            else mma_ts(d, tb + 480 + k2 * 8, b, id, 1u);
          }
          if (commit_every > 0 && (r & 1)) {          // every 8 MMAs: a stage hand-off (nobody waits on it) + `commit_every - 1`
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
            const long long t_end = clock64() + (commit_every - 1);      // cycles of issue-thread work before the next MMA
            while (clock64() < t_end) {}
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)pass) : "memory");
      t1 = clock64();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  const int smem = 3 * 40960 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 256;
  const int ns[] = {16, 64, 128, 144, 160, 192, 256};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 8; ++mode) {
      printf("grid %3d  A %-4s  acc %s  B %-6s :", grid, (mode & 1) ? "smem" : "tmem", (mode & 2) ? "2" : "1", (mode & 4) ? "rotate" : "same");
      for (int n : ns) {
        if ((mode & 2) && n > 240) { printf("  N%-3d    -", n); continue; }
        k<<<grid, 128, smem>>>(n, mode, reps, out, 0);
        long long h[148];
        cudaError_t e = cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf(" CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("  N%-3d %5.1f", n, (double)mx / (reps * 4));
      }
      printf("\n");
    }
  }
  // tcgen05.commit between groups of MMAs (as a pipelined kernel issues them: one or two per tile of 4 or 8 MMAs)
  for (int ce : {0, 1, 51, 101, 151, 201, 301, 401, 601}) {
    printf("grid 148  N144  bursts of 8 MMAs + commit + %3d cycles of issue-thread delay :", ce > 0 ? ce - 1 : -1);
    for (int mode : {0, 1}) {
      k<<<148, 128, smem>>>(144, mode, reps, out, ce);
      long long h[148];
      cudaMemcpy(h, out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("  %s %6.1f cycles / burst", mode ? "A smem" : "A tmem", (double)mx / (reps / 2));
    }
    printf("\n");
  }
  return 0;
}
