// synth.cu -- synthetic rows on the device, following the reference's data-generating process
// (stage_3_synthetic_data_generation.py:36-41) generalised to D feature columns:
//     X_ij ~ U(0, 100),  eps_i ~ N(0, 1),  y_i = alpha + beta * sum_j X_ij + sigma * eps_i
// Counter-based Philox4x32-10 keyed by the seed; the counter is (global row, column block), so any
// shard of the dataset can be generated independently (multi-GPU row sharding draws the same rows
// as a single GPU would).  Benchmark input only -- 51 GB never has to cross PCIe.
#include <cuda_bf16.h>

#include "b2_internal.cuh"

namespace b2 {
namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

template <typename T>
__global__ void synth_kernel(uint64_t seed, int64_t row_offset, int64_t n, int d, int64_t ldx, float alpha,
                             float beta, float sigma, T* __restrict__ X, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (int64_t row = warp; row < n; row += warps) {
    const uint64_t g = (uint64_t)(row + row_offset);
    float sum = 0.f;
    const int c0 = lane * 4;
    if (c0 < d) {
      const uint4 rnd = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)lane, 0u), key);
      const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c0 + k < d) {
          float x = 100.0f * u01(w[k]);
          if (sizeof(T) == 2) {
            const __nv_bfloat16 xb = __float2bfloat16_rn(x);
            x = __bfloat162float(xb);
            reinterpret_cast<__nv_bfloat16*>(X)[row * ldx + c0 + k] = xb;
          } else {
            reinterpret_cast<float*>(X)[row * ldx + c0 + k] = x;
          }
          sum += x;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) {
      const uint4 rnd = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0xFFFFFFFFu, 1u), key);
      const float u1 = fmaxf(u01(rnd.x), 5.9604645e-8f), u2 = u01(rnd.y);
      const float eps = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);  // Box-Muller
      y[row] = alpha + beta * sum + sigma * eps;
    }
  }
}

// ---- one reference tranche on the device: stage_3_synthetic_data_generation.py:28-43 -------------------------------
//   X ~ U(0, 100), eps ~ N(0, 1), y = alpha(day) + beta X + sigma eps, and rows with y < 0 are DROPPED (:43), the rest
//   keep their order.  Stable compaction in three small launches (flag + count per block, scan of the block counts,
//   regenerate + write at the scanned offset): the same seed gives the same tranche whatever the launch geometry.
constexpr int kTrThreads = 256;

__device__ __forceinline__ void tranche_row(uint2 key, uint64_t row, float alpha, float beta, float sigma, float* x, float* y) {
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)(row >> 32), 0x7A3Du, 2u), key);
  const float xv = 100.0f * u01(rnd.x);
  const float u1 = fmaxf(u01(rnd.y), 5.9604645e-8f), u2 = u01(rnd.z);
  const float eps = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  *x = xv;
  *y = alpha + beta * xv + sigma * eps;
}

__global__ void __launch_bounds__(kTrThreads)
tranche_count_kernel(uint64_t seed, int64_t n, float alpha, float beta, float sigma, long long* __restrict__ counts) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const int64_t row = (int64_t)blockIdx.x * kTrThreads + threadIdx.x;
  float x = 0.f, y = -1.f;
  if (row < n) tranche_row(key, (uint64_t)row, alpha, beta, sigma, &x, &y);
  const int c = __syncthreads_count(row < n && y >= 0.f);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
tranche_scan_kernel(long long* __restrict__ counts, int64_t n_blocks, long long* __restrict__ total) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n_blocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const long long v = i < n_blocks ? counts[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_sum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    const long long before = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + inc - v;
    if (i < n_blocks) counts[i] = before;                                 // exclusive prefix
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kTrThreads)
tranche_write_kernel(uint64_t seed, int64_t n, float alpha, float beta, float sigma, const long long* __restrict__ offsets,
                     float* __restrict__ X, float* __restrict__ yout) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const int64_t row = (int64_t)blockIdx.x * kTrThreads + threadIdx.x;
  float x = 0.f, y = -1.f;
  if (row < n) tranche_row(key, (uint64_t)row, alpha, beta, sigma, &x, &y);
  const bool keep = row < n && y >= 0.f;
  __shared__ int warp_cnt[kTrThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  __syncthreads();
  int before = __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) before += warp_cnt[w];
  if (keep) {
    const long long dst = offsets[blockIdx.x] + before;
    X[dst] = x;
    yout[dst] = y;
  }
}

}  // namespace

int launch_synth_tranche(b2_ctx* ctx, uint64_t seed, int64_t n, double alpha, double beta, double sigma, float* X, float* y,
                         int64_t* n_kept_dev) {
  long long* total = reinterpret_cast<long long*>(n_kept_dev);
  if (n <= 0) {
    B2_CUDA(cudaMemsetAsync(total, 0, sizeof(long long), ctx->stream));
    return B2_OK;
  }
  const int64_t n_blocks = (n + kTrThreads - 1) / kTrThreads;
  const int64_t cap = (int64_t)ctx->simt_ctas * kMaxS * kMaxS;      // the fp64 SIMT partial block doubles as scratch
  if (n_blocks > cap || n_blocks > 0x7fffffff) { set_error("b2_synth_tranche: at most %lld rows per call", (long long)cap * kTrThreads); return B2_E_ARG; }
  long long* counts = reinterpret_cast<long long*>(ctx->simt_part);
  tranche_count_kernel<<<(int)n_blocks, kTrThreads, 0, ctx->stream>>>(seed, n, (float)alpha, (float)beta, (float)sigma, counts);
  B2_CUDA(cudaGetLastError());
  tranche_scan_kernel<<<1, 1024, 0, ctx->stream>>>(counts, n_blocks, total);
  B2_CUDA(cudaGetLastError());
  tranche_write_kernel<<<(int)n_blocks, kTrThreads, 0, ctx->stream>>>(seed, n, (float)alpha, (float)beta, (float)sigma, counts, X, y);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 3;
  return B2_OK;
}

int launch_synth(b2_ctx* ctx, uint64_t seed, int64_t row_offset, int64_t n, int d, int64_t ldx, int x_dtype,
                 double alpha, double beta, double sigma, void* X, float* y) {
  if (n <= 0) return B2_OK;
  const int threads = 256;
  int64_t blocks = (n * 32 + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (x_dtype == B2_F32)
    synth_kernel<float><<<(int)blocks, threads, 0, ctx->stream>>>(seed, row_offset, n, d, ldx, (float)alpha,
                                                                   (float)beta, (float)sigma, static_cast<float*>(X), y);
  else
    synth_kernel<__nv_bfloat16><<<(int)blocks, threads, 0, ctx->stream>>>(
        seed, row_offset, n, d, ldx, (float)alpha, (float)beta, (float)sigma, static_cast<__nv_bfloat16*>(X), y);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

}  // namespace b2
