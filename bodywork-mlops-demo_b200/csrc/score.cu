// score.cu -- batch scoring yhat = X.coef + intercept fused with the hold-out metric reductions.
//
// Replaces  ols_regressor.predict(X_test)        stage_1_train_model.py:107
//           model.predict(X)                     stage_2_serve_model.py:78
//           model_metrics(y_actual, y_predicted) stage_1_train_model.py:79-90
//             MAPE = mean(|yhat-y| / max(|y|, eps_f64)), R^2 = 1 - SSres/SStot, max_error = max|y-yhat|
//
// HBM-bound: one pass over X (D*sizeof(x) bytes per row) + 4 B (y) + 4 B (yhat).  One warp per
// row, 128-bit loads, the dot product and all reductions in fp64 (products of an fp32 value with
// an fp64 coefficient; this reproduces the float64 predict of the oracle to ~1e-13).
#include <cuda_bf16.h>

#include "b2_internal.cuh"

namespace b2 {
namespace {

constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr double kEpsF64 = 2.220446049250313e-16;

constexpr int kNStats = 10;   // include/b2gram.h: b2_score stats_out layout
struct RowStats {
  double ape = 0.0, sse = 0.0, sy = 0.0, syy = 0.0, mx = 0.0, cnt = 0.0, sp = 0.0, spp = 0.0, syp = 0.0, mxape = 0.0;
  __device__ void add(double y, double p) {
    const double e = fabs(p - y);
    ape += e / fmax(fabs(y), kEpsF64);            // sklearn MAPE term (stage_1_train_model.py:81)
    sse += (y - p) * (y - p);
    sy += y;
    syy += y * y;
    mx = fmax(mx, e);
    cnt += 1.0;
    sp += p;                                      // Pearson correlation terms (stage_4...:103 "r_squared")
    spp += p * p;
    syp += y * p;
    mxape = fmax(mxape, e / fabs(y));             // |score/label - 1| (stage_4...:89,104); inf when label == 0
  }
};
__device__ __forceinline__ bool stat_is_max(int k) { return k == 4 || k == 9; }

template <typename T>
__device__ __forceinline__ double lane_dot(const T* __restrict__ row, int d, int lane, const double* cf, bool vec);

template <>
__device__ __forceinline__ double lane_dot<float>(const float* __restrict__ row, int d, int lane, const double* cf,
                                                  bool vec) {
  double acc = 0.0;
  if (vec) {
    if (lane * 4 < d) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(row) + lane);
      acc = fma((double)x.x, cf[0], acc);
      acc = fma((double)x.y, cf[1], acc);
      acc = fma((double)x.z, cf[2], acc);
      acc = fma((double)x.w, cf[3], acc);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lane + 32 * k;
      if (j < d) acc = fma((double)__ldg(row + j), cf[k], acc);
    }
  }
  return acc;
}

template <>
__device__ __forceinline__ double lane_dot<__nv_bfloat16>(const __nv_bfloat16* __restrict__ row, int d, int lane,
                                                          const double* cf, bool vec) {
  double acc = 0.0;
  if (vec) {
    if (lane * 4 < d) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(row) + lane);
      acc = fma((double)__uint_as_float(u.x << 16), cf[0], acc);
      acc = fma((double)__uint_as_float(u.x & 0xffff0000u), cf[1], acc);
      acc = fma((double)__uint_as_float(u.y << 16), cf[2], acc);
      acc = fma((double)__uint_as_float(u.y & 0xffff0000u), cf[3], acc);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = lane + 32 * k;
      if (j < d) acc = fma((double)__bfloat162float(row[j]), cf[k], acc);
    }
  }
  return acc;
}

template <typename T>
__global__ void __launch_bounds__(kScoreThreads)
score_kernel(const T* __restrict__ X, int64_t n, int d, int64_t ldx, const double* __restrict__ coef,
             const float* __restrict__ y, const uint8_t* __restrict__ mask, int keep, float* __restrict__ yhat,
             int vec, double* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double cf[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = vec ? lane * 4 + k : lane + 32 * k;
    cf[k] = j < d ? coef[j] : 0.0;
  }
  const double b0 = coef[kMaxD];
  RowStats st;
  const int64_t warps_total = (int64_t)gridDim.x * kScoreWarps;
  const int64_t gw = (int64_t)blockIdx.x * kScoreWarps + warp;
  constexpr int kU = 4;  // rows in flight per warp
  for (int64_t base = gw * kU; base < n; base += warps_total * kU) {
    double acc[kU];
    bool use[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t row = base + u;
      use[u] = row < n;
      if (use[u] && mask != nullptr) use[u] = (__ldg(mask + row) == (uint8_t)keep);
      acc[u] = use[u] ? lane_dot<T>(X + row * ldx, d, lane, cf, vec != 0) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t row = base + u;
        if (row < n) {
          const double p = acc[u] + b0;
          if (yhat != nullptr) yhat[row] = use[u] ? (float)p : 0.f;
          if (use[u] && y != nullptr) st.add((double)__ldg(y + row), p);
        }
      }
    }
  }
  __shared__ double red[kScoreWarps][kNStats];
  if (lane == 0) {
    red[warp][0] = st.ape; red[warp][1] = st.sse; red[warp][2] = st.sy;  red[warp][3] = st.syy; red[warp][4] = st.mx;
    red[warp][5] = st.cnt; red[warp][6] = st.sp;  red[warp][7] = st.spp; red[warp][8] = st.syp; red[warp][9] = st.mxape;
  }
  __syncthreads();
  if (threadIdx.x < kNStats) {
    const int k = threadIdx.x;
    double v = 0.0;
    for (int w = 0; w < kScoreWarps; ++w) v = stat_is_max(k) ? fmax(v, red[w][k]) : v + red[w][k];
    part[(size_t)blockIdx.x * kNStats + k] = v;
  }
}

// acc[0..5] (at part + n_ctas*6 ... see launch) = combine over CTAs in order; `first` overwrites.
__global__ void score_reduce_kernel(const double* __restrict__ part, int n_ctas, int first, double* __restrict__ acc) {
  const int k = threadIdx.x;
  if (k >= kNStats) return;
  double v = first ? 0.0 : acc[k];
  for (int c = 0; c < n_ctas; ++c)
    v = stat_is_max(k) ? fmax(v, part[(size_t)c * kNStats + k]) : v + part[(size_t)c * kNStats + k];
  acc[k] = v;
}

}  // namespace

// ctx->score_part layout: [score_ctas][10] partials, then 10 doubles of running totals.
int launch_score(b2_ctx* ctx, const void* X, int x_dtype, int64_t n, int d, int64_t ldx, const float* y,
                 const uint8_t* mask, int keep, float* yhat, bool first_block) {
  const int es = x_dtype == B2_F32 ? 4 : 2;
  const int vec = (d % 4 == 0) && ((ldx * es) % (4 * es) == 0) && ((reinterpret_cast<uintptr_t>(X) % (4 * es)) == 0);
  int64_t want = (n + kScoreWarps * 4 - 1) / (kScoreWarps * 4);
  if (want < 1) want = 1;
  const int grid = (int)(want < ctx->score_ctas ? want : ctx->score_ctas);
  double* acc = ctx->score_part + (size_t)ctx->score_ctas * kNStats;
  if (x_dtype == B2_F32)
    score_kernel<float><<<grid, kScoreThreads, 0, ctx->stream>>>(static_cast<const float*>(X), n, d, ldx,
                                                                 ctx->coef_dev, y, mask, keep, yhat, vec,
                                                                 ctx->score_part);
  else
    score_kernel<__nv_bfloat16><<<grid, kScoreThreads, 0, ctx->stream>>>(static_cast<const __nv_bfloat16*>(X), n, d,
                                                                         ldx, ctx->coef_dev, y, mask, keep, yhat,
                                                                         vec, ctx->score_part);
  B2_CUDA(cudaGetLastError());
  score_reduce_kernel<<<1, 32, 0, ctx->stream>>>(ctx->score_part, grid, first_block ? 1 : 0, acc);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return B2_OK;
}

}  // namespace b2
