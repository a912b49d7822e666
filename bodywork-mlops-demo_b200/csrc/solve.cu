// solve.cu -- single-SM fp64 solve of the centred (ridge) normal equations from S.
//
// Replaces scipy.linalg.lstsq + _set_intercept inside LinearRegression.fit
// (stage_1_train_model.py:105-106 -> sklearn/linear_model/_base.py: `linalg.lstsq(Xc, yc, cond=tol)`
// then `intercept_ = y_offset - X_offset @ coef_`).  Ridge term as sklearn/linear_model/_ridge.py
// (`(Xc^T Xc + alpha I) w = Xc^T yc`).
//
//   solve_cholesky_kernel : A = Xc^T Xc + alpha I = L L^T in shared memory, two triangular solves.
//   solve_spectral_kernel : one-sided Jacobi on A (A V = W, columns of W orthogonal => w_k = lambda_k v_k):
//                           singular_ = sqrt(lambda) (descending), rank_ = #{sqrt(lambda) > cond * max},
//                           coef = minimum-norm solution = what gelsd returns for rank-deficient X.
//
// One CTA: the matrices are <= 128 x 128 fp64 (132 KB with padding) -- latency bound, not a
// throughput problem (D^3/3 = 0.7 MFLOP).
#include "b2_internal.cuh"

namespace b2 {
namespace {

constexpr int kSolveThreads = 256;
constexpr int kOutIntercept = kMaxD;      // solve_out layout: [0,d) coef | intercept | info | rank | singular[d]
constexpr int kOutInfo = kMaxD + 1;
constexpr int kOutRank = kMaxD + 2;
constexpr int kOutSingular = kMaxD + 3;

// Builds A (pitch d+1) and r in shared memory from the raw statistic; returns means.
__device__ void build_normal_equations(const double* __restrict__ S, int d, double alpha, int fit_intercept,
                                       double* A, double* r, double* mean, double* ybar_out) {
  const int dp = d + 2, pitch = d + 1;
  const double n = S[d * dp + d];
  const double inv_n = n > 0.0 ? 1.0 / n : 0.0;
  for (int j = threadIdx.x; j < d; j += blockDim.x) mean[j] = fit_intercept ? S[j * dp + d] * inv_n : 0.0;
  __syncthreads();
  const double ybar = fit_intercept ? S[d * dp + d + 1] * inv_n : 0.0;
  for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
    const int i = idx / d, j = idx - i * d;
    // symmetrise explicitly; centre with n * mean_i * mean_j
    double v = 0.5 * (S[i * dp + j] + S[j * dp + i]) - n * mean[i] * mean[j];
    if (i == j) v += alpha;
    A[i * pitch + j] = v;
  }
  for (int i = threadIdx.x; i < d; i += blockDim.x) r[i] = S[i * dp + d + 1] - n * mean[i] * ybar;
  if (threadIdx.x == 0) *ybar_out = ybar;
  __syncthreads();
}

__global__ void __launch_bounds__(kSolveThreads, 1)
solve_cholesky_kernel(const double* __restrict__ S, int d, double alpha, int fit_intercept,
                      double* __restrict__ out) {
  extern __shared__ double sm[];
  const int pitch = d + 1;
  double* A = sm;                  // d x (d+1)
  double* r = A + d * pitch;       // d
  double* mean = r + d;            // d
  double* misc = mean + d;         // [0] ybar, [1] max diag, [2] info
  build_normal_equations(S, d, alpha, fit_intercept, A, r, mean, &misc[0]);
  if (threadIdx.x == 0) {
    double mx = 0.0;
    for (int i = 0; i < d; ++i) mx = fmax(mx, A[i * pitch + i]);
    misc[1] = mx;
    misc[2] = 0.0;
  }
  __syncthreads();
  const double tiny = misc[1] * 1e-12;

  // right-looking Cholesky, lower triangle in place
  for (int k = 0; k < d; ++k) {
    if (threadIdx.x == 0) {
      const double piv = A[k * pitch + k];
      if (!(piv > tiny)) { misc[2] = (double)(k + 1); A[k * pitch + k] = 1.0; }
      else A[k * pitch + k] = sqrt(piv);
    }
    __syncthreads();
    if (misc[2] != 0.0) break;
    const double inv = 1.0 / A[k * pitch + k];
    for (int i = k + 1 + threadIdx.x; i < d; i += blockDim.x) A[i * pitch + k] *= inv;
    __syncthreads();
    // trailing update: element (i, j), k < j <= i ; 2 threads per row
    const int rows = d - k - 1;
    for (int t = threadIdx.x; t < rows * 2; t += blockDim.x) {
      const int i = k + 1 + (t >> 1);
      const double lik = A[i * pitch + k];
      for (int j = k + 1 + (t & 1); j <= i; j += 2) A[i * pitch + j] -= lik * A[j * pitch + k];
    }
    __syncthreads();
  }
  const bool singular = misc[2] != 0.0;
  if (!singular) {
    // forward: L z = r (column oriented)
    for (int k = 0; k < d; ++k) {
      if (threadIdx.x == 0) r[k] /= A[k * pitch + k];
      __syncthreads();
      const double zk = r[k];
      for (int i = k + 1 + threadIdx.x; i < d; i += blockDim.x) r[i] -= A[i * pitch + k] * zk;
      __syncthreads();
    }
    // backward: L^T b = z
    for (int k = d - 1; k >= 0; --k) {
      if (threadIdx.x == 0) r[k] /= A[k * pitch + k];
      __syncthreads();
      const double bk = r[k];
      for (int i = threadIdx.x; i < k; i += blockDim.x) r[i] -= A[k * pitch + i] * bk;
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[i] = singular ? 0.0 : r[i];
  if (threadIdx.x == 0) {
    double b0 = misc[0];
    if (!singular) for (int i = 0; i < d; ++i) b0 -= mean[i] * r[i];
    out[kOutIntercept] = singular ? 0.0 : b0;
    out[kOutInfo] = misc[2];
  }
}

// pair (p, q) of slot k in round `round` of the round-robin tournament over m (even) players
__device__ __forceinline__ void rr_pair(int m, int round, int k, int* p, int* q) {
  const int mm = m - 1;
  int a, b;
  if (k == 0) { a = mm; b = round % mm; }
  else { a = (round + k) % mm; b = (round - k + mm) % mm; }
  *p = a < b ? a : b;
  *q = a < b ? b : a;
}

__global__ void __launch_bounds__(512, 1)
solve_spectral_kernel(const double* __restrict__ S, int d, double cond, int fit_intercept,
                      double* __restrict__ out) {
  extern __shared__ double sm[];
  const int pitch = d + 1;
  double* W = sm;                  // W^T: W[col * pitch + row]   (A is symmetric, so W0 = A either way)
  double* r = W + d * pitch;
  double* mean = r + d;
  double* lam = mean + d;          // column norms
  double* misc = lam + d;          // [0] ybar
  __shared__ int rotated;
  build_normal_equations(S, d, 0.0, fit_intercept, W, r, mean, &misc[0]);

  const int m = d + (d & 1);       // even player count; player d (if any) is a phantom
  const int pairs = m / 2;
  const int sub = threadIdx.x & 7; // 8 threads cooperate on one pair
  const int slot0 = threadIdx.x >> 3;
  for (int sweep = 0; sweep < 24 && d > 1; ++sweep) {
    if (threadIdx.x == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < m - 1; ++round) {
      for (int slot = slot0; slot < ((pairs + 63) / 64) * 64; slot += 64) {
        int p = 0, q = 0;
        const bool live_slot = slot < pairs;
        if (live_slot) rr_pair(m, round, slot, &p, &q);
        const bool live = live_slot && q < d;
        double a = 0.0, b = 0.0, g = 0.0;
        if (live) {
          for (int row = sub; row < d; row += 8) {
            const double wp = W[p * pitch + row], wq = W[q * pitch + row];
            a += wp * wp; b += wq * wq; g += wp * wq;
          }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
          g += __shfl_xor_sync(0xffffffffu, g, o);
        }
        if (live && fabs(g) > 1e-15 * sqrt(a * b) && a * b > 0.0) {
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
          for (int row = sub; row < d; row += 8) {
            const double wp = W[p * pitch + row], wq = W[q * pitch + row];
            W[p * pitch + row] = c * wp - s * wq;
            W[q * pitch + row] = s * wp + c * wq;
          }
          if (sub == 0 && fabs(g) > 1e-13 * sqrt(a * b)) rotated = 1;
        }
      }
      __syncthreads();
    }
    if (!rotated) break;
    __syncthreads();
  }
  // eigenvalues = column norms
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    double s2 = 0.0;
    for (int row = 0; row < d; ++row) s2 += W[k * pitch + row] * W[k * pitch + row];
    lam[k] = sqrt(s2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // descending singular values (selection sort on a copy in `out`), rank, min-norm coefficients
    double mx = 0.0;
    for (int k = 0; k < d; ++k) mx = fmax(mx, lam[k]);
    const double smax = sqrt(mx);
    int rank = 0;
    for (int k = 0; k < d; ++k) {
      out[kOutSingular + k] = sqrt(lam[k]);
      if (sqrt(lam[k]) > cond * smax) ++rank;
    }
    for (int i = 0; i < d; ++i) {
      int best = i;
      for (int j = i + 1; j < d; ++j) if (out[kOutSingular + j] > out[kOutSingular + best]) best = j;
      const double tmp = out[kOutSingular + i];
      out[kOutSingular + i] = out[kOutSingular + best];
      out[kOutSingular + best] = tmp;
    }
    out[kOutRank] = (double)rank;
    misc[1] = smax;
  }
  __syncthreads();
  const double smax = misc[1];
  // coef = sum_k w_k (w_k . r) / lambda_k^3 over kept k
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    double dot = 0.0;
    for (int row = 0; row < d; ++row) dot += W[k * pitch + row] * r[row];
    const bool keep = sqrt(lam[k]) > cond * smax && lam[k] > 0.0;
    lam[k] = keep ? dot / (lam[k] * lam[k] * lam[k]) : 0.0;  // reuse lam as the per-column weight
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    double b = 0.0;
    for (int k = 0; k < d; ++k) b += W[k * pitch + i] * lam[k];
    r[i] = b;
    out[i] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double b0 = misc[0];
    for (int i = 0; i < d; ++i) b0 -= mean[i] * r[i];
    out[kOutIntercept] = b0;
    out[kOutInfo] = 0.0;
  }
}

size_t solve_smem_bytes(int d) { return sizeof(double) * ((size_t)d * (d + 1) + 3 * d + 8); }

}  // namespace

int launch_solve_cholesky(b2_ctx* ctx, double alpha, int fit_intercept) {
  const size_t smem = solve_smem_bytes(ctx->d);
  B2_CUDA(cudaFuncSetAttribute(solve_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  solve_cholesky_kernel<<<1, kSolveThreads, smem, ctx->stream>>>(ctx->S, ctx->d, alpha, fit_intercept,
                                                                 ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

int launch_solve_spectral(b2_ctx* ctx, double cond, int fit_intercept) {
  const size_t smem = solve_smem_bytes(ctx->d);
  B2_CUDA(cudaFuncSetAttribute(solve_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  solve_spectral_kernel<<<1, 512, smem, ctx->stream>>>(ctx->S, ctx->d, cond, fit_intercept, ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

}  // namespace b2
