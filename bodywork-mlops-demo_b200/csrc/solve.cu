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
  // one warp per row, coalesced; S is symmetric by construction (tc_fold / the SIMT reduce write both halves)
  for (int i = threadIdx.x >> 5; i < d; i += blockDim.x >> 5) {
    const double mi = mean[i];
    for (int j = threadIdx.x & 31; j < d; j += 32) {
      double v = S[i * dp + j] - n * mi * mean[j];
      if (i == j) v += alpha;
      A[i * pitch + j] = v;
    }
  }
  for (int i = threadIdx.x; i < d; i += blockDim.x) r[i] = S[i * dp + d + 1] - n * mean[i] * ybar;
  if (threadIdx.x == 0) *ybar_out = ybar;
  __syncthreads();
}

// 1/sqrt(x) for normal positive x without the library's slow-path call (a call inside the unrolled pivot
// loop forces the register-resident block onto the stack): scale x by an even power of two into [1, 4),
// fp32 MUFU seed, two Newton steps in fp64 (2^-23 -> 2^-45 -> below 2^-53), undo the scaling.
__device__ __forceinline__ double rsqrt_pos(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int e2 = ((((hi >> 20) & 0x7ff) - 1023)) & ~1;                 // even exponent
  const double xs = __hiloint2double(hi - (e2 << 20), lo);              // x * 2^-e2 in [1, 4)
  double y = (double)rsqrtf((float)xs);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double t = xs * y, h = 0.5 * y;
    const double e = fma(-t, h, 0.5);
    y = fma(y, e, y);
  }
  return __hiloint2double(__double2hiint(y) - ((e2 >> 1) << 20), __double2loint(y));   // y * 2^-(e2/2)
}

// Blocked right-looking Cholesky (block 16) of the augmented matrix [A ; r^T]: carrying r as one extra
// row through the panel/update steps leaves z = L^-1 r in that row, so no forward substitution is needed.
// fp64 arithmetic on this part has ~25-cycle dependent latency, so every phase is written to keep the
// dependent chains short:
//   (1) diagonal block: one warp, rows in registers, pivots/multipliers by shuffle, no selects (the part of
//       a row right of the diagonal may hold garbage -- it is never read), rsqrt_pos instead of sqrt/div;
//   (2) panel: one thread per row, "right-looking" inside the row (2 dependent ops per column);
//   (3) trailing update: 4 independent accumulators per thread;
//   (4) back substitution per block in registers + shuffles.
// A is (d+1) x (d+1) with row pitch d+1 (fp64, shared memory); row d = r^T.
constexpr int kNB = 16;
constexpr int kCholThreads = 512;

__global__ void __launch_bounds__(kCholThreads, 1)
solve_cholesky_kernel(const double* __restrict__ S, int d, double alpha, int fit_intercept,
                      double* __restrict__ out) {
  extern __shared__ double sm[];
  const int pitch = d + 1;
  double* A = sm;                        // rows 0..d-1 = A, row d = r^T
  double* r = A + d * pitch;             // alias of row d
  double* mean = A + (d + 1) * pitch;    // d
  double* invd = mean + d;               // d: 1 / L[k][k]
  double* misc = invd + d;               // [0] ybar, [1] max diag, [2] info (1-based failing pivot, 0 = ok)
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  long long tm[5] = {0, 0, 0, 0, 0};     // phase cycle counters: build, diag, panel, update, backward
  long long tc0 = clock64();
  build_normal_equations(S, d, alpha, fit_intercept, A, r, mean, &misc[0]);
  tm[0] = clock64() - tc0;
  if (warp == 0) {
    double mx = 0.0;
    for (int i = lane; i < d; i += 32) mx = fmax(mx, A[i * pitch + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) { misc[1] = mx; misc[2] = 0.0; }
  }
  __syncthreads();
  const double tiny = misc[1] * 1e-12;
  const int rows = d + 1;                // including the augmented row

  for (int kb = 0; kb < d; kb += kNB) {
    const int nb = (d - kb) < kNB ? (d - kb) : kNB;
    tc0 = clock64();
    // ---- (1) diagonal block -------------------------------------------------------------------------
    if (warp == 0) {
      if (nb == kNB) {
        double a[kNB];
        const int row = kb + (lane & (kNB - 1));
#pragma unroll
        for (int c = 0; c < kNB; ++c) a[c] = lane < kNB ? A[row * pitch + kb + c] : 0.0;   // only c <= lane is meaningful
        double my_inv = 1.0;
        int first_bad = 0;
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
          const double piv = __shfl_sync(0xffffffffu, a[k], k);
          const bool bad = !(piv > tiny);
          const double inv = bad ? 1.0 : rsqrt_pos(piv);
          const double l = a[k] * inv;                  // lane > k: L[r][k]; lane == k: sqrt(piv)
          a[k] = l;
          my_inv = (lane == k) ? inv : my_inv;
          first_bad = (bad && first_bad == 0) ? (kb + k + 1) : first_bad;
#pragma unroll
          for (int c = k + 1; c < kNB; ++c) a[c] = fma(-l, __shfl_sync(0xffffffffu, l, c), a[c]);
        }
        __syncwarp();
        if (lane < kNB) {
#pragma unroll
          for (int c = 0; c < kNB; ++c) if (c <= lane) A[row * pitch + kb + c] = a[c];
          invd[kb + lane] = my_inv;
        }
        if (lane == 0 && first_bad != 0 && misc[2] == 0.0) misc[2] = (double)first_bad;
      } else {   // ragged last block (d % 16 != 0): plain shared-memory version
        const int row = kb + lane;
        for (int k = 0; k < nb; ++k) {
          const int kk = kb + k;
          const double piv = A[kk * pitch + kk];
          __syncwarp();
          const bool bad = !(piv > tiny);
          const double inv = bad ? 1.0 : rsqrt_pos(piv);
          if (lane == k) {
            if (bad && misc[2] == 0.0) misc[2] = (double)(kk + 1);
            A[kk * pitch + kk] = bad ? 1.0 : piv * inv;
            invd[kk] = inv;
          }
          double l = 0.0;
          if (lane > k && lane < nb) {
            l = A[row * pitch + kk] * inv;
            A[row * pitch + kk] = l;
          }
          __syncwarp();
          if (lane > k && lane < nb)
            for (int c = k + 1; c <= lane; ++c) A[row * pitch + kb + c] -= l * A[(kb + c) * pitch + kk];
          __syncwarp();
        }
      }
    }
    __syncthreads();
    tm[1] += clock64() - tc0; tc0 = clock64();
    if (misc[2] != 0.0) break;
    // ---- (2) panel: rows below the block (incl. the r row): x L_bb^T = a ----------------------------------
    const int below = rows - kb - nb;
    for (int t = tid; t < below; t += blockDim.x) {
      const int i = kb + nb + t;
      double sv[kNB];
#pragma unroll
      for (int c = 0; c < kNB; ++c) sv[c] = c < nb ? A[i * pitch + kb + c] : 0.0;
#pragma unroll
      for (int m = 0; m < kNB; ++m) {
        if (m < nb) {
          const double xm = sv[m] * invd[kb + m];
          sv[m] = xm;
#pragma unroll
          for (int c = m + 1; c < kNB; ++c) if (c < nb) sv[c] = fma(-xm, A[(kb + c) * pitch + kb + m], sv[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < kNB; ++c) if (c < nb) A[i * pitch + kb + c] = sv[c];
    }
    __syncthreads();
    tm[2] += clock64() - tc0; tc0 = clock64();
    // ---- (3) trailing update A[i][j] -= sum_m P[i][m] P[j][m], i >= j >= kb+nb (i up to the r row) -----------
    // each thread owns a 2-row x 4-column register tile: 8 independent fp64 chains, one shared-memory load per
    // two FMAs (the panel rows P[i][.] stay in registers)
    const int ty = tid >> 4, tx = tid & 15;           // 32 x 16 thread grid
    const int base = kb + nb;
    for (int i0 = base + ty; i0 < rows; i0 += 64) {
      const int i1 = i0 + 32;
      const bool has1 = i1 < rows;
      double p0[kNB], p1[kNB];
#pragma unroll
      for (int m = 0; m < kNB; ++m) {
        p0[m] = m < nb ? A[i0 * pitch + kb + m] : 0.0;
        p1[m] = (m < nb && has1) ? A[i1 * pitch + kb + m] : 0.0;
      }
      const int jmax0 = i0 < d ? i0 : d - 1;
      const int jmax1 = has1 ? (i1 < d ? i1 : d - 1) : -1;
      const int jmax = jmax1 > jmax0 ? jmax1 : jmax0;
      for (int j0 = base + tx; j0 <= jmax; j0 += 64) {
        double a0[4] = {0.0, 0.0, 0.0, 0.0}, a1[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 0; m < kNB; ++m) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + 16 * u;
            const double pj = (m < nb && j <= jmax) ? A[j * pitch + kb + m] : 0.0;
            a0[u] = fma(p0[m], pj, a0[u]);
            a1[u] = fma(p1[m], pj, a1[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + 16 * u;
          if (j <= jmax0) A[i0 * pitch + j] -= a0[u];
          if (j <= jmax1) A[i1 * pitch + j] -= a1[u];
        }
      }
    }
    __syncthreads();
    tm[3] += clock64() - tc0;
  }
  tc0 = clock64();
  const bool singular = misc[2] != 0.0;
  if (!singular) {
    // row d now holds z = L^-1 r.  backward: L^T b = z, blocked from the bottom
    for (int kb = ((d - 1) / kNB) * kNB; kb >= 0; kb -= kNB) {
      const int nb = (d - kb) < kNB ? (d - kb) : kNB;
      if (warp == 0) {
        if (nb == kNB) {
          const int col = lane & (kNB - 1);
          double lt[kNB];                                // lt[k] = L[kb+k][kb+col]
#pragma unroll
          for (int k = 0; k < kNB; ++k) lt[k] = lane < kNB ? A[(kb + k) * pitch + kb + col] : 0.0;
          double z = lane < kNB ? r[kb + col] : 0.0;
          const double dinv = lane < kNB ? invd[kb + col] : 0.0;
#pragma unroll
          for (int k = kNB - 1; k >= 0; --k) {
            const double bk = __shfl_sync(0xffffffffu, z * dinv, k);
            z = (lane == k) ? bk : ((lane < k) ? fma(-lt[k], bk, z) : z);   // lanes > k are already final
          }
          __syncwarp();
          if (lane < kNB) r[kb + lane] = z;
        } else {
          for (int k = nb - 1; k >= 0; --k) {
            if (lane == k) r[kb + k] *= invd[kb + k];
            __syncwarp();
            if (lane < k) r[kb + lane] -= A[(kb + k) * pitch + kb + lane] * r[kb + k];
            __syncwarp();
          }
        }
      }
      __syncthreads();
      for (int i = tid; i < kb; i += blockDim.x) {
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int m = 0; m < kNB; m += 2) {
          if (m < nb) acc0 = fma(A[(kb + m) * pitch + i], r[kb + m], acc0);
          if (m + 1 < nb) acc1 = fma(A[(kb + m + 1) * pitch + i], r[kb + m + 1], acc1);
        }
        r[i] -= acc0 + acc1;
      }
      __syncthreads();
    }
  }
  tm[4] = clock64() - tc0;
  if (tid == 0)
    for (int k = 0; k < 5; ++k) out[kOutSingular + k] = (double)tm[k];
  for (int i = tid; i < d; i += blockDim.x) out[i] = singular ? 0.0 : r[i];
  if (warp == 0) {
    double part = 0.0;
    for (int i = lane; i < d; i += 32) part += mean[i] * r[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) {
      out[kOutIntercept] = singular ? 0.0 : misc[0] - part;
      out[kOutInfo] = misc[2];
    }
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

size_t solve_smem_bytes(int d) { return sizeof(double) * ((size_t)(d + 1) * (d + 1) + 3 * d + 8); }

}  // namespace

int launch_solve_cholesky(b2_ctx* ctx, double alpha, int fit_intercept) {
  const size_t smem = solve_smem_bytes(ctx->d);
  if (!ctx->solve_attr_set) {
    B2_CUDA(cudaFuncSetAttribute(solve_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(kMaxD)));
    B2_CUDA(cudaFuncSetAttribute(solve_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(kMaxD)));
    ctx->solve_attr_set = true;
  }
  solve_cholesky_kernel<<<1, kCholThreads, smem, ctx->stream>>>(ctx->S, ctx->d, alpha, fit_intercept,
                                                                 ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

int launch_solve_spectral(b2_ctx* ctx, double cond, int fit_intercept) {
  const size_t smem = solve_smem_bytes(ctx->d);
  if (!ctx->solve_attr_set) {
    B2_CUDA(cudaFuncSetAttribute(solve_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(kMaxD)));
    B2_CUDA(cudaFuncSetAttribute(solve_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)solve_smem_bytes(kMaxD)));
    ctx->solve_attr_set = true;
  }
  solve_spectral_kernel<<<1, 512, smem, ctx->stream>>>(ctx->S, ctx->d, cond, fit_intercept, ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

}  // namespace b2
