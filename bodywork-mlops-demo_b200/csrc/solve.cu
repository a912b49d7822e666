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
#include "b2_xchg.cuh"

namespace b2 {
namespace {

constexpr int kSolveThreads = 256;
constexpr int kOutIntercept = kMaxD;      // solve_out layout: [0,d) coef | intercept | info | rank | singular[d]
constexpr int kOutInfo = kMaxD + 1;
constexpr int kOutRank = kMaxD + 2;
constexpr int kOutSingular = kMaxD + 3;
constexpr int kOutRows = kOutSingular + kMaxD;   // eigvals kernel only

// Builds A (pitch d+1) and r in shared memory from the raw statistic; returns means.
__device__ void build_normal_equations(const double* S_in, int d, double alpha, int fit_intercept,
                                       double* A, double* r, double* mean, double* ybar_out) {
  const volatile double* S = S_in;   // the fused solve has just written S itself: plain loads, not the read-only path
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

// 1/x for normal positive x without the library's slow-path division (a call inside the unrolled pivot loop forces the
// register-resident block onto the stack, and a correctly rounded fp64 division is ~25 dependent instructions):
// scale x by a power of two into [1, 2), fp32 MUFU seed y0 (relative error e ~ 2^-22), one cubic step
// y = y0 (1 + e + e^2) (error e^3 ~ 2^-66, i.e. below the rounding of the last FMA), undo the scaling.
// 6 dependent instructions after the conversion; the pivot recurrence of the factorisation is paced by this chain.
__device__ __forceinline__ double rcp_pos(double x) {
  const int hi = __double2hiint(x), lo = __double2loint(x);
  const int e = ((hi >> 20) & 0x7ff) - 1023;
  const double xs = __hiloint2double(hi - (e << 20), lo);               // x * 2^-e in [1, 2)
  float y0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)xs));
  double y = (double)y0;
  const double er = fma(-xs, y, 1.0);
  const double t = fma(er, er, er);
  y = fma(y, t, y);
  return __hiloint2double(__double2hiint(y) - (e << 20), __double2loint(y));   // y * 2^-e
}

// Peer exchange consumed by the solve (fused fit, b2_fit): wait for the exchange, sum the slots into S.
struct SolveXchg {
  double* own;                  // nullptr: S is already complete
  int n_ranks;
  unsigned int epoch;
  unsigned long long timeout_ns;
};

// Blocked right-looking LDL^T (block 16, no square roots) of the augmented matrix [A ; r^T]:  A = M D M^T with M unit
// lower triangular.  Carrying r as one extra row through the panel / update steps leaves w = D^-1 M^-1 r in that row, so
// there is no forward substitution; the back substitution M^T b = w needs no division.  fp64 arithmetic here is
// latency bound (the whole solve is 0.7 MFLOP), so every phase is written to keep the dependent chains short:
//   (1) diagonal block: one warp, rows in registers, pivots by shuffle; the products u_ik u_ck are formed before the
//       reciprocal of the pivot arrives, so the recurrence pivot -> next pivot is shuffle + rcp_pos (6) + one FMA;
//   (2) panel: one thread per row, 2 dependent operations per column;
//   (3) trailing update A[i][j] -= sum_m M[i][m] U[j][m] (U = M D, the unscaled entries): 2 x 4 register tiles;
//   (4) back substitution per block in registers: shuffle + FMA per unknown.
// A is (d+1) x (d+1) with row pitch d+1 (fp64, shared memory); row d = r^T.  U: (d+1) x 16 panel scratch.
constexpr int kNB = 16;
constexpr int kCholThreads = 512;
constexpr int kUPitch = kNB + 1;

__global__ void __launch_bounds__(kCholThreads, 1)
solve_cholesky_kernel(double* S, int d, double alpha, int fit_intercept, double* __restrict__ out, const SolveXchg xc) {
  extern __shared__ double sm[];
  const int pitch = d + 1;
  double* A = sm;                        // rows 0..d-1 = A, row d = r^T
  double* r = A + d * pitch;             // alias of row d
  double* mean = A + (d + 1) * pitch;    // d
  double* invd = mean + d;               // d: 1 / D[k]
  double* misc = invd + d;               // [0] ybar, [1] max diag, [2] info (1-based failing pivot, 0 = ok)
  double* U = misc + 8;                  // (d+1) x kUPitch: unscaled panel entries of the current block column
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  long long tm[5] = {0, 0, 0, 0, 0};     // phase cycle counters: build, diag, panel, update, backward
  long long tc0 = clock64();
  if (xc.own != nullptr) {
    // fused fit: this rank's partial S went to every peer from the Gram kernel's fold; gather = wait + sum, here
    __shared__ int ok;
    if (tid == 0) ok = xchg_wait(xc.own, xc.n_ranks, xc.epoch, xc.timeout_ns) ? 1 : 0;
    __syncthreads();
    if (!ok) {
      if (tid == 0) {
        xchg_flags(xc.own)[kXchgStatusWord] = xc.epoch;
        out[kOutInfo] = -1.0;             // the host turns this into B2_E_COMM (never a fit on a partial statistic)
      }
      return;
    }
    const int dp = d + 2;
    for (int idx = tid; idx < dp * dp; idx += blockDim.x) S[idx] = xchg_sum(xc.own, xc.n_ranks, xc.epoch, idx);
    __threadfence();
    __syncthreads();
  }
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
    // ---- (1) diagonal block: rows kb .. kb+nb-1 in the registers of lanes 0 .. nb-1 -------------------------------
    if (warp == 0) {
      double a[kNB], mm[kNB];
      const int lrow = lane & (kNB - 1);
      const int row = kb + (lrow < nb ? lrow : 0);
      const bool act = lane < nb;
#pragma unroll
      for (int c = 0; c < kNB; ++c) { a[c] = (act && c < nb) ? A[row * pitch + kb + c] : 0.0; mm[c] = 0.0; }
      double my_rc = 1.0;
      int first_bad = 0;
#pragma unroll
      for (int k = 0; k < kNB; ++k) {
        if (k < nb) {
          const double piv = __shfl_sync(0xffffffffu, a[k], k);
          const bool bad = !(piv > tiny);
          const double rc = bad ? 1.0 : rcp_pos(piv);
          const double u = a[k];                        // lane > k: unscaled entry u_ik = m_ik * D_k
          mm[k] = u * rc;                               // m_ik
          my_rc = (lane == k) ? rc : my_rc;
          first_bad = (bad && first_bad == 0) ? (kb + k + 1) : first_bad;
#pragma unroll
          for (int c = k + 1; c < kNB; ++c) {
            const double uc = __shfl_sync(0xffffffffu, u, c);   // u_ck
            a[c] = fma(-(u * uc), rc, a[c]);                    // a_ic -= u_ik u_ck / D_k   (only c <= lane is meaningful)
          }
        }
      }
      __syncwarp();
      if (act) {
#pragma unroll
        for (int c = 0; c < kNB; ++c) {
          if (c < lane) { A[row * pitch + kb + c] = mm[c]; U[row * kUPitch + c] = a[c]; }
        }
        invd[kb + lane] = my_rc;
      }
      if (lane == 0 && first_bad != 0 && misc[2] == 0.0) misc[2] = (double)first_bad;
    }
    __syncthreads();
    tm[1] += clock64() - tc0; tc0 = clock64();
    if (misc[2] != 0.0) break;
    // ---- (2) panel: rows below the block (incl. the r row) ----------------------------------------------------------
    const int below = rows - kb - nb;
    for (int t = tid; t < below; t += blockDim.x) {
      const int i = kb + nb + t;
      double sv[kNB];
#pragma unroll
      for (int c = 0; c < kNB; ++c) sv[c] = c < nb ? A[i * pitch + kb + c] : 0.0;
#pragma unroll
      for (int m = 0; m < kNB; ++m) {
        if (m < nb) {
          const double um = sv[m];
          const double xm = um * invd[kb + m];
          A[i * pitch + kb + m] = xm;
          U[i * kUPitch + m] = um;
#pragma unroll
          for (int c = m + 1; c < kNB; ++c) if (c < nb) sv[c] = fma(-xm, U[(kb + c) * kUPitch + m], sv[c]);
        }
      }
    }
    __syncthreads();
    tm[2] += clock64() - tc0; tc0 = clock64();
    // ---- (3) trailing update A[i][j] -= sum_m M[i][m] U[j][m], i >= j >= kb+nb (i up to the r row) ---------------
    // each thread owns a 2-row x 4-column register tile: 8 independent fp64 chains, one shared-memory load per
    // two FMAs (the panel rows M[i][.] stay in registers)
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
            const double pj = (m < nb && j <= jmax) ? U[j * kUPitch + m] : 0.0;
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
    // row d now holds w = D^-1 M^-1 r.  backward: M^T b = w (unit diagonal), blocked from the bottom
    for (int kb = ((d - 1) / kNB) * kNB; kb >= 0; kb -= kNB) {
      const int nb = (d - kb) < kNB ? (d - kb) : kNB;
      if (warp == 0) {
        const int col = lane & (kNB - 1);
        const bool act = lane < nb;
        double lt[kNB];                                // lt[k] = M[kb+k][kb+col], k > col
#pragma unroll
        for (int k = 0; k < kNB; ++k) lt[k] = (act && k < nb && k > col) ? A[(kb + k) * pitch + kb + col] : 0.0;
        double z = act ? r[kb + col] : 0.0;
#pragma unroll
        for (int k = kNB - 1; k >= 0; --k) {
          if (k < nb) {
            const double bk = __shfl_sync(0xffffffffu, z, k);   // lane k is final: every i > k has been subtracted
            z = (lane < k) ? fma(-lt[k], bk, z) : z;
          }
        }
        __syncwarp();
        if (act) r[kb + lane] = z;
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

// ---- eigenvalues only: singular_ and rank_ of the fitted estimator -------------------------------------------------
// LinearRegression.fit stores `singular_` (singular values of the centred X, descending) and `rank_` next to the
// coefficients (sklearn/linear_model/_base.py: `self.coef_, _, self.rank_, self.singular_ = linalg.lstsq(...)`), and the
// joblib artefact of stage_1_train_model.py:113-114 carries them.  They are sqrt(eig(Xc^T Xc)), i.e. eigenvalues of the
// same centred Gram matrix the Cholesky solve factors -- no eigenvectors are needed unless the matrix is rank deficient
// (then solve_spectral_kernel computes the minimum-norm coefficients).  One CTA:
//   (a) Householder tridiagonalisation T = Q^T A Q in shared memory (d - 2 reflections; matvec + rank-2 update by all
//       threads, 3 block syncs per reflection);
//   (b) eigenvalues of T by multisection on Sturm counts: 4 threads per eigenvalue evaluate the division-free
//       characteristic-polynomial recurrence (one dependent FMA per row, power-of-two rescaling every 8 rows) at the 4
//       interior points of its bracket, so every round shrinks every bracket 5x with no block-level synchronisation.
constexpr int kEigThreads = 512;

__device__ __forceinline__ int sturm_count(const double* __restrict__ dd, const double* __restrict__ ee2, int d, double x) {
  // number of eigenvalues of T below x = sign changes of p_0 = 1, p_1 = d_0 - x, p_i = (d_{i-1} - x) p_{i-1} - e_{i-2}^2 p_{i-2}
  double pm = 1.0, p = dd[0] - x;
  int count = p < 0.0 ? 1 : 0;
  if (p == 0.0) { p = -1e-300; count = 1; }
  for (int i = 1; i < d; ++i) {
    double pn = fma(dd[i] - x, p, -(ee2[i - 1] * pm));
    if (pn == 0.0) pn = (p < 0.0) ? 1e-300 : -1e-300;               // a zero counts as a sign change
    count += ((pn < 0.0) != (p < 0.0)) ? 1 : 0;
    pm = p; p = pn;
    if ((i & 7) == 7) {                                              // keep the pair in range: scale both by 2^-exponent(p)
      const int eb = (__double2hiint(p) >> 20) & 0x7ff;
      int sh = 1023 - eb;
      sh = sh > 1000 ? 1000 : (sh < -1000 ? -1000 : sh);
      const double sc = __hiloint2double((1023 + sh) << 20, 0);
      p *= sc; pm *= sc;
    }
  }
  return count;
}

__global__ void __launch_bounds__(kEigThreads, 1)
solve_eigvals_kernel(const double* S, int d, double cond, int fit_intercept, double* __restrict__ out) {
  extern __shared__ double sm[];
  const int pitch = d + 1;
  double* A = sm;                         // symmetric, full storage (both triangles kept current)
  double* r = A + d * pitch;              // (unused here; build_normal_equations fills it)
  double* mean = A + (d + 1) * pitch;
  double* misc = mean + 2 * d;
  double* v = misc + 8;                   // Householder vector (v[k+1] = 1)
  double* pv = v + d;                     // tau * A v
  double* dd = pv + d;                    // diagonal of T
  double* ee2 = dd + d;                   // squared off-diagonal of T
  double* lam = ee2 + d;                  // eigenvalues, ascending
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  build_normal_equations(S, d, 0.0, fit_intercept, A, r, mean, &misc[0]);

  for (int k = 0; k + 2 < d; ++k) {
    const int m = d - k - 1;              // length of the column below the diagonal: rows k+1 .. d-1
    // (a1) reflection: x = A[k+1.., k] (read as the row k, A is symmetric); beta = -sign(x0) |x|, tau = (beta - x0) / beta
    if (warp == 0) {
      double s2 = 0.0;
      for (int i = lane + 1; i < m; i += 32) { const double x = A[k * pitch + k + 1 + i]; s2 = fma(x, x, s2); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      const double x0 = A[k * pitch + k + 1];
      double tau = 0.0, beta = x0, scale = 0.0;
      if (s2 > 0.0) {
        const double nrm = sqrt(fma(x0, x0, s2));
        beta = x0 >= 0.0 ? -nrm : nrm;
        tau = (beta - x0) / beta;
        scale = 1.0 / (x0 - beta);
      }
      for (int i = lane; i < m; i += 32) v[k + 1 + i] = (i == 0) ? 1.0 : A[k * pitch + k + 1 + i] * scale;
      if (lane == 0) { misc[1] = tau; dd[k] = A[k * pitch + k]; ee2[k] = beta * beta; }
    }
    __syncthreads();
    const double tau = misc[1];
    if (tau != 0.0) {                     // block-uniform
      // (a2) p = tau * A22 v: 4 threads per row
      {
        const int row = tid >> 2, q = tid & 3;
        double acc = 0.0;
        if (row < m) {
          const double* ar = A + (k + 1 + row) * pitch + k + 1;
          for (int j = q; j < m; j += 4) acc = fma(ar[j], v[k + 1 + j], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (row < m && q == 0) pv[k + 1 + row] = tau * acc;
      }
      __syncthreads();
      // (a3) K = -tau/2 (p . v), w = p + K v (every warp recomputes K: no extra sync); A22 -= v w^T + w v^T
      double dot = 0.0;
      for (int i = lane; i < m; i += 32) dot = fma(pv[k + 1 + i], v[k + 1 + i], dot);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const double K = -0.5 * tau * dot;
      {
        const int row = tid >> 2, q = tid & 3;
        if (row < m) {
          const double vi = v[k + 1 + row], wi = fma(K, vi, pv[k + 1 + row]);
          double* ar = A + (k + 1 + row) * pitch + k + 1;
          for (int j = q; j < m; j += 4) {
            const double vj = v[k + 1 + j], wj = fma(K, vj, pv[k + 1 + j]);
            ar[j] -= fma(vi, wj, wi * vj);
          }
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (d >= 2) { dd[d - 2] = A[(d - 2) * pitch + d - 2]; ee2[d - 2] = A[(d - 1) * pitch + d - 2] * A[(d - 1) * pitch + d - 2]; }
    dd[d - 1] = A[(d - 1) * pitch + d - 1];
    // Gershgorin interval, then a power-of-two scaling so that |d_i - x| <= 2 and e_i^2 <= 1 in the recurrence
    double glo = dd[0], ghi = dd[0];
    for (int i = 0; i < d; ++i) {
      const double rad = (i > 0 ? sqrt(ee2[i - 1]) : 0.0) + (i + 1 < d ? sqrt(ee2[i]) : 0.0);
      glo = fmin(glo, dd[i] - rad); ghi = fmax(ghi, dd[i] + rad);
    }
    const double span = fmax(fmax(fabs(glo), fabs(ghi)), 1e-300);
    int ex = ((__double2hiint(span) >> 20) & 0x7ff) - 1023 + 1;
    ex = ex > 1000 ? 1000 : (ex < -1000 ? -1000 : ex);
    misc[2] = __hiloint2double((1023 - ex) << 20, 0);   // 2^-ex
    misc[3] = __hiloint2double((1023 + ex) << 20, 0);   // 2^ex
    misc[4] = glo; misc[5] = ghi;
  }
  __syncthreads();
  const double sdown = misc[2], sup = misc[3];
  for (int i = tid; i < d; i += blockDim.x) { dd[i] *= sdown; if (i + 1 < d) ee2[i] *= sdown * sdown; }
  __syncthreads();
  // (b) multisection: quad (4 consecutive lanes) owns eigenvalue index e; bracket invariant count(lo) <= e < count(hi)
  for (int e0 = 0; e0 < d; e0 += kEigThreads / 4) {
    const int e = e0 + (tid >> 2), q = tid & 3;
    const bool live = e < d;
    double lo = misc[4] * sdown, hi = misc[5] * sdown;
    const double w0 = hi - lo;
    lo -= 1e-3 * w0 + 1e-300; hi += 1e-3 * w0 + 1e-300;
    for (int round = 0; round < 26; ++round) {
      const double step = (hi - lo) * 0.2;
      const double x = lo + step * (double)(q + 1);
      const int c = live ? sturm_count(dd, ee2, d, x) : 0;
      const bool below = c <= e;                       // x is still a lower bound of eigenvalue e
      // the 4 points are increasing in q: new lo = the last `below` point, new hi = the first non-`below` point
      const unsigned int quad_shift = (unsigned int)(lane & ~3);
      const unsigned int bal = (__ballot_sync(0xffffffffu, below) >> quad_shift) & 0xFu;
      const int nb = __popc(bal);                      // below is monotone in x: the first nb points are lower bounds
      const double nlo = nb > 0 ? lo + step * (double)nb : lo;
      const double nhi = nb < 4 ? lo + step * (double)(nb + 1) : hi;
      lo = nlo; hi = nhi;
    }
    if (live && q == 0) lam[e] = 0.5 * (lo + hi) * sup;
  }
  __syncthreads();
  // singular values descending, rank = #{s > cond * s_max}
  const double lmax = fmax(lam[d - 1], 0.0);
  const double smax = sqrt(lmax);
  int rk = 0;
  for (int i = tid; i < d; i += blockDim.x) {
    const double sv = sqrt(fmax(lam[d - 1 - i], 0.0));
    out[kOutSingular + i] = sv;
    rk += (sv > cond * smax) ? 1 : 0;
  }
  __shared__ int rank_total;
  if (tid == 0) rank_total = 0;
  __syncthreads();
  if (rk) atomicAdd(&rank_total, rk);
  __syncthreads();
  if (tid == 0) {
    out[kOutRank] = (double)rank_total;
    out[kOutInfo] = 0.0;
    out[kOutRows] = S[d * (d + 2) + d];        // rows in the statistic: singular_ has min(rows, d) entries
  }
}

size_t solve_smem_bytes(int d) {
  return sizeof(double) * ((size_t)(d + 1) * (d + 1) + 3 * d + 16 + (size_t)(d + 1) * kUPitch + 8 * d);
}

int ensure_solve_attrs(b2_ctx* ctx) {
  if (!ctx->solve_attr_set) {
    const int bytes = (int)solve_smem_bytes(kMaxD);
    B2_CUDA(cudaFuncSetAttribute(solve_cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    B2_CUDA(cudaFuncSetAttribute(solve_spectral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    B2_CUDA(cudaFuncSetAttribute(solve_eigvals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    ctx->solve_attr_set = true;
  }
  return B2_OK;
}

}  // namespace

// The Cholesky (LDL^T) kernel writes its result straight into the pinned host mirror (no D2H copy node): the caller
// synchronises the stream and reads ctx->solve_host.
int launch_solve_cholesky(b2_ctx* ctx, double alpha, int fit_intercept, unsigned int gather_epoch) {
  if (int r = ensure_solve_attrs(ctx)) return r;
  SolveXchg xc;
  xc.own = gather_epoch != 0 ? ctx->xchg : nullptr;
  xc.n_ranks = ctx->n_ranks;
  xc.epoch = gather_epoch;
  xc.timeout_ns = ctx->xchg_timeout_ns;
  solve_cholesky_kernel<<<1, kCholThreads, solve_smem_bytes(ctx->d), ctx->stream>>>(ctx->S, ctx->d, alpha, fit_intercept,
                                                                                    ctx->solve_host, xc);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

int launch_solve_spectral(b2_ctx* ctx, double cond, int fit_intercept) {
  if (int r = ensure_solve_attrs(ctx)) return r;
  solve_spectral_kernel<<<1, 512, solve_smem_bytes(ctx->d), ctx->stream>>>(ctx->S, ctx->d, cond, fit_intercept, ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

int launch_solve_eigvals(b2_ctx* ctx, double cond, int fit_intercept) {
  if (int r = ensure_solve_attrs(ctx)) return r;
  solve_eigvals_kernel<<<1, kEigThreads, solve_smem_bytes(ctx->d), ctx->stream>>>(ctx->S, ctx->d, cond, fit_intercept,
                                                                                  ctx->solve_out);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return B2_OK;
}

}  // namespace b2
