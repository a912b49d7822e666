/* A plain C99 client of the drop-in boundary: nothing but include/b2gram.h and libb2gram.so.
 *
 * It does what the reference's retrain does around its three scikit-learn calls
 * (mlops_simulation/stage_1_train_model.py:93-108): hold rows in ordinary host memory, split, fit, score the
 * hold-out rows, read the coefficients and metrics back -- here through the C-ABI a non-Python caller would bind
 * (b2_gram_* + b2_solve, the one-call b2_fit, b2_solve_eigvals, b2_split_mask, b2_score).  The check value is a
 * double-precision normal-equation solve written out below (3 features, so Cramer-free Gaussian elimination).
 *
 * exit 0: fitted coefficients agree to 1e-4;  exit 3: no CUDA device (the library has no CPU path);
 * exit 1: anything else.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "b2gram.h"

#define N 20000
#define D 3

static uint64_t lcg_state = 88172645463325252ull;
static double uniform01(void) {
  lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(lcg_state >> 11) / 9007199254740992.0;
}

/* solve the (D+1) x (D+1) normal equations [X 1]^T [X 1] b = [X 1]^T y in double */
static void reference_fit(const float* X, const float* y, double* coef, double* intercept) {
  double A[D + 1][D + 2];
  int i, j, k, r;
  for (i = 0; i <= D; ++i)
    for (j = 0; j <= D + 1; ++j) A[i][j] = 0.0;
  for (r = 0; r < N; ++r) {
    double row[D + 1];
    for (j = 0; j < D; ++j) row[j] = (double)X[r * D + j];
    row[D] = 1.0;
    for (i = 0; i <= D; ++i) {
      for (j = 0; j <= D; ++j) A[i][j] += row[i] * row[j];
      A[i][D + 1] += row[i] * (double)y[r];
    }
  }
  for (k = 0; k <= D; ++k) {
    int p = k;
    for (i = k + 1; i <= D; ++i)
      if (fabs(A[i][k]) > fabs(A[p][k])) p = i;
    for (j = 0; j <= D + 1; ++j) { double t = A[k][j]; A[k][j] = A[p][j]; A[p][j] = t; }
    for (i = k + 1; i <= D; ++i) {
      double f = A[i][k] / A[k][k];
      for (j = k; j <= D + 1; ++j) A[i][j] -= f * A[k][j];
    }
  }
  {
    double b[D + 1];
    for (i = D; i >= 0; --i) {
      double s = A[i][D + 1];
      for (j = i + 1; j <= D; ++j) s -= A[i][j] * b[j];
      b[i] = s / A[i][i];
    }
    for (j = 0; j < D; ++j) coef[j] = b[j];
    *intercept = b[D];
  }
}

int main(void) {
  int n_dev = 0, rc, j;
  b2_ctx* ctx = NULL;
  float* X = (float*)malloc(sizeof(float) * N * D);
  float* y = (float*)malloc(sizeof(float) * N);
  double coef[D], intercept = 0.0, want[D], want_b0 = 0.0, worst = 0.0;
  if (X == NULL || y == NULL) return 1;
  printf("libb2gram ABI version %d\n", b2_abi_version());
  if (b2_device_count(&n_dev) != B2_OK || n_dev == 0) {
    rc = b2_ctx_create(0, &ctx);
    printf("no CUDA device: b2_ctx_create -> %d (%s)\n", rc, b2_last_error());
    return rc == B2_OK ? 1 : 3;
  }
  for (j = 0; j < N; ++j) {
    double s = 1.0;
    int k;
    for (k = 0; k < D; ++k) {
      X[j * D + k] = (float)(100.0 * uniform01());
      s += (0.5 + 0.25 * k) * (double)X[j * D + k];
    }
    y[j] = (float)(s + 10.0 * (uniform01() + uniform01() + uniform01() - 1.5));
  }
  reference_fit(X, y, want, &want_b0);

  if ((rc = b2_ctx_create(0, &ctx)) != B2_OK) goto fail;
  if ((rc = b2_gram_reset(ctx, D)) != B2_OK) goto fail;
  /* rows live in pageable host memory: the library streams them to the device itself */
  if ((rc = b2_gram_accumulate(ctx, X, B2_F32, y, N, D, D, B2_MEM_HOST, NULL, 1)) != B2_OK) goto fail;
  if ((rc = b2_solve(ctx, 0.0, 1, coef, &intercept)) != B2_OK) goto fail;
  for (j = 0; j < D; ++j) {
    double e = fabs(coef[j] - want[j]);
    if (e > worst) worst = e;
    printf("coef[%d] = %.9f (want %.9f)\n", j, coef[j], want[j]);
  }
  printf("intercept = %.6f (want %.6f); worst coefficient error %.3g\n", intercept, want_b0, worst);
  {
    /* the same fit through the one-call entry point, and the spectrum attributes of the estimator */
    double coef2[D], b2 = 0.0, sing[D];
    int rank = 0;
    int64_t rows = 0;
    if ((rc = b2_fit(ctx, X, B2_F32, y, N, D, D, B2_MEM_HOST, NULL, 1, 0.0, 1, coef2, &b2)) != B2_OK) goto fail;
    for (j = 0; j < D; ++j)
      if (coef2[j] != coef[j]) { fprintf(stderr, "b2_fit differs from the four-call sequence\n"); return 1; }
    if (b2 != intercept) return 1;
    if ((rc = b2_solve_eigvals(ctx, 1e-6, 1, sing, &rank, &rows)) != B2_OK) goto fail;
    printf("rank %d, rows %lld, singular values %.4f .. %.4f\n", rank, (long long)rows, sing[0], sing[D - 1]);
    if (rank != D || rows != N || !(sing[0] >= sing[D - 1] && sing[D - 1] > 0.0)) return 1;
  }
  {
    /* train_test_split's membership (random_state = 42, test_size = 0.2) as a row mask: masked fit + hold-out metrics */
    uint8_t* mask = (uint8_t*)malloc(N);
    double stats[10], coef3[D], b3 = 0.0;
    int64_t n_test = (N + 4) / 5, k, zeros = 0;
    if (mask == NULL) return 1;
    if ((rc = b2_split_mask(N, n_test, 42u, mask)) != B2_OK) goto fail;
    for (k = 0; k < N; ++k) zeros += mask[k] == 0;
    if (zeros != n_test) { fprintf(stderr, "split mask holds %lld test rows, want %lld\n", (long long)zeros, (long long)n_test); return 1; }
    if ((rc = b2_fit(ctx, X, B2_F32, y, N, D, D, B2_MEM_HOST, mask, 1, 0.0, 1, coef3, &b3)) != B2_OK) goto fail;
    if ((rc = b2_score(ctx, X, B2_F32, N, D, D, B2_MEM_HOST, coef3, b3, y, mask, 0, NULL, stats)) != B2_OK) goto fail;
    printf("hold-out rows %.0f, MAPE %.5f, max |residual| %.3f\n", stats[5], stats[0] / stats[5], stats[4]);
    if ((int64_t)stats[5] != n_test) return 1;
    free(mask);
  }
  b2_ctx_destroy(ctx);
  free(X);
  free(y);
  return (worst < 1e-4 && fabs(intercept - want_b0) < 1e-2) ? 0 : 1;
fail:
  fprintf(stderr, "libb2gram call failed: %d (%s)\n", rc, b2_last_error());
  if (ctx != NULL) b2_ctx_destroy(ctx);
  return 1;
}
