// ubench_fp64.cu -- dependent-chain latencies that pace the single-SM solve (solve.cu): fp64 FMA / MUL, a 64-bit
// shuffle, the fp32-seeded reciprocal, a shared-memory load.  One warp, clock64() around N dependent operations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

__device__ __forceinline__ double rcp_pos(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double er = fma(-x, y, 1.0);
  const double t = fma(er, er, er);
  return fma(y, t, y);
}

__global__ void bench(double a, double b, double* out, long long* cyc) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a + threadIdx.x;
  __syncthreads();
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  cyc[0] = t1 - t0;
  double y = a;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = y * b;
  t1 = clock64();
  cyc[1] = t1 - t0;
  double z = a + threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = __shfl_sync(0xffffffffu, z, (i + 1) & 31);
  t1 = clock64();
  cyc[2] = t1 - t0;
  double r = a + 2.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) r = rcp_pos(r) + 1.5;
  t1 = clock64();
  cyc[3] = t1 - t0;
  double d = a + 3.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) d = 1.0 / d + 1.5;
  t1 = clock64();
  cyc[4] = t1 - t0;
  float f = (float)a;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) f = fmaf(f, (float)b, (float)a);
  t1 = clock64();
  cyc[5] = t1 - t0;
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = (int)sm[idx & 31] & 31;
  t1 = clock64();
  cyc[6] = t1 - t0;
  double q = a + 4.0;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) q = sqrt(q) + 1.5;
  t1 = clock64();
  cyc[7] = t1 - t0;
  out[threadIdx.x] = x + y + z + r + d + f + idx + q;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * sizeof(double));
  cudaMallocManaged(&cyc, 8 * sizeof(long long));
  for (int rep = 0; rep < 2; ++rep) { bench<<<1, 32>>>(1.0000001, 0.9999999, out, cyc); cudaDeviceSynchronize(); }
  const char* names[8] = {"DFMA", "DMUL", "SHFL.64", "rcp_pos+DADD", "1.0/x+DADD", "FFMA", "LDS.64+cvt", "sqrt+DADD"};
  for (int k = 0; k < 8; ++k) printf("%-14s %7.1f cycles per dependent op\n", names[k], (double)cyc[k] / N);
  return 0;
}
