// gram_tc.cu -- the hot kernel: row-block streaming Gram accumulator on tcgen05 (sm_100a).
//
// Replaces the pass over the training rows inside LinearRegression.fit
// (stage_1_train_model.py:105-106 -> sklearn/linear_model/_base.py: centre + LAPACK gelsd).
//
// Data flow per CTA (persistent, one CTA per SM, contiguous range of 32-row tiles):
//
//   HBM --TMA(cp.async.bulk.tensor)--> smem raw tile [32 rows][D] (+ y, + row mask)
//       --transform warps: v = x - c (per-column shift), bf16 split v = hi + lo,
//         CUDA-core side sums  sum v, sum v*y', sum y', sum y'^2, row count  (fp32 -> fp64)
//       --> smem operand tile, K-major canonical layout (8x16B core matrices, no swizzle)
//       --tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M=128 N=256 K=16:
//             D[i][j]      += sum_r hi[r][i] * hi[r][j]        (columns   0..127)
//             D[i][128+j]  += sum_r hi[r][i] * lo[r][j]        (columns 128..255)
//         accumulators live in TMEM (2 x 256 columns, double buffered)
//       --every `drain_rows` rows: epilogue warps tcgen05.ld the 128x256 fp32 block and fold it
//         into this CTA's fp64 partial in global memory (L2 resident).
//
// Why the shift and the split: the tensor core accumulates fp32 with truncation, so raw
// (uncentred) second moments cannot reach the 1e-4 coefficient tolerance; after the shift the
// Gram is ~diagonal and the centring in the solve subtracts almost nothing.  hi+lo carries
// 16 mantissa bits, i.e. products are accurate to ~2^-17 relative (lo*lo is dropped).
//
// The finalize kernels reduce the per-CTA partials in a fixed order (deterministic), undo the
// shift in fp64 and add the result to the context's raw statistic S = [X 1 y]^T [X 1 y].
#include <cuda_bf16.h>

#include "b2_internal.cuh"

namespace b2 {
namespace {

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
constexpr int kRawStages = 5;
constexpr int kOpStages = 3;
constexpr int kThreads = 512;           // 16 warps: 0 TMA, 1 MMA(+TMEM alloc), 2-3 idle, 4-7 epilogue, 8-15 transform
constexpr uint32_t kRawStageBytes = 16384;  // 32 rows x 128 fp32 (max)
constexpr uint32_t kOpLBO = 4096;       // bytes between the 8-row K groups (core matrices along K)
constexpr uint32_t kOpSBO = 128;        // bytes between 8-feature groups (core matrices along M/N)
constexpr uint32_t kOpStageBytes = (kTcRows / 8) * kOpLBO;  // 16384: [4 kgroups][256 j][8 k] bf16
constexpr uint32_t kOffRaw = 0;
constexpr uint32_t kOffOp = kOffRaw + kRawStages * kRawStageBytes;          // 81920
constexpr uint32_t kOffY = kOffOp + kOpStages * kOpStageBytes;              // 131072
constexpr uint32_t kOffMask = kOffY + kRawStages * 128;                     // 131712
constexpr uint32_t kOffBar = kOffMask + kRawStages * 128;                   // 132352
constexpr int kNumBars = 2 * kRawStages + 2 * kOpStages + 4;                // 20
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;                    // 132512
constexpr uint32_t kOffShift = kOffTmemPtr + 16;                            // 132528
constexpr uint32_t kSmemBytes = kOffShift + (kMaxD + 4) * 4 + 1024;         // + alignment slack

// instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) |
                            ((uint32_t)(kTcM >> 4) << 24);

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded spin: a protocol bug must end in a trap (clean launch failure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 24)) {
      if (err != nullptr) {
        *reinterpret_cast<volatile int*>(err) = code | (blockIdx.x << 8);
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;  // L2 cache hint: streaming data, read once

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const CUtensorMap* tm, int c0, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3}], [%2], %4;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(kOpLBO >> 4) << 16) |
         ((uint64_t)(kOpSBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint32_t (&v)[4]) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}

template <typename T>
__device__ __forceinline__ float raw_ld(const T* p);
template <>
__device__ __forceinline__ float raw_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float raw_ld<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __uint_as_float(((uint32_t) * reinterpret_cast<const unsigned short*>(p)) << 16);
}

// ------------------------------------------------------------------------------------------
// per-column shift c: mean of a strided row sample (any value near the column mean will do;
// the algebra in tc_fold_kernel is exact for every c)
// ------------------------------------------------------------------------------------------
constexpr int kShiftBlocks = 64;                 // partial sums of the row sample, one per block
constexpr int kShiftStride = kMaxD + 1;          // floats per partial: features, then y (slot kMaxD)

__host__ __device__ __forceinline__ int64_t shift_samples(int64_t n) { return n < 2048 ? n : 2048; }

// c_j from the 64 partial sums; bf16-representable so that (bf16 input - c) is exact in fp32.
// Called with identical arguments by the Gram kernel and by tc_fold_kernel -> identical c.
__device__ __forceinline__ float shift_value(const float* __restrict__ sp, int j, int64_t n) {
  float acc = 0.f;
#pragma unroll 8
  for (int b = 0; b < kShiftBlocks; ++b) acc += sp[b * kShiftStride + j];
  return __bfloat162float(__float2bfloat16_rn(acc / (float)shift_samples(n)));
}

template <typename T>
__global__ void tc_shift_kernel(const T* __restrict__ X, const float* __restrict__ y, int64_t n, int d,
                                int64_t ldx, float* __restrict__ sp) {
  const int j = threadIdx.x;
  if (j > d) return;
  const int64_t samples = shift_samples(n);
  const int64_t stride = n / samples;
  const int64_t per = (samples + kShiftBlocks - 1) / kShiftBlocks;
  const int64_t s0 = blockIdx.x * per;
  const int64_t s1 = (s0 + per < samples) ? s0 + per : samples;
  float acc = 0.f;
#pragma unroll 8
  for (int64_t s = s0; s < s1; ++s) {
    const int64_t row = s * stride;
    acc += (j < d) ? raw_ld<T>(X + row * ldx + j) : __ldg(y + row);
  }
  sp[blockIdx.x * kShiftStride + (j == d ? kMaxD : j)] = acc;
}

// ------------------------------------------------------------------------------------------
// the Gram kernel
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmM, int y_map_2d, int has_mask, int keep,
               int64_t n_rows, int d, const float* __restrict__ shift, int chunk_tiles,
               double* __restrict__ part, double* __restrict__ side, int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t bar_raw_full = sbase + kOffBar;                       // [kRawStages]
  const uint32_t bar_raw_empty = bar_raw_full + 8 * kRawStages;        // [kRawStages]
  const uint32_t bar_op_full = bar_raw_empty + 8 * kRawStages;         // [kOpStages]
  const uint32_t bar_op_empty = bar_op_full + 8 * kOpStages;           // [kOpStages]
  const uint32_t bar_acc_full = bar_op_empty + 8 * kOpStages;          // [2]
  const uint32_t bar_acc_empty = bar_acc_full + 16;                    // [2]
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);
  float* shift_s = reinterpret_cast<float*>(smem + kOffShift);

  // contiguous tile range of this CTA
  const int64_t total_tiles = (n_rows + kTcRows - 1) / kTcRows;
  const int64_t tile_begin = (int64_t)blockIdx.x * total_tiles / gridDim.x;
  const int64_t tile_end = (int64_t)(blockIdx.x + 1) * total_tiles / gridDim.x;
  const int my_tiles = (int)(tile_end - tile_begin);
  const int n_chunks = (my_tiles + chunk_tiles - 1) / chunk_tiles;

  // ---- one-time setup --------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRawStages; ++s) {
      mbar_init(bar_raw_full + 8 * s, 1);
      mbar_init(bar_raw_empty + 8 * s, kTcXformWarps);
    }
    for (int s = 0; s < kOpStages; ++s) {
      mbar_init(bar_op_full + 8 * s, kTcXformWarps);
      mbar_init(bar_op_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_acc_full + 8 * b, 1);
      mbar_init(bar_acc_empty + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
    if (has_mask) tma_prefetch_desc(&tmM);
  }
  if (warp == 1) {  // TMEM: all 512 columns (two 128x256 fp32 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sbase + kOffTmemPtr)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand stages once: feature columns >= d are never written and must read as 0
  for (uint32_t o = threadIdx.x * 16; o < kOpStages * kOpStageBytes; o += kThreads * 16)
    *reinterpret_cast<uint4*>(smem + kOffOp + o) = make_uint4(0, 0, 0, 0);
  for (int j = threadIdx.x; j <= kMaxD; j += kThreads)
    shift_s[j] = (j < d || j == kMaxD) ? shift_value(shift, j, n_rows) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // ---- warp roles --------------------------------------------------------------------------
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t x_bytes = (uint32_t)(kTcRows * d * sizeof(T));
      const uint32_t tx = x_bytes + kTcRows * 4 + (has_mask ? kTcRows : 0);
      for (int it = 0; it < my_tiles; ++it) {
        const int s = it % kRawStages;
        const uint32_t ph = (it / kRawStages) & 1;
        mbar_wait(bar_raw_empty + 8 * s, ph ^ 1, err, 1);
        const uint32_t full = bar_raw_full + 8 * s;
        mbar_expect_tx(full, tx);
        const int64_t row0 = (tile_begin + it) * kTcRows;
        tma_load_2d(sbase + kOffRaw + s * kRawStageBytes, &tmX, 0, (int)row0, full);
        if (y_map_2d) tma_load_2d(sbase + kOffY + s * 128, &tmY, 0, (int)(row0 >> 2), full);
        else tma_load_1d(sbase + kOffY + s * 128, &tmY, (int)row0, full);
        if (has_mask) tma_load_1d(sbase + kOffMask + s * 128, &tmM, (int)row0, full);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      for (int it = 0; it < my_tiles; ++it) {
        const int os = it % kOpStages;
        const uint32_t oph = (it / kOpStages) & 1;
        const int chunk = it / chunk_tiles;
        const int in_chunk = it - chunk * chunk_tiles;
        const int b = chunk & 1;
        if (in_chunk == 0) {  // first tile of a chunk: the TMEM buffer must have been drained
          mbar_wait(bar_acc_empty + 8 * b, ((chunk >> 1) & 1) ^ 1, err, 2);
          tc_fence_after();
        }
        mbar_wait(bar_op_full + 8 * os, oph, err, 3);
        tc_fence_after();
        const uint32_t op_addr = sbase + kOffOp + os * kOpStageBytes;
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * kTcN);
#pragma unroll
        for (int k2 = 0; k2 < kTcRows / 16; ++k2) {
          const uint64_t desc = make_smem_desc(op_addr + k2 * 2 * kOpLBO);
          umma_bf16(tmem_d, desc, desc, (in_chunk > 0 || k2 > 0) ? 1u : 0u);
        }
        umma_commit(bar_op_empty + 8 * os);  // frees the operand stage when these MMAs retire
        if (in_chunk == chunk_tiles - 1 || it == my_tiles - 1) umma_commit(bar_acc_full + 8 * b);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue: TMEM -> registers -> fp64 partial in global (column-major [col][feature]) =====
    const int w = warp & 3;  // TMEM lane quadrant this warp may access
    double* my_part = part + (size_t)blockIdx.x * kTcAccElems + w * 32 + lane;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
      const int b = chunk & 1;
      mbar_wait(bar_acc_full + 8 * b, (chunk >> 1) & 1, err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(b * kTcN);
#pragma unroll 1
      for (int p = 0; p < kTcN / 32; ++p) {
        uint32_t r[32];
        tmem_ld32(taddr + p * 32, r);
        tmem_ld_wait();
        double* dst = my_part + (size_t)(p * 32) * kTcM;
        if (chunk == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * kTcM] = (double)__uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[(size_t)j * kTcM] += (double)__uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * b);
    }
  } else if (warp >= 8) {
    // ===== transform: shift, bf16 hi/lo split, K-major operand store, CUDA-core side sums =====
    const int t = warp - 8;
    const int q = t & 3;        // feature quad: features q*32 .. q*32+31
    const int gsel = t >> 2;    // row groups {gsel, gsel+2} of each tile
    const int i = q * 32 + lane;
    const bool active = i < d;
    const float c_i = shift_s[active ? i : 0] * (active ? 1.f : 0.f);
    const float c_y = shift_s[kMaxD];
    double s1 = 0.0, sxy = 0.0, sy = 0.0, syy = 0.0;
    long long cnt = 0;
    const uint32_t st_off = (uint32_t)((i >> 3) * kOpSBO + (i & 7) * 16);
    for (int it = 0; it < my_tiles; ++it) {
      const int rs = it % kRawStages;
      const uint32_t rph = (it / kRawStages) & 1;
      const int os = it % kOpStages;
      const uint32_t oph = (it / kOpStages) & 1;
      mbar_wait(bar_raw_full + 8 * rs, rph, err, 5);
      mbar_wait(bar_op_empty + 8 * os, oph ^ 1, err, 6);
      tc_fence_after();
      const int64_t row0 = (tile_begin + it) * kTcRows;
      const bool full_tile = (!has_mask) && (row0 + kTcRows <= n_rows);
      const T* rawp = reinterpret_cast<const T*>(smem + kOffRaw + rs * kRawStageBytes);
      const float* yp = reinterpret_cast<const float*>(smem + kOffY + rs * 128);
      const uint8_t* mp = smem + kOffMask + rs * 128;
      const uint32_t op_addr = sbase + kOffOp + os * kOpStageBytes + st_off;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int gi = gsel + 2 * half;
        const int r0 = gi * 8;
        float v[8], yv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xv = active ? raw_ld<T>(rawp + (r0 + k) * d + i) : 0.f;
          v[k] = xv - c_i;
          yv[k] = yp[r0 + k] - c_y;
        }
        int used = 8;
        if (!full_tile) {
          used = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            bool use = (row0 + r0 + k) < n_rows;
            if (use && has_mask) use = (mp[r0 + k] == (uint8_t)keep);
            if (!use) { v[k] = 0.f; yv[k] = 0.f; }
            used += use ? 1 : 0;
          }
        }
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { t1 += v[k]; t2 = fmaf(v[k], yv[k], t2); }
        s1 += (double)t1;
        sxy += (double)t2;
        if (q == 0) {
          float a = 0.f, bb = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) { a += yv[k]; bb = fmaf(yv[k], yv[k], bb); }
          sy += (double)a;
          syy += (double)bb;
          cnt += used;
        }
        uint32_t hp[4], lp[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * p], v[2 * p + 1]);
          const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
          const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
          const __nv_bfloat162 l = __floats2bfloat162_rn(v[2 * p] - h0, v[2 * p + 1] - h1);
          hp[p] = hb;
          lp[p] = *reinterpret_cast<const uint32_t*>(&l);
        }
        if (active) {
          st_shared_v4(op_addr + gi * kOpLBO, hp);
          st_shared_v4(op_addr + gi * kOpLBO + (kTcM / 8) * kOpSBO, lp);
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_op_full + 8 * os);
        mbar_arrive(bar_raw_empty + 8 * rs);
      }
    }
    double* my_side = side + (size_t)blockIdx.x * kTcSideDoubles;
    my_side[(t * 32 + lane) * 2 + 0] = s1;
    my_side[(t * 32 + lane) * 2 + 1] = sxy;
    if (q == 0 && lane == 0) {
      double* ys = my_side + kTcXformWarps * 32 * 2 + gsel * 3;
      ys[0] = sy;
      ys[1] = syy;
      ys[2] = (double)cnt;
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// finalize 1: reduce the per-CTA partials in CTA order (deterministic)
//   red[0 .. 32768)            Gp[col][i]       (col 0..127: hi*hi, col 128..255: hi*lo)
//   red[32768 + i]             s1[i]  = sum (x_i - c_i)
//   red[32768 + 128 + i]       sxy[i] = sum (x_i - c_i)(y - c_y)
//   red[32768 + 256 + 0..2]    sum y', sum y'^2, rows used
// ------------------------------------------------------------------------------------------
__global__ void tc_reduce_kernel(const double* __restrict__ part, const double* __restrict__ side, int n_ctas,
                                 double* __restrict__ red) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < kTcAccElems) {
    double s = 0.0;
    for (int c = 0; c < n_ctas; ++c) s += part[(size_t)c * kTcAccElems + idx];
    red[idx] = s;
  } else if (idx < kTcAccElems + 2 * kMaxD) {
    const int k = idx - kTcAccElems;
    const int which = k / kMaxD, i = k % kMaxD;
    const int q = i >> 5, lane = i & 31;
    double s = 0.0;
    for (int c = 0; c < n_ctas; ++c) {
      const double* sd = side + (size_t)c * kTcSideDoubles;
      s += sd[((q)*32 + lane) * 2 + which] + sd[((q + 4) * 32 + lane) * 2 + which];
    }
    red[idx] = s;
  } else if (idx < kTcAccElems + 2 * kMaxD + 3) {
    const int k = idx - kTcAccElems - 2 * kMaxD;
    double s = 0.0;
    for (int c = 0; c < n_ctas; ++c) {
      const double* ys = side + (size_t)c * kTcSideDoubles + kTcXformWarps * 32 * 2;
      s += ys[k] + ys[3 + k];
    }
    red[idx] = s;
  }
}

// ------------------------------------------------------------------------------------------
// finalize 2: undo the shift in fp64 and add into the raw statistic S ((d+2)^2, row stride d+2)
// ------------------------------------------------------------------------------------------
__global__ void tc_fold_kernel(const double* __restrict__ red, const float* __restrict__ shift, int64_t n_rows,
                               int d, double* __restrict__ S) {
  const int dp = d + 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dp * dp) return;
  const int a = idx / dp, b = idx % dp;
  const double* s1 = red + kTcAccElems;
  const double* sxy = s1 + kMaxD;
  const double sy = red[kTcAccElems + 2 * kMaxD + 0];
  const double syy = red[kTcAccElems + 2 * kMaxD + 1];
  const double n = red[kTcAccElems + 2 * kMaxD + 2];
  const double cy = (double)shift_value(shift, kMaxD, n_rows);
  double val;
  if (a < d && b < d) {
    const double ca = (double)shift_value(shift, a, n_rows), cb = (double)shift_value(shift, b, n_rows);
    // G'(a,b) = sum (x_a-c_a)(x_b-c_b) ~= hh + hl + hl^T   (lo*lo dropped, ~2^-18 relative)
    const double hh = 0.5 * (red[(size_t)b * kTcM + a] + red[(size_t)a * kTcM + b]);
    const double hl = red[(size_t)(kTcM + b) * kTcM + a] + red[(size_t)(kTcM + a) * kTcM + b];
    val = hh + hl + ca * s1[b] + cb * s1[a] + n * ca * cb;
  } else if (a < d || b < d) {
    const int i = a < d ? a : b;
    const int o = a < d ? b : a;  // d (ones) or d+1 (y)
    const double ci = (double)shift_value(shift, i, n_rows);
    if (o == d) val = s1[i] + n * ci;
    else val = sxy[i] + cy * s1[i] + ci * sy + n * ci * cy;
  } else if (a == d && b == d) {
    val = n;
  } else if (a == d + 1 && b == d + 1) {
    val = syy + 2.0 * cy * sy + n * cy * cy;
  } else {
    val = sy + n * cy;
  }
  S[idx] += val;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

bool gram_tc_supported(const void* X, int x_dtype, const float* y, int64_t n, int d, int64_t ldx) {
  const int es = x_dtype == B2_F32 ? 4 : 2;
  if (d < 4 || d > kMaxD) return false;
  if ((d * es) % 16 != 0) return false;
  if ((ldx * es) % 16 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(X) & 15) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0) return false;
  if (n < kTcRows) return false;
  if (n > (int64_t)0x7fffffff) return false;  // TMA coordinates are int32
  return true;
}

int launch_gram_tc(b2_ctx* ctx, const void* X, int x_dtype, const float* y, int64_t n, int d, int64_t ldx,
                   const uint8_t* mask, int keep) {
  PFN_encodeTiled encode = get_encode();
  if (encode == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return B2_E_CUDA;
  }
  const int es = x_dtype == B2_F32 ? 4 : 2;
  CUtensorMap tmX, tmY, tmM;
  memset(&tmM, 0, sizeof(tmM));
  {
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * es};
    cuuint32_t box[2] = {(cuuint32_t)d, (cuuint32_t)kTcRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmX, x_dtype == B2_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                        2, const_cast<void*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(X) failed with %d (n=%lld d=%d ldx=%lld)", (int)r, (long long)n, d,
                (long long)ldx);
      return B2_E_CUDA;
    }
  }
  int y_map_2d = 0;
  {
    cuuint64_t dims[1] = {(cuuint64_t)n};
    cuuint64_t strides[1] = {0};
    cuuint32_t box[1] = {(cuuint32_t)kTcRows};
    cuuint32_t estr[1] = {1};
    CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, const_cast<float*>(y), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      // rank-1 maps refused: view y as [ceil(n/4)][4] (16-byte rows) -- same bytes land in smem
      cuuint64_t dims2[2] = {4, (cuuint64_t)((n + 3) / 4)};
      cuuint64_t strides2[1] = {16};
      cuuint32_t box2[2] = {4, (cuuint32_t)(kTcRows / 4)};
      cuuint32_t estr2[2] = {1, 1};
      r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(y), dims2, strides2, box2, estr2,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      y_map_2d = 1;
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(y) failed with %d", (int)r);
        return B2_E_CUDA;
      }
    }
  }
  const bool mask_tma = mask != nullptr && (reinterpret_cast<uintptr_t>(mask) & 15) == 0;
  if (mask != nullptr && !mask_tma) {
    set_error("row_mask must be 16-byte aligned for the tcgen05 path");
    return B2_E_ARG;
  }
  if (mask != nullptr) {
    cuuint64_t dims[1] = {(cuuint64_t)n};
    cuuint64_t strides[1] = {0};
    cuuint32_t box[1] = {(cuuint32_t)kTcRows};
    cuuint32_t estr[1] = {1};
    CUresult r = encode(&tmM, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, const_cast<uint8_t*>(mask), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(mask) failed with %d", (int)r);
      return B2_E_CUDA;
    }
  }

  const int64_t total_tiles = (n + kTcRows - 1) / kTcRows;
  const int grid = (int)(total_tiles < ctx->sm_count ? total_tiles : ctx->sm_count);
  int chunk_tiles = ctx->drain_rows / kTcRows;
  if (chunk_tiles < 1) chunk_tiles = 1;

  if (!ctx->tc_attr_set) {
    B2_CUDA(cudaFuncSetAttribute(gram_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    B2_CUDA(cudaFuncSetAttribute(gram_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemBytes));
    ctx->tc_attr_set = true;
  }

  if (x_dtype == B2_F32)
    tc_shift_kernel<float><<<kShiftBlocks, 160, 0, ctx->stream>>>(static_cast<const float*>(X), y, n, d, ldx, ctx->shift);
  else
    tc_shift_kernel<__nv_bfloat16><<<kShiftBlocks, 160, 0, ctx->stream>>>(static_cast<const __nv_bfloat16*>(X), y, n, d, ldx,
                                                                ctx->shift);
  B2_CUDA(cudaGetLastError());

  const int pair = ctx->k_pairs % kKernelEventPairs;
  B2_CUDA(cudaEventRecord(ctx->ev_k[pair][0], ctx->stream));
  if (x_dtype == B2_F32)
    gram_tc_kernel<float><<<grid, kThreads, kSmemBytes, ctx->stream>>>(
        tmX, tmY, tmM, y_map_2d, mask != nullptr ? 1 : 0, keep, n, d, ctx->shift, chunk_tiles, ctx->tc_part,
        ctx->tc_side, nullptr);
  else
    gram_tc_kernel<__nv_bfloat16><<<grid, kThreads, kSmemBytes, ctx->stream>>>(
        tmX, tmY, tmM, y_map_2d, mask != nullptr ? 1 : 0, keep, n, d, ctx->shift, chunk_tiles, ctx->tc_part,
        ctx->tc_side, nullptr);
  B2_CUDA(cudaGetLastError());
  B2_CUDA(cudaEventRecord(ctx->ev_k[pair][1], ctx->stream));
  ctx->k_pairs += 1;

  const int red_elems = kTcAccElems + 2 * kMaxD + 3;
  tc_reduce_kernel<<<(red_elems + 255) / 256, 256, 0, ctx->stream>>>(ctx->tc_part, ctx->tc_side, grid, ctx->tc_red);
  B2_CUDA(cudaGetLastError());
  const int dp = d + 2;
  tc_fold_kernel<<<(dp * dp + 255) / 256, 256, 0, ctx->stream>>>(ctx->tc_red, ctx->shift, n, d, ctx->S);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 4;
  ctx->k_launches += 4;
  return B2_OK;
}

}  // namespace b2
