// b2_api.cu -- the C-ABI of libb2gram.so (declared in include/b2gram.h): context, caller-buffer
// helpers, dispatch between the tcgen05 and CUDA-core Gram kernels, host-streamed accumulation
// (pinned ring -> HBM staging, copy/compute overlap), NCCL all-reduce of the statistic, timing.
#include <dlfcn.h>
#include <stdarg.h>
#include <stdlib.h>

#include <new>

#include "b2_internal.cuh"

namespace b2 {

static thread_local char g_err[768] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

// ---- NCCL through dlopen: no link-time dependency, the library loads on CPU-only boxes ------------
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = reinterpret_cast<int (*)(NcclUid*)>(dlsym(api.lib, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclUid, int)>(dlsym(api.lib, "ncclCommInitRank"));
      api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
      api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
          dlsym(api.lib, "ncclAllReduce"));
      api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce) return nullptr;
  return &api;
}

#define B2_NCCL(api, call)                                                                     \
  do {                                                                                         \
    int r_ = (call);                                                                           \
    if (r_ != 0) {                                                                             \
      set_error("%s failed: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r_) : "?"); \
      return B2_E_NCCL;                                                                        \
    }                                                                                          \
  } while (0)

int use_device(b2_ctx* ctx) {
  if (ctx == nullptr) {
    set_error("null context");
    return B2_E_ARG;
  }
  B2_CUDA(cudaSetDevice(ctx->device));
  return B2_OK;
}

int check_shape(b2_ctx* ctx, int x_dtype, int64_t n, int d, int64_t ldx, int mem_kind) {
  if (x_dtype != B2_F32 && x_dtype != B2_BF16) { set_error("x_dtype must be B2_F32 or B2_BF16"); return B2_E_ARG; }
  if (d < 1 || d > kMaxD) { set_error("d=%d out of range [1,%d]", d, kMaxD); return B2_E_ARG; }
  if (n < 0) { set_error("n_rows < 0"); return B2_E_ARG; }
  if (ldx < d) { set_error("ldx=%lld < d=%d", (long long)ldx, d); return B2_E_ARG; }
  if (mem_kind != B2_MEM_DEVICE && mem_kind != B2_MEM_HOST) { set_error("bad mem_kind %d", mem_kind); return B2_E_ARG; }
  (void)ctx;
  return B2_OK;
}

// One device-resident block through the selected Gram kernel.
int gram_block(b2_ctx* ctx, const void* X, int x_dtype, const float* y, int64_t n, int d, int64_t ldx,
               const uint8_t* mask, int keep) {
  if (n == 0) return B2_OK;
  const bool tc_ok = gram_tc_supported(X, x_dtype, y, n, d, ldx) &&
                     (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 15) == 0);
  const bool nw_ok = gram_narrow_supported(X, x_dtype, y, n, d, ldx, mask);
  int mode = ctx->kernel_mode;
  if (mode == B2_KERNEL_TCGEN05 && !tc_ok) {
    set_error("tcgen05 path needs d%%4==0 (fp32) / d%%8==0 (bf16), 16-byte aligned X/y/mask/row pitch, n>=32");
    return B2_E_UNSUPPORTED;
  }
  if (mode == B2_KERNEL_NARROW && !nw_ok) {
    set_error("narrow path needs d<=16, contiguous rows (ldx==d) and 16-byte aligned X/y/mask");
    return B2_E_UNSUPPORTED;
  }
  if (mode == B2_KERNEL_AUTO) {
    // narrow rows stream through the CUDA-core pipeline (HBM-bound); wide rows go to the tensor core;
    // tiny tranches (the reference's 1 440-row day) and odd layouts stay on the exact fp64 kernel
    if (nw_ok && n >= 4096) mode = B2_KERNEL_NARROW;
    else mode = (tc_ok && n >= 2048) ? B2_KERNEL_TCGEN05 : B2_KERNEL_SIMT;
  }
  if (mode == B2_KERNEL_NARROW) return launch_gram_narrow(ctx, X, x_dtype, y, n, d, ldx, mask, keep);
  if (mode == B2_KERNEL_TCGEN05) return launch_gram_tc(ctx, X, x_dtype, y, n, d, ldx, mask, keep);
  return launch_gram_simt(ctx, X, x_dtype, y, n, d, ldx, mask, keep);
}

int ensure_staging(b2_ctx* ctx) {
  if (ctx->stage_x[0] != nullptr) return B2_OK;
  ctx->stage_rows = 1 << 18;                                  // 262 144 rows per block (134 MB at 128 x fp32)
  ctx->stage_bytes_x = (size_t)ctx->stage_rows * kMaxD * 4;
  for (int b = 0; b < 2; ++b) {
    B2_CUDA(cudaMalloc(&ctx->stage_x[b], ctx->stage_bytes_x));
    B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stage_y[b]), (size_t)ctx->stage_rows * 4));
    B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->stage_m[b]), (size_t)ctx->stage_rows));
    B2_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&ctx->ev_consumed[b], cudaEventDisableTiming));
  }
  return B2_OK;
}

// copy rows [r0, r0+rows) of a host matrix into a compact (ldx == d) staging block
int stage_rows_h2d(b2_ctx* ctx, int buf, const void* X, int es, const float* y, const uint8_t* mask, int64_t r0,
                   int64_t rows, int d, int64_t ldx) {
  const char* src = static_cast<const char*>(X) + (size_t)r0 * ldx * es;
  if (ldx == d) {
    B2_CUDA(cudaMemcpyAsync(ctx->stage_x[buf], src, (size_t)rows * d * es, cudaMemcpyHostToDevice, ctx->copy_stream));
  } else {
    B2_CUDA(cudaMemcpy2DAsync(ctx->stage_x[buf], (size_t)d * es, src, (size_t)ldx * es, (size_t)d * es, rows,
                              cudaMemcpyHostToDevice, ctx->copy_stream));
  }
  if (y != nullptr)
    B2_CUDA(cudaMemcpyAsync(ctx->stage_y[buf], y + r0, (size_t)rows * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
  if (mask != nullptr)
    B2_CUDA(cudaMemcpyAsync(ctx->stage_m[buf], mask + r0, (size_t)rows, cudaMemcpyHostToDevice, ctx->copy_stream));
  return B2_OK;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" {

int b2_abi_version(void) { return B2_ABI_VERSION; }
const char* b2_last_error(void) { return g_err; }

int b2_device_count(int* n_out) {
  if (n_out == nullptr) { set_error("n_out is null"); return B2_E_ARG; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *n_out = 0;
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return B2_E_CUDA;
  }
  *n_out = n;
  return B2_OK;
}

// streams, events and device scratch of a fresh context (b2_ctx_create destroys the context if this fails)
static int ctx_allocate(b2_ctx* ctx) {
  B2_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  B2_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  B2_CUDA(cudaEventCreate(&ctx->ev_t0));
  B2_CUDA(cudaEventCreate(&ctx->ev_t1));
  for (int i = 0; i < kKernelEventPairs; ++i) {
    B2_CUDA(cudaEventCreate(&ctx->ev_k[i][0]));
    B2_CUDA(cudaEventCreate(&ctx->ev_k[i][1]));
  }
  ctx->simt_ctas = ctx->sm_count;
  ctx->score_ctas = ctx->sm_count * 8;
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->S), sizeof(double) * kMaxS * kMaxS));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tc_part), sizeof(double) * (size_t)ctx->sm_count * kTcAccElems));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tc_side), sizeof(double) * (size_t)ctx->sm_count * kTcSideDoubles));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tc_red), sizeof(double) * (kTcAccElems + 16 + kMaxD + 8)));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->shift), sizeof(float) * 64 * (kMaxD + 1)));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->simt_part), sizeof(double) * (size_t)ctx->simt_ctas * kMaxS * kMaxS));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->score_part), sizeof(double) * ((size_t)ctx->score_ctas + 2) * 10));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->coef_dev), sizeof(double) * (kMaxD + 1)));
  B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->solve_out), sizeof(double) * (2 * kMaxD + 8)));
  B2_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->solve_host), sizeof(double) * (2 * kMaxD + 8), cudaHostAllocDefault));
  B2_CUDA(cudaMemset(ctx->S, 0, sizeof(double) * kMaxS * kMaxS));
  B2_CUDA(cudaMemset(ctx->tc_side, 0, sizeof(double) * (size_t)ctx->sm_count * kTcSideDoubles));
  B2_CUDA(cudaMemset(ctx->tc_red, 0, sizeof(double) * (kTcAccElems + 16 + kMaxD + 8)));
  return B2_OK;
}

int b2_ctx_create(int device, b2_ctx** out) {
  if (out == nullptr) { set_error("out is null"); return B2_E_ARG; }
  *out = nullptr;
  int n = 0;
  if (b2_device_count(&n) != B2_OK || n == 0) {
    set_error("no usable CUDA device (libb2gram has no CPU fallback)");
    return B2_E_CUDA;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return B2_E_ARG; }
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libb2gram is built for sm_100a only", device, prop.major, prop.minor);
    return B2_E_UNSUPPORTED;
  }
  b2_ctx* ctx = new (std::nothrow) b2_ctx();
  if (ctx == nullptr) { set_error("out of host memory"); return B2_E_STATE; }
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->hbm_bytes = prop.totalGlobalMem;
  snprintf(ctx->name, sizeof(ctx->name), "%s", prop.name);
  if (int r = ctx_allocate(ctx)) {   // streams, events, scratch: release whatever was created before the failure
    b2_ctx_destroy(ctx);
    return r;
  }
  *out = ctx;
  return B2_OK;
}

int b2_ctx_destroy(b2_ctx* ctx) {
  if (ctx == nullptr) return B2_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm != nullptr) b2_comm_destroy(ctx);
  b2_comm_p2p_detach(ctx);
  if (ctx->xchg != nullptr) cudaFree(ctx->xchg);
  void* bufs[] = {ctx->S, ctx->tc_part, ctx->tc_side, ctx->tc_red, ctx->shift, ctx->simt_part, ctx->score_part,
                  ctx->coef_dev, ctx->solve_out, ctx->stage_x[0], ctx->stage_x[1], ctx->stage_y[0], ctx->stage_y[1],
                  ctx->stage_m[0], ctx->stage_m[1]};
  for (void* p : bufs) if (p != nullptr) cudaFree(p);
  if (ctx->solve_host != nullptr) cudaFreeHost(ctx->solve_host);
  for (int b = 0; b < 2; ++b) {
    if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
    if (ctx->ev_consumed[b]) cudaEventDestroy(ctx->ev_consumed[b]);
  }
  for (int i = 0; i < kKernelEventPairs; ++i)
    for (int e = 0; e < 2; ++e)
      if (ctx->ev_k[i][e] != nullptr) cudaEventDestroy(ctx->ev_k[i][e]);
  if (ctx->ev_t0 != nullptr) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1 != nullptr) cudaEventDestroy(ctx->ev_t1);
  if (ctx->stream != nullptr) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream != nullptr) cudaStreamDestroy(ctx->copy_stream);
  cudaGetLastError();   // a half-built context may have left a sticky-free error code behind
  delete ctx;
  return B2_OK;
}

int b2_ctx_sync(b2_ctx* ctx) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaStreamSynchronize(ctx->copy_stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

int b2_ctx_info(b2_ctx* ctx, char* name, int name_cap, int* sm_count, size_t* hbm_bytes) {
  if (ctx == nullptr) { set_error("null context"); return B2_E_ARG; }
  if (name != nullptr && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", ctx->name);
  if (sm_count != nullptr) *sm_count = ctx->sm_count;
  if (hbm_bytes != nullptr) *hbm_bytes = ctx->hbm_bytes;
  return B2_OK;
}

int b2_ctx_set_kernel(b2_ctx* ctx, int kernel) {
  if (ctx == nullptr || kernel < B2_KERNEL_AUTO || kernel > B2_KERNEL_NARROW) { set_error("bad kernel id"); return B2_E_ARG; }
  ctx->kernel_mode = kernel;
  return B2_OK;
}

int b2_ctx_set_drain_rows(b2_ctx* ctx, int rows) {
  if (ctx == nullptr || rows < kTcRows || rows % kTcRows != 0) { set_error("drain_rows must be a positive multiple of %d", kTcRows); return B2_E_ARG; }
  ctx->drain_rows = rows;
  return B2_OK;
}

int b2_ctx_set_precision(b2_ctx* ctx, int precision) {
  if (ctx == nullptr || (precision != B2_PRECISION_SPLIT && precision != B2_PRECISION_BF16)) {
    set_error("precision must be B2_PRECISION_SPLIT or B2_PRECISION_BF16");
    return B2_E_ARG;
  }
  ctx->precision = precision;
  return B2_OK;
}

// ---- buffers ------------------------------------------------------------------------------------
int b2_dev_alloc(b2_ctx* ctx, size_t bytes, void** out) {
  if (int r = use_device(ctx)) return r;
  if (out == nullptr) { set_error("out is null"); return B2_E_ARG; }
  B2_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return B2_OK;
}
int b2_dev_free(b2_ctx* ctx, void* p) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaFree(p));
  return B2_OK;
}
int b2_host_alloc(b2_ctx* ctx, size_t bytes, void** out) {
  if (int r = use_device(ctx)) return r;
  if (out == nullptr) { set_error("out is null"); return B2_E_ARG; }
  B2_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return B2_OK;
}
int b2_host_free(b2_ctx* ctx, void* p) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaFreeHost(p));
  return B2_OK;
}
int b2_copy_h2d(b2_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}
int b2_copy_d2h(b2_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}
int b2_dev_memset(b2_ctx* ctx, void* dst, int value, size_t bytes) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return B2_OK;
}

// ---- Gram -----------------------------------------------------------------------------------------
int b2_gram_reset(b2_ctx* ctx, int d) {
  if (int r = use_device(ctx)) return r;
  if (d < 1 || d > kMaxD) { set_error("d=%d out of range [1,%d]", d, kMaxD); return B2_E_ARG; }
  ctx->d = d;
  B2_CUDA(cudaMemsetAsync(ctx->S, 0, sizeof(double) * kMaxS * kMaxS, ctx->stream));
  return B2_OK;
}

int b2_gram_accumulate(b2_ctx* ctx, const void* X, int x_dtype, const float* y, int64_t n_rows, int d, int64_t ldx,
                       int mem_kind, const uint8_t* row_mask, int mask_keep) {
  if (int r = use_device(ctx)) return r;
  if (int r = check_shape(ctx, x_dtype, n_rows, d, ldx, mem_kind)) return r;
  if (ctx->d == 0) { set_error("b2_gram_reset has not been called"); return B2_E_STATE; }
  if (d != ctx->d) { set_error("d=%d differs from the statistic's d=%d", d, ctx->d); return B2_E_ARG; }
  if (n_rows > 0 && (X == nullptr || y == nullptr)) { set_error("X / y is null"); return B2_E_ARG; }
  ctx->k_launches = 0;
  if (mem_kind == B2_MEM_DEVICE) return gram_block(ctx, X, x_dtype, y, n_rows, d, ldx, row_mask, mask_keep);

  // host rows: stream blocks through a 2-deep HBM staging ring, copies overlapping the kernels
  if (int r = ensure_staging(ctx)) return r;
  const int es = x_dtype == B2_F32 ? 4 : 2;
  int64_t blk = 0;
  int rc = B2_OK;
  auto step = [&](int64_t r0, int buf, int64_t rows) -> int {
    if (blk >= 2) B2_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0));
    if (int r = stage_rows_h2d(ctx, buf, X, es, y, row_mask, r0, rows, d, ldx)) return r;
    B2_CUDA(cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream));
    B2_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0));
    if (int r = gram_block(ctx, ctx->stage_x[buf], x_dtype, ctx->stage_y[buf], rows, d, d,
                           row_mask ? ctx->stage_m[buf] : nullptr, mask_keep))
      return r;
    B2_CUDA(cudaEventRecord(ctx->ev_consumed[buf], ctx->stream));
    return B2_OK;
  };
  for (int64_t r0 = 0; r0 < n_rows && rc == B2_OK; r0 += ctx->stage_rows, ++blk) {
    const int64_t rows = (n_rows - r0 < ctx->stage_rows) ? n_rows - r0 : ctx->stage_rows;
    rc = step(r0, (int)(blk & 1), rows);
  }
  // the caller may reuse its host buffers on return -- also when a block failed
  const cudaError_t drained = cudaStreamSynchronize(ctx->copy_stream);
  if (rc != B2_OK) return rc;
  B2_CUDA(drained);
  return B2_OK;
}

int b2_gram_allreduce(b2_ctx* ctx) {
  if (int r = use_device(ctx)) return r;
  if (ctx->d == 0) { set_error("b2_gram_reset has not been called"); return B2_E_STATE; }
  if (ctx->n_ranks > 1 && ctx->p2p_ready) return launch_p2p_allreduce(ctx);   // peer-memory one-shot exchange
  if (ctx->n_ranks == 1 || ctx->comm == nullptr) return B2_OK;
  NcclApi* api = nccl();
  if (api == nullptr) { set_error("libnccl.so.2 could not be loaded"); return B2_E_NCCL; }
  const size_t count = (size_t)(ctx->d + 2) * (ctx->d + 2);
  B2_NCCL(api, api->AllReduce(ctx->S, ctx->S, count, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
  return B2_OK;
}

int b2_gram_export(b2_ctx* ctx, double* S_out, int64_t* n_rows_out) {
  if (int r = use_device(ctx)) return r;
  if (ctx->d == 0) { set_error("b2_gram_reset has not been called"); return B2_E_STATE; }
  const int dp = ctx->d + 2;
  if (S_out == nullptr) { set_error("S_out is null"); return B2_E_ARG; }
  B2_CUDA(cudaMemcpyAsync(S_out, ctx->S, sizeof(double) * dp * dp, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_rows_out != nullptr) *n_rows_out = (int64_t)(S_out[ctx->d * dp + ctx->d] + 0.5);
  return B2_OK;
}

int b2_gram_import(b2_ctx* ctx, const double* S_in, int d) {
  if (int r = use_device(ctx)) return r;
  if (d < 1 || d > kMaxD || S_in == nullptr) { set_error("bad arguments to b2_gram_import"); return B2_E_ARG; }
  ctx->d = d;
  B2_CUDA(cudaMemsetAsync(ctx->S, 0, sizeof(double) * kMaxS * kMaxS, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(ctx->S, S_in, sizeof(double) * (d + 2) * (d + 2), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

// ---- solve ------------------------------------------------------------------------------------------
static int fetch_solution(b2_ctx* ctx, double* coef, double* intercept, double* singular, int* rank, double* info) {
  double* host = ctx->solve_host;
  B2_CUDA(cudaMemcpyAsync(host, ctx->solve_out, sizeof(double) * (2 * kMaxD + 8), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  if (coef != nullptr) memcpy(coef, host, sizeof(double) * ctx->d);
  if (intercept != nullptr) *intercept = host[kMaxD];
  *info = host[kMaxD + 1];
  if (rank != nullptr) *rank = (int)host[kMaxD + 2];
  if (singular != nullptr) memcpy(singular, host + kMaxD + 3, sizeof(double) * ctx->d);
  return B2_OK;
}

int b2_solve(b2_ctx* ctx, double alpha, int fit_intercept, double* coef, double* intercept) {
  if (int r = use_device(ctx)) return r;
  if (ctx->d == 0) { set_error("b2_gram_reset has not been called"); return B2_E_STATE; }
  if (!(alpha >= 0.0)) { set_error("alpha must be >= 0"); return B2_E_ARG; }
  if (int r = launch_solve_cholesky(ctx, alpha, fit_intercept)) return r;
  double info = 0.0;
#ifdef B2_DEV_KNOBS
  double phase[kMaxD];
  if (int r = fetch_solution(ctx, coef, intercept, getenv("B2_SOLVE_TIMING") ? phase : nullptr, nullptr, &info)) return r;
  if (getenv("B2_SOLVE_TIMING"))
    fprintf(stderr, "[b2_solve] cycles: build %.0f diag %.0f panel %.0f update %.0f backward %.0f\n", phase[0], phase[1],
            phase[2], phase[3], phase[4]);
#else
  if (int r = fetch_solution(ctx, coef, intercept, nullptr, nullptr, &info)) return r;
#endif
  if (info != 0.0) {
    set_error("Cholesky pivot %d is not positive: the centred Gram matrix is rank deficient "
              "(use alpha > 0 or b2_solve_spectral)", (int)info);
    return B2_E_SINGULAR;
  }
  return B2_OK;
}

int b2_solve_spectral(b2_ctx* ctx, double cond, int fit_intercept, double* coef, double* intercept, double* singular,
                      int* rank) {
  if (int r = use_device(ctx)) return r;
  if (ctx->d == 0) { set_error("b2_gram_reset has not been called"); return B2_E_STATE; }
  if (int r = launch_solve_spectral(ctx, cond, fit_intercept)) return r;
  double info = 0.0;
  return fetch_solution(ctx, coef, intercept, singular, rank, &info);
}

// ---- scoring ---------------------------------------------------------------------------------------
int b2_score(b2_ctx* ctx, const void* X, int x_dtype, int64_t n_rows, int d, int64_t ldx, int mem_kind,
             const double* coef, double intercept, const float* y, const uint8_t* row_mask, int mask_keep,
             float* yhat, double* stats_out) {
  if (int r = use_device(ctx)) return r;
  if (int r = check_shape(ctx, x_dtype, n_rows, d, ldx, mem_kind)) return r;
  if (coef == nullptr || (n_rows > 0 && X == nullptr)) { set_error("coef / X is null"); return B2_E_ARG; }
  double cbuf[kMaxD + 1];
  memset(cbuf, 0, sizeof(cbuf));
  memcpy(cbuf, coef, sizeof(double) * d);
  cbuf[kMaxD] = intercept;
  B2_CUDA(cudaMemcpyAsync(ctx->coef_dev, cbuf, sizeof(cbuf), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));  // cbuf is on this stack frame
  double* acc = ctx->score_part + (size_t)ctx->score_ctas * 10;
  if (n_rows == 0) {
    B2_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 10, ctx->stream));
  } else if (mem_kind == B2_MEM_DEVICE) {
    if (int r = launch_score(ctx, X, x_dtype, n_rows, d, ldx, y, row_mask, mask_keep, yhat, true)) return r;
  } else {
    if (int r = ensure_staging(ctx)) return r;
    const int es = x_dtype == B2_F32 ? 4 : 2;
    // predictions of a staged block land in a device block of their own and are copied back behind the kernel
    float* yhat_dev[2] = {nullptr, nullptr};
    if (yhat != nullptr) {
      for (int b = 0; b < 2; ++b) {
        if (cudaMalloc(reinterpret_cast<void**>(&yhat_dev[b]), (size_t)ctx->stage_rows * 4) != cudaSuccess) {
          cudaGetLastError();
          if (yhat_dev[0] != nullptr) cudaFree(yhat_dev[0]);
          set_error("out of device memory for the prediction staging blocks");
          return B2_E_CUDA;
        }
      }
    }
    int64_t blk = 0;
    int rc = B2_OK;
    for (int64_t r0 = 0; r0 < n_rows && rc == B2_OK; r0 += ctx->stage_rows, ++blk) {
      const int buf = (int)(blk & 1);
      const int64_t rows = (n_rows - r0 < ctx->stage_rows) ? n_rows - r0 : ctx->stage_rows;
      if (blk >= 2) cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[buf], 0);
      rc = stage_rows_h2d(ctx, buf, X, es, y, row_mask, r0, rows, d, ldx);
      if (rc != B2_OK) break;
      cudaEventRecord(ctx->ev_copied[buf], ctx->copy_stream);
      cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[buf], 0);
      rc = launch_score(ctx, ctx->stage_x[buf], x_dtype, rows, d, d, y ? ctx->stage_y[buf] : nullptr,
                        row_mask ? ctx->stage_m[buf] : nullptr, mask_keep, yhat ? yhat_dev[buf] : nullptr, blk == 0);
      if (rc != B2_OK) break;
      if (yhat != nullptr)
        cudaMemcpyAsync(yhat + r0, yhat_dev[buf], (size_t)rows * 4, cudaMemcpyDeviceToHost, ctx->stream);
      cudaEventRecord(ctx->ev_consumed[buf], ctx->stream);
    }
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (int b = 0; b < 2; ++b) if (yhat_dev[b]) cudaFree(yhat_dev[b]);
    if (rc != B2_OK) return rc;
    B2_CUDA(cudaGetLastError());
  }
  if (stats_out != nullptr && y != nullptr) {
    B2_CUDA(cudaMemcpyAsync(stats_out, acc, sizeof(double) * 10, cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return B2_OK;
}

int b2_score_allreduce(b2_ctx* ctx, double* stats) {
  if (int r = use_device(ctx)) return r;
  if (stats == nullptr) { set_error("stats is null"); return B2_E_ARG; }
  if (ctx->n_ranks == 1 || ctx->comm == nullptr) return B2_OK;
  NcclApi* api = nccl();
  if (api == nullptr) { set_error("libnccl.so.2 could not be loaded"); return B2_E_NCCL; }
  double* acc = ctx->score_part + (size_t)ctx->score_ctas * 10;   // 10 sums
  double* mx = acc + 10;                                          // 2 maxima
  double host[10], hmax[2];
  memcpy(host, stats, sizeof(host));
  hmax[0] = host[4]; hmax[1] = host[9];
  host[4] = 0.0; host[9] = 0.0;
  B2_CUDA(cudaMemcpyAsync(acc, host, sizeof(host), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(mx, hmax, sizeof(hmax), cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));   // host / hmax live on this stack frame
  B2_NCCL(api, api->AllReduce(acc, acc, 10, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
  B2_NCCL(api, api->AllReduce(mx, mx, 2, kNcclFloat64, kNcclMax, ctx->comm, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(host, acc, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(hmax, mx, sizeof(hmax), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  host[4] = hmax[0]; host[9] = hmax[1];
  memcpy(stats, host, sizeof(host));
  return B2_OK;
}

// ---- synthetic rows -------------------------------------------------------------------------------------
int b2_synth(b2_ctx* ctx, uint64_t seed, int64_t row_offset, int64_t n_rows, int d, int64_t ldx, int x_dtype,
             double alpha, double beta, double sigma, void* X_dev, float* y_dev) {
  if (int r = use_device(ctx)) return r;
  if (int r = check_shape(ctx, x_dtype, n_rows, d, ldx, B2_MEM_DEVICE)) return r;
  if (n_rows > 0 && (X_dev == nullptr || y_dev == nullptr)) { set_error("X / y is null"); return B2_E_ARG; }
  return launch_synth(ctx, seed, row_offset, n_rows, d, ldx, x_dtype, alpha, beta, sigma, X_dev, y_dev);
}

// ---- multi-GPU --------------------------------------------------------------------------------------------
int b2_comm_unique_id(char* id_out) {
  if (id_out == nullptr) { set_error("id_out is null"); return B2_E_ARG; }
  NcclApi* api = nccl();
  if (api == nullptr) { set_error("libnccl.so.2 could not be loaded"); return B2_E_NCCL; }
  NcclUid uid;
  B2_NCCL(api, api->GetUniqueId(&uid));
  memcpy(id_out, uid.internal, 128);
  return B2_OK;
}

int b2_comm_init(b2_ctx* ctx, int n_ranks, int rank, const char* id) {
  if (int r = use_device(ctx)) return r;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || id == nullptr) { set_error("bad communicator arguments"); return B2_E_ARG; }
  if (ctx->comm != nullptr) { set_error("communicator already initialised"); return B2_E_STATE; }
  NcclApi* api = nccl();
  if (api == nullptr) { set_error("libnccl.so.2 could not be loaded"); return B2_E_NCCL; }
  NcclUid uid;
  memcpy(uid.internal, id, 128);
  B2_NCCL(api, api->CommInitRank(&ctx->comm, n_ranks, uid, rank));
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  return B2_OK;
}

int b2_comm_destroy(b2_ctx* ctx) {
  if (ctx == nullptr || ctx->comm == nullptr) return B2_OK;
  NcclApi* api = nccl();
  if (api != nullptr && api->CommDestroy != nullptr) api->CommDestroy(ctx->comm);
  ctx->comm = nullptr;
  ctx->n_ranks = 1;
  ctx->rank = 0;
  return B2_OK;
}

int b2_comm_p2p_export(b2_ctx* ctx, char* handle_out) {
  if (int r = use_device(ctx)) return r;
  if (handle_out == nullptr) { set_error("handle_out is null"); return B2_E_ARG; }
  if (ctx->xchg == nullptr) {
    B2_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->xchg), kXchgBytes));
    B2_CUDA(cudaMemset(ctx->xchg, 0, kXchgBytes));
    B2_CUDA(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, ctx->xchg));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  return B2_OK;
}

int b2_comm_p2p_attach(b2_ctx* ctx, int n_ranks, int rank, const char* handles) {
  if (int r = use_device(ctx)) return r;
  if (n_ranks < 2 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || handles == nullptr || ctx->xchg == nullptr) {
    set_error("b2_comm_p2p_attach: bad arguments (2..%d ranks; call b2_comm_p2p_export first)", kMaxRanks);
    return B2_E_ARG;
  }
  if (ctx->p2p_ready) { set_error("peer exchange already attached"); return B2_E_STATE; }
  for (int r = 0; r < n_ranks; ++r) {
    if (r == rank) { ctx->xchg_peer[r] = ctx->xchg; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    B2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->xchg_peer[r] = static_cast<double*>(p);
  }
  ctx->n_ranks = n_ranks;
  ctx->rank = rank;
  ctx->xchg_epoch = 0;
  ctx->p2p_ready = true;
  return B2_OK;
}

int b2_comm_p2p_detach(b2_ctx* ctx) {
  if (ctx == nullptr) { set_error("null context"); return B2_E_ARG; }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->p2p_ready)
    for (int r = 0; r < ctx->n_ranks; ++r)
      if (r != ctx->rank && ctx->xchg_peer[r] != nullptr) cudaIpcCloseMemHandle(ctx->xchg_peer[r]);
  for (int r = 0; r < kMaxRanks; ++r) ctx->xchg_peer[r] = nullptr;
  ctx->p2p_ready = false;
  return B2_OK;
}

int b2_comm_barrier(b2_ctx* ctx) {
  if (int r = use_device(ctx)) return r;
  if (ctx->n_ranks == 1 || ctx->comm == nullptr) return b2_ctx_sync(ctx);
  NcclApi* api = nccl();
  if (api == nullptr) { set_error("libnccl.so.2 could not be loaded"); return B2_E_NCCL; }
  double* slot = ctx->tc_red + kTcAccElems + 8;  // spare scratch
  B2_NCCL(api, api->AllReduce(slot, slot, 1, kNcclFloat64, kNcclSum, ctx->comm, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return B2_OK;
}

// ---- timing ---------------------------------------------------------------------------------------------------
int b2_timer_start(b2_ctx* ctx) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
  return B2_OK;
}
int b2_timer_stop(b2_ctx* ctx, double* ms_out) {
  if (int r = use_device(ctx)) return r;
  B2_CUDA(cudaEventRecord(ctx->ev_t1, ctx->stream));
  B2_CUDA(cudaEventSynchronize(ctx->ev_t1));
  float ms = 0.f;
  B2_CUDA(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
  if (ms_out != nullptr) *ms_out = (double)ms;
  return B2_OK;
}
int b2_last_kernel_ms(b2_ctx* ctx, double* gram_ms_out, int* launches_out) {
  if (int r = use_device(ctx)) return r;
  double total = 0.0;
  const int n = ctx->k_pairs < kKernelEventPairs ? ctx->k_pairs : kKernelEventPairs;
  for (int i = 0; i < n; ++i) {
    B2_CUDA(cudaEventSynchronize(ctx->ev_k[i][1]));
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, ctx->ev_k[i][0], ctx->ev_k[i][1]));
    total += ms;
  }
  if (gram_ms_out != nullptr) *gram_ms_out = total;
  if (launches_out != nullptr) *launches_out = n;
  ctx->k_pairs = 0;
  return B2_OK;
}
int b2_launch_count(b2_ctx* ctx, int64_t* n_out) {
  if (ctx == nullptr || n_out == nullptr) { set_error("null argument"); return B2_E_ARG; }
  *n_out = ctx->launches;
  return B2_OK;
}

}  // extern "C"
