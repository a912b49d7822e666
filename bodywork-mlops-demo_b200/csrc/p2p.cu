// p2p.cu -- one-shot all-reduce of the (D+2)^2 fp64 statistic over NVLink peer memory (no NCCL launch).
//
// The exchange step of the row-sharded fit: after the Gram kernel every rank
// holds its partial S.  Each rank owns an exchange buffer that all peers have mapped through CUDA IPC:
//
//   scatter kernel : every rank stores its S into slot[parity][rank] of EVERY rank's buffer (plain st.global on
//                    peer pointers: NVLink 5 / NVSwitch), then -- after a system-scope fence and a last-block
//                    ticket -- writes the epoch number into flag[rank] of every buffer;
//   gather kernel  : spins (bounded) until all n flags in its own buffer carry this epoch, then sums the n slots
//                    in rank order (bit-identical S on every rank, deterministic) into S.
//
// Slots are double buffered by epoch parity: a rank can be at most one exchange ahead of a peer (its next gather
// waits for that peer's next flag), so epoch k+1 data never overwrites slots a slow peer is still summing.
// 135 KB per rank: latency bound -- the point is to remove the collective launch + protocol latency from a
// ~1 ms step, and to keep compute (fold) and exchange in adjacent tiny kernels on one stream.
#include "b2_internal.cuh"

namespace b2 {
namespace {

struct PeerPtrs { double* p[kMaxRanks]; };

__host__ __device__ __forceinline__ unsigned int* flags_of(double* buf) {
  return reinterpret_cast<unsigned int*>(buf + kXchgDataDoubles);
}

__global__ void p2p_scatter_kernel(const double* __restrict__ S, int n_elems, PeerPtrs peers, int n_ranks, int rank,
                                   unsigned int epoch) {
  const size_t slot = ((size_t)(epoch & 1u) * kMaxRanks + rank) * kXchgSlotDoubles;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_elems; idx += gridDim.x * blockDim.x) {
    const double v = S[idx];
#pragma unroll 1
    for (int r = 0; r < n_ranks; ++r) peers.p[r][slot + idx] = v;      // r == rank: own buffer
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* ticket = flags_of(peers.p[rank]) + 32;               // own memory
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x < n_ranks) {
    __threadfence_system();
    volatile unsigned int* f = flags_of(peers.p[threadIdx.x]) + rank;  // "rank has delivered epoch"
    *f = epoch;
    __threadfence_system();
  }
}

__global__ void p2p_gather_kernel(double* __restrict__ S, int n_elems, double* own, int n_ranks, unsigned int epoch,
                                  int* __restrict__ status) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    volatile unsigned int* f = flags_of(own);
    long long t0 = clock64();
    int good = 1;
    for (int r = 0; r < n_ranks && good; ++r) {
      while ((int)(f[r] - epoch) < 0) {                                  // epochs are monotonic
        if (clock64() - t0 > 4000000000ll) { good = 0; break; }          // ~2 s: a peer died -- do not hang
      }
    }
    __threadfence_system();
    ok = good;
    if (!good && blockIdx.x == 0) *status = 1;
  }
  __syncthreads();
  if (!ok) return;
  const size_t base = (size_t)(epoch & 1u) * kMaxRanks * kXchgSlotDoubles;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_elems; idx += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < n_ranks; ++r) s += __ldcg(own + base + (size_t)r * kXchgSlotDoubles + idx);   // bypass L1
    S[idx] = s;
  }
}

}  // namespace

int launch_p2p_allreduce(b2_ctx* ctx) {
  const int n_elems = (ctx->d + 2) * (ctx->d + 2);
  PeerPtrs peers;
  for (int r = 0; r < kMaxRanks; ++r) peers.p[r] = ctx->xchg_peer[r];
  const unsigned int epoch = ++ctx->xchg_epoch;
  p2p_scatter_kernel<<<16, 256, 0, ctx->stream>>>(ctx->S, n_elems, peers, ctx->n_ranks, ctx->rank, epoch);
  B2_CUDA(cudaGetLastError());
  int* status = reinterpret_cast<int*>(flags_of(ctx->xchg) + 48);
  p2p_gather_kernel<<<16, 256, 0, ctx->stream>>>(ctx->S, n_elems, ctx->xchg, ctx->n_ranks, epoch, status);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return B2_OK;
}

}  // namespace b2
