// p2p.cu -- one-shot all-reduce of the (D+2)^2 fp64 statistic over NVLink peer memory (no NCCL launch).
//
// The exchange step of the row-sharded fit: after the Gram kernel every rank
// holds its partial S.  Protocol and buffer layout: b2_xchg.cuh.
//
//   p2p_scatter_kernel : stores S into slot[parity][rank] of EVERY rank's buffer, then -- after a system-scope fence
//                        and a last-block ticket -- publishes the exchange number in flag[rank] of every buffer;
//   p2p_gather_kernel  : waits (bounded) until all n flags in its own buffer carry this exchange, then sums the n
//                        slots in rank order (bit-identical S on every rank, deterministic) into S.  On a timeout it
//                        leaves S alone and raises the status word, which the host reads with the next result it
//                        fetches (b2_solve / b2_gram_export return B2_E_COMM instead of a fit on a partial statistic).
//
// These two launches serve the stand-alone b2_gram_allreduce call.  The fused fit (b2_fit, gram_tc.cu + solve.cu)
// issues the same stores from the finalize kernel's fold and the same wait + sum from the solve kernel's prologue.
#include "b2_xchg.cuh"

namespace b2 {
namespace {

__global__ void p2p_scatter_kernel(const double* __restrict__ S, int n_elems, PeerPtrs peers, int n_ranks, int rank,
                                   unsigned int epoch) {
  const size_t slot = xchg_slot_offset(epoch, rank);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_elems; idx += gridDim.x * blockDim.x)
    xchg_store_all(peers, n_ranks, slot, idx, S[idx]);
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    unsigned int* ticket = xchg_flags(peers.p[rank]) + kXchgTicketWord;   // own memory
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (last) xchg_publish(peers, n_ranks, rank, epoch);
}

__global__ void p2p_gather_kernel(double* __restrict__ S, int n_elems, double* own, int n_ranks, unsigned int epoch,
                                  unsigned long long timeout_ns) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const bool good = xchg_wait(own, n_ranks, epoch, timeout_ns);
    ok = good ? 1 : 0;
    if (!good && blockIdx.x == 0) xchg_flags(own)[kXchgStatusWord] = epoch;
  }
  __syncthreads();
  if (!ok) return;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_elems; idx += gridDim.x * blockDim.x)
    S[idx] = xchg_sum(own, n_ranks, epoch, idx);
}

}  // namespace

constexpr int kP2pCtas = 8;   // 135 KB per rank: latency bound; few CTAs, so a waiting gather leaves the SMs to the peers

int launch_p2p_allreduce(b2_ctx* ctx) {
  const int n_elems = (ctx->d + 2) * (ctx->d + 2);
  PeerPtrs peers;
  for (int r = 0; r < kMaxRanks; ++r) peers.p[r] = ctx->xchg_peer[r];
  const unsigned int epoch = ++ctx->xchg_epoch;
  p2p_scatter_kernel<<<kP2pCtas, 256, 0, ctx->stream>>>(ctx->S, n_elems, peers, ctx->n_ranks, ctx->rank, epoch);
  B2_CUDA(cudaGetLastError());
  p2p_gather_kernel<<<kP2pCtas, 256, 0, ctx->stream>>>(ctx->S, n_elems, ctx->xchg, ctx->n_ranks, epoch, ctx->xchg_timeout_ns);
  B2_CUDA(cudaGetLastError());
  ctx->launches += 2;
  ctx->xchg_pending = true;
  return B2_OK;
}

}  // namespace b2
