// quickrank_b200 — multi-GPU plumbing (one process per GPU, NCCL loaded at run time).
#pragma once

#include <vector>

#include "qr_internal.cuh"

namespace qr {

int comm_create(const unsigned char id[QR_COMM_ID_BYTES], int rank, int world, Comm **out);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_world(const Comm *c);

int comm_allreduce_sum_f64(Comm *c, double *buf, size_t count, cudaStream_t st);
int comm_allreduce_max_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st);
int comm_allreduce_sum_u32(Comm *c, uint32_t *buf, size_t count, cudaStream_t st);
int comm_allreduce_sum_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st);
// every rank contributes `bytes` bytes from host memory (same size on all ranks); `out` receives
// all contributions, rank-major
int comm_allgather_host(Comm *c, const void *send, size_t bytes, std::vector<unsigned char> *out,
                        cudaStream_t st);

// all-reduce the freshly built per-bin histograms (fixed-point sums + counts) and the squares
// sums of the first `ntasks` tasks in ctx->d_tasks
int comm_reduce_tasks(qr_ctx *ctx, uint32_t ntasks, bool root);
// export this rank's histogram pool over CUDA IPC, map the peers' and agree on the exchange path
// (peer-memory kernel when every rank can map every other, NCCL all-reduces otherwise)
int comm_setup_peers(qr_ctx *ctx);
// 0: single GPU, 1: NCCL all-reduces, 2: peer-memory kernel
int comm_transport(const Comm *c);
// fills the view finalize_kernel needs for the fused exchange of the current round (takes one barrier epoch)
void comm_peer_view(qr_ctx *ctx, bool with_counts, PeerView *pv);
// all-reduce the per-leaf (sum lambda, sum weight) pairs and recompute the leaf outputs
int comm_leaf_values(qr_ctx *ctx, uint32_t nleaves);

}  // namespace qr
