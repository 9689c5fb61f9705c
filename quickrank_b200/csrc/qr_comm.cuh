// quickrank_b200 — multi-GPU plumbing (one process per GPU, NCCL loaded at run time).
#pragma once

#include "qr_internal.cuh"

namespace qr {

struct LeafSeg;

void comm_destroy(Comm *c);
int comm_allreduce_sum_f64(Comm *c, double *buf, size_t count, cudaStream_t st);
int comm_allreduce_max_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st);
// all-reduce the freshly built per-bin histograms (fixed-point sums + counts) and the squares
// partials of the first `ntasks` tasks in ctx->d_tasks
int comm_reduce_tasks(qr_ctx *ctx, uint32_t ntasks, bool root);
// all-reduce the per-leaf (sum lambda, sum weight) pairs and recompute the leaf outputs
int comm_leaf_values(qr_ctx *ctx, uint32_t nleaves);

}  // namespace qr
