// quickrank_b200 — multi-GPU plumbing (one process per GPU, NCCL loaded at run time).
#pragma once

#include "qr_internal.cuh"

namespace qr {

struct LeafSeg;

void comm_destroy(Comm *c);
int comm_allreduce_sum_f64(Comm *c, double *buf, size_t count, cudaStream_t st);
int comm_allreduce_max_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st);
// all-reduce a node's per-bin histogram (fixed-point sums + counts) and its squares partials
int comm_reduce_hist(qr_ctx *ctx, unsigned long long *hsum, uint32_t *hcnt, uint32_t *n_partials);
int comm_local_lcount(qr_ctx *ctx, int node);
int comm_leaf_fit(qr_ctx *ctx, const LeafSeg *d_segs, uint32_t nleaves, bool root_only);

}  // namespace qr
