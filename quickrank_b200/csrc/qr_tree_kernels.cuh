// quickrank_b200 — device kernels of tree growth (sm_100a): histograms, split scan, partition,
// leaf fit.  Every kernel works on a BATCH of node expansions described by NodeTask records, so
// one launch serves all the frontier nodes a growth round expands (qr_train.cu explains why that
// reproduces the reference's one-node-at-a-time heap order exactly).
#pragma once

#include "qr_kernels.cuh"
#include "qr_task.cuh"
#include "qr_fast_kernels.cuh"
#include "qr_exact_kernels.cuh"

namespace qr {

// Clears the histogram slot each task builds into.  For the root refresh the per-bin counts never
// change from tree to tree ("count doesn't change, so no need to re-compute",
// rtnode_histogram.cc:149): they are copied from the table made at init instead of being recounted.
__global__ void prep_slots_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ TaskPack pack,
                                  unsigned long long *hsum, uint32_t *hcnt, uint32_t ncells,
                                  const uint32_t *__restrict__ root_cnt) {
  if (pack.n) tasks = pack.t;
  const NodeTask t = tasks[blockIdx.y];
  unsigned long long *s = hsum + (size_t) build_slot(t) * ncells;
  uint32_t *c = hcnt + (size_t) build_slot(t) * ncells;
  const bool copy = t.whole && root_cnt != nullptr;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) {
    s[i] = 0ull;
    c[i] = copy ? root_cnt[i] : 0u;
  }
}

// ------------------------------------------------------------------------------------------
// Stable partition of each task's document list by bin(f, doc) <= t  (rt.cc:325-334; equal to
// the reference's float test because thresholds ascend, SURVEY.md section 7.1 "Bins").
// count -> per-task exclusive prefix -> scatter.
// ------------------------------------------------------------------------------------------
template <typename BinT>
__global__ void __launch_bounds__(256)
part_count_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint4 *__restrict__ panels,
                  size_t N, const uint32_t *__restrict__ ids0, const uint32_t *__restrict__ ids1,
                  uint32_t *blockcnt) {
  __shared__ uint32_t s_task;
  __shared__ uint32_t w[8];
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, blockIdx.x, false);
  __syncthreads();
  const NodeTask t = tasks[s_task];
  const uint32_t *src = t.src == 1 ? ids1 : ids0;
  const uint32_t b0 = (blockIdx.x - t.part_blk0) * kPartItems, e = min(t.n, b0 + kPartItems);
  uint32_t c = 0;
  for (uint32_t i = b0 + threadIdx.x; i < e; i += 256) {
    const uint32_t d = t.src == 2 ? t.lo + i : src[t.lo + i];
    c += load_bin<BinT>(panels, N, t.f, d) <= t.t;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane_id() == 0) w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int k = 0; k < 8; ++k) s += w[k];
    blockcnt[blockIdx.x] = s;
  }
}

// one block per task: blockcnt[range] -> exclusive prefix in place, total -> lcount[task]
__global__ void __launch_bounds__(256)
part_prefix_kernel(const NodeTask *__restrict__ tasks, uint32_t *blockcnt, uint32_t *lcount) {
  const NodeTask t = tasks[blockIdx.x];
  const uint32_t nb = (t.n + kPartItems - 1) / kPartItems;
  uint32_t *bc = blockcnt + t.part_blk0;
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t b0 = 0; b0 < nb; b0 += 256) {
    const uint32_t b = b0 + threadIdx.x;
    const uint32_t v = b < nb ? bc[b] : 0u;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int) lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t add = carry;
    for (uint32_t k = 0; k < warp; ++k) add += wsum[k];
    if (b < nb) bc[b] = add + inc - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = add + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) lcount[blockIdx.x] = carry;
}

template <typename BinT>
__global__ void __launch_bounds__(256)
part_scatter_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint4 *__restrict__ panels,
                    size_t N, const uint32_t *__restrict__ ids0, const uint32_t *__restrict__ ids1,
                    uint32_t *out0, uint32_t *out1, const uint32_t *__restrict__ blockcnt,
                    const uint32_t *__restrict__ lcount) {
  __shared__ uint32_t s_task;
  __shared__ uint32_t wcnt[8];
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, blockIdx.x, false);
  __syncthreads();
  const NodeTask t = tasks[s_task];
  const uint32_t *src = t.src == 1 ? ids1 : ids0;
  uint32_t *dst = t.dst == 1 ? out1 : out0;
  const uint32_t lc = lcount[s_task];
  const uint32_t b0 = (blockIdx.x - t.part_blk0) * kPartItems, e = min(t.n, b0 + kPartItems);
  uint32_t left_run = blockcnt[blockIdx.x];   // lefts before this block (exclusive prefix)
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t r0 = b0; r0 < e; r0 += 256) {
    const uint32_t i = r0 + threadIdx.x;
    const bool act = i < e;
    uint32_t d = 0;
    bool goes_left = false;
    if (act) {
      d = t.src == 2 ? t.lo + i : src[t.lo + i];
      goes_left = load_bin<BinT>(panels, N, t.f, d) <= t.t;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, goes_left);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (uint32_t k = 0; k < 8; ++k) { if (k < warp) before += wcnt[k]; total += wcnt[k]; }
    const uint32_t lrank = left_run + before + __popc(bal & ((1u << lane) - 1u));
    if (act) {
      if (goes_left) dst[t.lo + lrank] = d;
      else dst[t.lo + lc + (i - lrank)] = d;   // rights before i = i - lefts before i
    }
    left_run += total;
    __syncthreads();
  }
}

__global__ void update_scores_kernel(const uint32_t *__restrict__ leaf_of_doc,
                                     const double *__restrict__ leafval, double weight, size_t N,
                                     double *scores) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) scores[i] = fma(weight, leafval[leaf_of_doc[i]], scores[i]);   // mart.cc:466 (fused)
}

// scores[i] += sum_t weight_t * tree_t(doc_i) for arbitrary trees expressed on this context's bins
// (Dart::update_modelscores, dart.cc:634-650; validation-set update of Mart::learn, mart.cc:356):
// one pass over the documents, trees in array order.
struct PackedNode { int32_t feature; uint32_t tidx; int32_t left, right; double value; };

template <typename BinT>
__global__ void apply_trees_kernel(const uint4 *__restrict__ panels, size_t N, const PackedNode *__restrict__ nodes,
                                   const uint32_t *__restrict__ root_of, const double *__restrict__ weights,
                                   uint32_t ntrees, double *scores) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double s = scores[i];
  for (uint32_t t = 0; t < ntrees; ++t) {
    const PackedNode *tn = nodes + root_of[t];
    int32_t nd = 0;
    while (tn[nd].feature >= 0)
      nd = load_bin<BinT>(panels, N, (uint32_t) tn[nd].feature, (uint32_t) i) <= tn[nd].tidx ? tn[nd].left : tn[nd].right;
    s = fma(weights[t], tn[nd].value, s);   // dart.cc:644-646 (fused like mart.cc:466)
  }
  scores[i] = s;
}

// Dart::update_contribution_scores (dart.cc:689-706): sum over the documents of |tree_t(doc)| (the unweighted leaf
// output) for each of `ntrees` trees; block partials in a fixed shape, reduced in block order by
// contrib_reduce_kernel, so the result does not depend on the launch.
constexpr uint32_t kContribThreads = 256;
template <typename BinT>
__global__ void __launch_bounds__(kContribThreads)
tree_contrib_kernel(const uint4 *__restrict__ panels, size_t N, const PackedNode *__restrict__ nodes,
                    const uint32_t *__restrict__ root_of, uint32_t ntrees, double *partials) {
  __shared__ double s_w[kContribThreads / 32];
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t t = 0; t < ntrees; ++t) {
    double v = 0.0;
    if (i < N) {
      const PackedNode *tn = nodes + root_of[t];
      int32_t nd = 0;
      while (tn[nd].feature >= 0)
        nd = load_bin<BinT>(panels, N, (uint32_t) tn[nd].feature, (uint32_t) i) <= tn[nd].tidx ? tn[nd].left : tn[nd].right;
      v = fabs(tn[nd].value);
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0;
      for (uint32_t w = 0; w < kContribThreads / 32; ++w) a += s_w[w];
      partials[(size_t) blockIdx.x * ntrees + t] = a;
    }
    __syncthreads();
  }
}
__global__ void contrib_reduce_kernel(const double *__restrict__ partials, uint32_t nblocks, uint32_t ntrees, double *out) {
  const uint32_t t = blockIdx.x, lane = threadIdx.x;   // one warp per tree
  double a = 0.0;
  for (uint32_t b = lane; b < nblocks; b += 32) a += partials[(size_t) b * ntrees + t];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) out[t] = a;
}

// ------------------------------------------------------------------------------------------
// Oblivious trees (ObliviousRT::fit / fill, ot.cc:32-201): per level, sum the split gain of every
// (f, t) over the level's nodes in node order; a cell is invalid as soon as one node violates the
// minimum leaf support; the single best cell (> 0, first maximum) splits every node.
// ------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void obv_level_kernel(const unsigned long long *__restrict__ hsum, const uint32_t *__restrict__ hcnt,
                                 uint32_t ncells, const int *__restrict__ slots, uint32_t nnodes,
                                 const uint32_t *__restrict__ thr_off, uint32_t F, uint32_t minls,
                                 const int *qexp, double *cell_score) {
  const uint32_t f = blockIdx.x;
  const uint32_t c0 = thr_off[f], cells = thr_off[f + 1] - c0;
  const double inv = EXACT ? 1.0 : ldexp(1.0, -*qexp);
  const double invalid = -DBL_MAX;
  for (uint32_t t = threadIdx.x; t < cells; t += blockDim.x) {
    double acc = 0.0;
    for (uint32_t k = 0; k < nnodes; ++k) {
      const unsigned long long *S = hsum + (size_t) slots[k] * ncells + c0;
      const uint32_t *C = hcnt + (size_t) slots[k] * ncells + c0;
      if (acc != invalid) {
        const uint32_t cn = C[cells - 1], lc = C[t], rc = cn - lc;
        if (lc >= minls && rc >= minls) {
          const double s = cell_value(EXACT, S[cells - 1], inv);
          const double ls = cell_value(EXACT, S[t], inv);
          const double rs = s - ls;
          acc += ls * ls / (double) lc + rs * rs / (double) rc;   // ot.cc:194-195
        } else {
          acc = invalid;
        }
      }
    }
    cell_score[c0 + t] = acc;
  }
}

// first maximum strictly greater than 0 (ot.cc:72-96); also gathers each node's left count
__global__ void obv_argmax_kernel(const double *__restrict__ cell_score, const uint32_t *__restrict__ thr_off,
                                  uint32_t F, const uint32_t *__restrict__ hcnt, uint32_t ncells,
                                  const int *__restrict__ slots, uint32_t nnodes, SplitResult *res,
                                  uint64_t *lcounts) {
  __shared__ double sb[256];
  __shared__ uint32_t sc[256];
  const uint32_t total = thr_off[F];
  double best = 0.0;
  uint32_t bc = 0xffffffffu;
  const uint32_t per = (total + blockDim.x - 1) / blockDim.x;
  const uint32_t b = threadIdx.x * per, e = min(total, b + per);
  for (uint32_t c = b; c < e; ++c) {
    const double v = cell_score[c];
    if (v != -DBL_MAX && v > best) { best = v; bc = c; }
  }
  sb[threadIdx.x] = best; sc[threadIdx.x] = bc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t k = 1; k < blockDim.x; ++k)
      if (sb[k] > best) { best = sb[k]; bc = sc[k]; }   // chunks ascend with k: first maximum wins
    SplitResult r{};
    r.score = best;
    r.valid = bc != 0xffffffffu && best != 0.0;
    if (r.valid) {
      uint32_t f = 0;
      while (thr_off[f + 1] <= bc) ++f;
      r.feature = f;
      r.threshold_idx = bc - thr_off[f];
      for (uint32_t k = 0; k < nnodes; ++k)
        lcounts[k] = hcnt[(size_t) slots[k] * ncells + bc];
    }
    res[0] = r;
  }
}

}  // namespace qr
