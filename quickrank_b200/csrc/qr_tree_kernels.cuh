// quickrank_b200 — device kernels of tree growth (sm_100a): histograms, split scan, partition,
// leaf fit.  Every kernel works on a BATCH of node expansions described by NodeTask records, so
// one launch serves all the frontier nodes a growth round expands (qr_train.cu explains why that
// reproduces the reference's one-node-at-a-time heap order exactly).
#pragma once

#include "qr_kernels.cuh"

namespace qr {

// One node being expanded (RegressionTree::split, rt.cc:209-362): its document list is cut by
// bin(f, doc) <= t, the histogram of one child is built from that child's documents and the
// sibling's is derived as parent - built (rt.cc:337-347).
struct NodeTask {
  uint32_t lo, n;        // the node's segment of the id buffer (local documents)
  uint32_t src, dst;     // id buffers: 0 / 1, src == 2 means the identity list (root)
  uint32_t f, t;         // split feature and threshold index
  uint32_t build_left;   // 1: the left child's histogram is built from samples, 0: the right one's
  uint32_t whole;        // 1: histogram of the whole node (root refresh, mart.cc:335); no split
  int32_t slotP, slotB, slotD;  // histogram slots: parent, built child, derived child
  uint32_t part_blk0;    // first flat partition block of this task
  uint32_t hist_blk0;    // first flat histogram slice of this task
  uint32_t hist_dpb;     // documents per histogram slice
  uint32_t sq0;          // first squares partial of this task
  uint32_t fused_sq;     // REFERENCE: 1 = fma chain (child ctor), 0 = mul+add (root update)
  double parent_squares;
};

constexpr uint32_t kPartItems = 2048;   // documents per partition block
constexpr uint32_t kSqParts = 16;       // squares partials per task (FAST)

__device__ __forceinline__ uint32_t find_task_by(const NodeTask *tasks, uint32_t ntasks, uint32_t blk,
                                                 bool hist) {
  uint32_t lo = 0, hi = ntasks - 1;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    const uint32_t b0 = hist ? tasks[mid].hist_blk0 : tasks[mid].part_blk0;
    if (b0 <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// segment (in buffer `dst`, or identity when whole && src == 2) whose histogram is built
__device__ __forceinline__ void built_segment(const NodeTask &t, uint32_t lcount, uint32_t &begin,
                                              uint32_t &len) {
  if (t.whole) { begin = t.lo; len = t.n; }
  else if (t.build_left) { begin = t.lo; len = lcount; }
  else { begin = t.lo + lcount; len = t.n - lcount; }
}

__global__ void zero_slots_kernel(const NodeTask *__restrict__ tasks, unsigned long long *hsum,
                                  uint32_t *hcnt, uint32_t ncells) {
  const int slot = tasks[blockIdx.y].slotB;
  unsigned long long *s = hsum + (size_t) slot * ncells;
  uint32_t *c = hcnt + (size_t) slot * ncells;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += gridDim.x * blockDim.x) {
    s[i] = 0ull;
    c[i] = 0u;
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

// ------------------------------------------------------------------------------------------
// FAST histograms: RTNodeHistogram::update / RTNodeHistogram(parent, sampleids, ...) scatter loops
// (rtnode_histogram.cc:51-58, 183-191) in 64-bit fixed point.  Shared memory has no native 64-bit
// add, so each cell is two 32-bit limbs updated with native shared atomics: the low limb's atomic
// returns the old value, which tells this very addition whether it carried into the high limb.
// One block per (document slice, panel); lanes walk the panel's features in rotated order so that
// the lanes of a warp hit different features' cells.  Integer sums are order-independent: the
// result is deterministic and identical for any slicing (and any number of GPUs).
// ------------------------------------------------------------------------------------------
template <typename BinT, bool SMEM>
__global__ void __launch_bounds__(256)
hist_limb_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint32_t *__restrict__ lcount,
                 const uint4 *__restrict__ panels, size_t N, const uint32_t *__restrict__ ids0,
                 const uint32_t *__restrict__ ids1, const long long *__restrict__ lamq,
                 const uint32_t *__restrict__ thr_off, uint32_t F, unsigned long long *hsum,
                 uint32_t *hcnt, uint32_t ncells) {
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_base[FPP];
  __shared__ uint32_t s_task;
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, blockIdx.x, true);
  __syncthreads();
  const NodeTask t = tasks[s_task];
  uint32_t seg0, seglen;
  built_segment(t, t.whole ? 0u : lcount[s_task], seg0, seglen);
  const uint32_t begin = (blockIdx.x - t.hist_blk0) * t.hist_dpb;
  if (begin >= seglen) return;
  const uint32_t end = min(seglen, begin + t.hist_dpb);

  const uint32_t p = blockIdx.y;
  const uint32_t f0 = p * FPP;
  const uint32_t nf = min(FPP, F - f0);
  const uint32_t cell0 = thr_off[f0];
  const uint32_t cells = thr_off[f0 + nf] - cell0;
  uint32_t *s_lo = reinterpret_cast<uint32_t *>(smem_raw);
  int32_t *s_hi = reinterpret_cast<int32_t *>(s_lo + (SMEM ? cells : 0));
  uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_hi + (SMEM ? cells : 0));
  if (threadIdx.x < FPP) s_base[threadIdx.x] = threadIdx.x < nf ? thr_off[f0 + threadIdx.x] - cell0 : 0u;
  if (SMEM)
    for (uint32_t i = threadIdx.x; i < cells; i += 256) { s_lo[i] = 0u; s_hi[i] = 0; s_cnt[i] = 0u; }
  __syncthreads();

  unsigned long long *gs = hsum + (size_t) t.slotB * ncells + cell0;
  uint32_t *gc = hcnt + (size_t) t.slotB * ncells + cell0;
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t rot = lane_id() & (FPP - 1);
  const uint4 *prow = panels + (size_t) p * N;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    const uint32_t d = identity ? seg0 + i : ids[seg0 + i];
    const uint4 row = rotate_bytes(prow[d], rot * (uint32_t) sizeof(BinT));
    const long long q = lamq[d];
    if (SMEM) {
      const uint32_t qlo = (uint32_t) q;
      const int32_t qhi = (int32_t) (q >> 32);
      uint32_t cell[FPP], old[FPP];
#pragma unroll
      for (int j = 0; j < (int) FPP; ++j) {
        const uint32_t slot = (j + rot) & (FPP - 1);
        cell[j] = slot < nf ? s_base[slot] + extract_bin<BinT>(row, j) : 0xffffffffu;
        if (cell[j] != 0xffffffffu) old[j] = atomicAdd(s_lo + cell[j], qlo);
      }
#pragma unroll
      for (int j = 0; j < (int) FPP; ++j) {
        if (cell[j] != 0xffffffffu) {
          const int32_t carry = (old[j] + qlo) < old[j];
          atomicAdd(s_hi + cell[j], qhi + carry);
          atomicAdd(s_cnt + cell[j], 1u);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < (int) FPP; ++j) {
        const uint32_t slot = (j + rot) & (FPP - 1);
        if (slot < nf) {
          const uint32_t c = s_base[slot] + extract_bin<BinT>(row, j);
          atomicAdd(gs + c, (unsigned long long) q);
          atomicAdd(gc + c, 1u);
        }
      }
    }
  }
  if (SMEM) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cells; i += 256) {
      const uint32_t cn = s_cnt[i];
      if (cn) {
        const long long v = ((long long) s_hi[i] << 32) + (long long) s_lo[i];
        atomicAdd(gs + i, (unsigned long long) v);
        atomicAdd(gc + i, cn);
      }
    }
  }
}

// REFERENCE order: one warp per (feature, task) walks the node's documents in list order;
// documents of a 32-wide chunk that fall in the same bin are added one after the other in document
// order (__match_any_sync ranks them), so every per-bin FP64 sum sees its addends in the sequence
// the reference's loop does (rtnode_histogram.cc:51-58).  Then the sequential inclusive prefix
// over bins (rtnode_histogram.cc:59-62).  The slot must be zero on entry.
template <typename BinT>
__global__ void __launch_bounds__(128)
hist_exact_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                  const uint4 *__restrict__ panels, size_t N, const uint32_t *__restrict__ ids0,
                  const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                  const uint32_t *__restrict__ thr_off, uint32_t F, unsigned long long *hsum,
                  uint32_t *hcnt, uint32_t ncells) {
  const uint32_t f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= F) return;
  const NodeTask t = tasks[blockIdx.y];
  uint32_t seg0, n;
  built_segment(t, t.whole ? 0u : lcount[blockIdx.y], seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t lane = lane_id();
  double *sum = reinterpret_cast<double *>(hsum + (size_t) t.slotB * ncells) + thr_off[f];
  uint32_t *cnt = hcnt + (size_t) t.slotB * ncells + thr_off[f];
  const uint32_t cells = thr_off[f + 1] - thr_off[f];
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    const bool act = i < n;
    uint32_t b = 0xffffffffu;   // inactive lanes share a bin no document can have
    double v = 0.0;
    if (act) {
      const uint32_t d = identity ? seg0 + i : ids[seg0 + i];
      b = load_bin<BinT>(panels, N, f, d);
      v = lam[d];
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, b);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t maxr = act ? __popc(peers) : 0u;
    for (int o = 16; o > 0; o >>= 1) maxr = max(maxr, __shfl_xor_sync(0xffffffffu, maxr, o));
    for (uint32_t r = 0; r < maxr; ++r) {
      if (act && rank == r) { sum[b] += v; cnt[b] += 1u; }
      __syncwarp();
    }
  }
  __syncwarp();
  if (lane == 0)
    for (uint32_t k = 1; k < cells; ++k) { sum[k] += sum[k - 1]; cnt[k] += cnt[k - 1]; }
}

// squares_sum_ (rtnode_histogram.cc:65-69, 199-203), sequential in list order; one warp per task.
__global__ void squares_exact_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                                     const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                                     const uint32_t *__restrict__ ids1, double *partials) {
  const NodeTask t = tasks[blockIdx.x];
  uint32_t seg0, n;
  built_segment(t, t.whole ? 0u : lcount[blockIdx.x], seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t lane = lane_id();
  double acc = 0.0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    double v = 0.0;
    if (i < n) v = lam[identity ? seg0 + i : ids[seg0 + i]];
    const uint32_t cntk = min(32u, n - base);
    if (t.fused_sq) {
      for (uint32_t k = 0; k < cntk; ++k) { const double vk = __shfl_sync(0xffffffffu, v, k); acc = fma(vk, vk, acc); }
    } else {
      for (uint32_t k = 0; k < cntk; ++k) { const double vk = __shfl_sync(0xffffffffu, v, k); acc = __dadd_rn(acc, __dmul_rn(vk, vk)); }
    }
  }
  if (lane == 0) partials[t.sq0] = acc;
}

// FAST: deterministic two-level sum of squares (fixed partition into kSqParts chunks, fixed tree).
__global__ void __launch_bounds__(256)
squares_fast_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                    const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                    const uint32_t *__restrict__ ids1, double *partials) {
  __shared__ double part[256];
  const NodeTask t = tasks[blockIdx.y];
  uint32_t seg0, n;
  built_segment(t, t.whole ? 0u : lcount[blockIdx.y], seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t per = (n + kSqParts - 1) / kSqParts;
  const uint32_t b = blockIdx.x * per, e = min(n, b + per);
  double acc = 0.0;
  for (uint32_t i = b + threadIdx.x; i < e; i += 256) {
    const double v = lam[identity ? seg0 + i : ids[seg0 + i]];
    acc = fma(v, v, acc);
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) part[threadIdx.x] += part[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[t.sq0 + blockIdx.x] = part[0];
}

// ------------------------------------------------------------------------------------------
// Finalize: cumulative histograms, derived = parent - built (rtnode_histogram.cc:59-62, 79-85,
// 209-216) and the split scan of every (feature, threshold) (rt.cc:257-292).  One block per
// (feature, task).  Per-feature winners go to fbest_*[task][child][f], child 0 = left.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double cell_value(bool exact, unsigned long long raw, double inv) {
  return exact ? __longlong_as_double((long long) raw) : (double) (long long) raw * inv;
}

template <bool EXACT>
__global__ void __launch_bounds__(256)
finalize_kernel(const NodeTask *__restrict__ tasks, unsigned long long *hsum, uint32_t *hcnt,
                uint32_t ncells, const uint32_t *__restrict__ thr_off, uint32_t F, uint32_t minls,
                const int *__restrict__ qexp, double *fbest_score, uint32_t *fbest_t) {
  const uint32_t f = blockIdx.x, task = blockIdx.y;
  const NodeTask t = tasks[task];
  const uint32_t c0 = thr_off[f], cells = thr_off[f + 1] - c0;
  unsigned long long *Bs = hsum + (size_t) t.slotB * ncells + c0;
  uint32_t *Bc = hcnt + (size_t) t.slotB * ncells + c0;
  __shared__ long long w_sum[8];
  __shared__ uint32_t w_cnt[8];
  __shared__ long long carry_sum;
  __shared__ uint32_t carry_cnt;
  __shared__ double wb[8];
  __shared__ uint32_t wt[8];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;

  if (!EXACT) {
    // inclusive scan of the per-bin fixed-point sums and counts, tiles of 256 bins
    if (threadIdx.x == 0) { carry_sum = 0; carry_cnt = 0; }
    __syncthreads();
    for (uint32_t t0 = 0; t0 < cells; t0 += 256) {
      const uint32_t k = t0 + threadIdx.x;
      long long v = k < cells ? (long long) Bs[k] : 0;
      uint32_t cv = k < cells ? Bc[k] : 0u;
      for (int o = 1; o < 32; o <<= 1) {
        const long long pv = __shfl_up_sync(0xffffffffu, v, o);
        const uint32_t pc = __shfl_up_sync(0xffffffffu, cv, o);
        if ((int) lane >= o) { v += pv; cv += pc; }
      }
      if (lane == 31) { w_sum[warp] = v; w_cnt[warp] = cv; }
      __syncthreads();
      long long add = carry_sum;
      uint32_t addc = carry_cnt;
      for (uint32_t w = 0; w < warp; ++w) { add += w_sum[w]; addc += w_cnt[w]; }
      v += add; cv += addc;
      if (k < cells) { Bs[k] = (unsigned long long) v; Bc[k] = cv; }
      __syncthreads();
      if (threadIdx.x == 255) { carry_sum = v; carry_cnt = cv; }
      __syncthreads();
    }
  }
  __syncthreads();
  const double inv = EXACT ? 1.0 : ldexp(1.0, -*qexp);
  const int nchild = t.whole ? 1 : 2;

  for (int pass = 0; pass < nchild; ++pass) {
    unsigned long long *S = Bs;
    uint32_t *C = Bc;
    if (pass == 1) {
      const unsigned long long *Ps = hsum + (size_t) t.slotP * ncells + c0;
      const uint32_t *Pc = hcnt + (size_t) t.slotP * ncells + c0;
      S = hsum + (size_t) t.slotD * ncells + c0;
      C = hcnt + (size_t) t.slotD * ncells + c0;
      for (uint32_t k = threadIdx.x; k < cells; k += 256) {
        if (EXACT) {
          const double pv = __longlong_as_double((long long) Ps[k]);
          const double bv = __longlong_as_double((long long) Bs[k]);
          S[k] = (unsigned long long) __double_as_longlong(pv - bv);   // rtnode_histogram.cc:82
        } else {
          S[k] = Ps[k] - Bs[k];
        }
        C[k] = Pc[k] - Bc[k];
      }
      __syncthreads();
    }
    // split scan (rt.cc:272-291): strict '>' in ascending t, start value -1
    const double s = cell_value(EXACT, S[cells - 1], inv);
    const uint32_t cn = C[cells - 1];
    double best = -1.0;
    uint32_t best_t = 0xffffffffu;
    for (uint32_t k = threadIdx.x; k < cells; k += 256) {
      const uint32_t lc = C[k], rc = cn - lc;
      if (lc >= minls && rc >= minls) {
        const double ls = cell_value(EXACT, S[k], inv);
        const double rs = s - ls;
        const double score = ls * ls / (double) lc + rs * rs / (double) rc;
        if (score > best) { best = score; best_t = k; }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {   // arg-max, ties to the smaller t
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, best_t, o);
      if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; }
    }
    if (lane == 0) { wb[warp] = best; wt[warp] = best_t; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w)
        if (wb[w] > best || (wb[w] == best && wt[w] < best_t)) { best = wb[w]; best_t = wt[w]; }
      // pass 0 scanned the built child, pass 1 the derived one; child 0 = left
      const int child = t.whole ? 0 : ((pass == 0) == (t.build_left != 0) ? 0 : 1);
      fbest_score[((size_t) task * 2 + child) * F + f] = best;
      fbest_t[((size_t) task * 2 + child) * F + f] = best_t;
    }
    __syncthreads();
  }
}

// Arg-max over features (first maximum wins: rt.cc:297-306 with GCC's static schedule) and the
// node statistics of RTNode(sampleids, hist) (rtnode.h:97-107).  One block per task.
template <bool EXACT>
__global__ void __launch_bounds__(128)
finalize2_kernel(const NodeTask *__restrict__ tasks, const unsigned long long *__restrict__ hsum,
                 const uint32_t *__restrict__ hcnt, uint32_t ncells, const uint32_t *__restrict__ thr_off,
                 uint32_t F, const int *__restrict__ qexp, const double *__restrict__ fbest_score,
                 const uint32_t *__restrict__ fbest_t, const double *__restrict__ sq_partials, uint32_t n_sq,
                 SplitResult *res) {
  const uint32_t task = blockIdx.x;
  const NodeTask t = tasks[task];
  const double inv = EXACT ? 1.0 : ldexp(1.0, -*qexp);
  __shared__ double wb[4];
  __shared__ uint32_t wf[4];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const int nchild = t.whole ? 1 : 2;
  for (int child = 0; child < nchild; ++child) {
    // first maximum over features: strict '>' in ascending f
    double best = -1.0;
    uint32_t bf = 0xffffffffu;
    for (uint32_t f = threadIdx.x; f < F; f += 128) {
      const double sc = fbest_score[((size_t) task * 2 + child) * F + f];
      if (sc > best) { best = sc; bf = f; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const uint32_t of = __shfl_xor_sync(0xffffffffu, bf, o);
      if (ob > best || (ob == best && of < bf)) { best = ob; bf = of; }
    }
    if (lane == 0) { wb[warp] = best; wf[warp] = bf; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 4; ++w)
        if (wb[w] > best || (wb[w] == best && wf[w] < bf)) { best = wb[w]; bf = wf[w]; }
      double sqB = 0.0;
      for (uint32_t i = 0; i < n_sq; ++i) sqB += sq_partials[t.sq0 + i];
      const bool built = t.whole || ((child == 0) == (t.build_left != 0));
      const int slot = built ? t.slotB : t.slotD;
      const unsigned long long *S = hsum + (size_t) slot * ncells;
      const uint32_t *C = hcnt + (size_t) slot * ncells;
      SplitResult r;
      const uint32_t last0 = thr_off[1] - 1;
      r.n = C[last0];
      r.sum = cell_value(EXACT, S[last0], inv);
      r.squares = built ? sqB : t.parent_squares - sqB;          // rtnode_histogram.cc:86,207
      r.deviance = r.squares - r.sum * r.sum / (double) r.n;      // rtnode.h:106
      r.score = best;
      r.valid = best != -1.0;
      r.feature = bf;
      r.threshold_idx = r.valid ? fbest_t[((size_t) task * 2 + child) * F + bf] : 0xffffffffu;
      r.lcount = r.valid ? C[thr_off[bf] + r.threshold_idx] : 0;
      r.pad = 0;
      res[(size_t) task * 2 + child] = r;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Leaf outputs (RegressionTree::update_output, rt.cc:165-207) and score update
// (Mart::update_modelscores, mart.cc:459-468).
// ------------------------------------------------------------------------------------------
struct LeafSeg { uint32_t lo, n; uint32_t buf; uint32_t blk0; };  // buf 2 = identity (unsplit root)

constexpr uint32_t kLeafItems = 4096;   // documents per block of the FAST leaf pass

// FAST: per (leaf, chunk) partial sums with a fixed tree shape; also writes the doc -> leaf map.
__global__ void __launch_bounds__(256)
leaf_partial_kernel(const LeafSeg *__restrict__ segs, uint32_t nleaves, const uint32_t *__restrict__ ids0,
                    const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                    const double *__restrict__ wgt, double2 *partials, uint32_t *__restrict__ leaf_of_doc) {
  __shared__ uint32_t s_leaf;
  __shared__ double p1[256], p2[256];
  if (threadIdx.x == 0) {
    uint32_t lo = 0, hi = nleaves - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (segs[mid].blk0 <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_leaf = lo;
  }
  __syncthreads();
  const uint32_t leaf = s_leaf;
  const LeafSeg sg = segs[leaf];
  const uint32_t *ids = sg.buf == 1 ? ids1 : ids0;
  const uint32_t b0 = (blockIdx.x - sg.blk0) * kLeafItems, e = min(sg.n, b0 + kLeafItems);
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t i = b0 + threadIdx.x; i < e; i += 256) {
    const uint32_t d = sg.buf == 2 ? sg.lo + i : ids[sg.lo + i];
    leaf_of_doc[d] = leaf;
    s1 += lam[d];
    if (wgt) s2 += wgt[d];
  }
  p1[threadIdx.x] = s1; p2[threadIdx.x] = s2;
  __syncthreads();
  for (uint32_t st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) { p1[threadIdx.x] += p1[threadIdx.x + st]; p2[threadIdx.x] += p2[threadIdx.x + st]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = make_double2(p1[0], p2[0]);
}

__global__ void leaf_final_kernel(const LeafSeg *__restrict__ segs, uint32_t nleaves,
                                  const double2 *__restrict__ partials, bool newton, double2 *leafsum,
                                  double *leafval) {
  const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= nleaves) return;
  const LeafSeg sg = segs[leaf];
  const uint32_t nb = (sg.n + kLeafItems - 1) / kLeafItems;
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t b = 0; b < nb; ++b) { s1 += partials[sg.blk0 + b].x; s2 += partials[sg.blk0 + b].y; }
  leafsum[leaf] = make_double2(s1, s2);
  if (newton) leafval[leaf] = s2 >= DBL_EPSILON ? s1 / s2 : 0.0;   // rt.cc:200
  else leafval[leaf] = s1 / (double) sg.n;                         // rt.cc:178
}

// REFERENCE: one warp per leaf, sums in list order.
__global__ void __launch_bounds__(32)
leaf_exact_kernel(const LeafSeg *__restrict__ segs, const uint32_t *__restrict__ ids0,
                  const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                  const double *__restrict__ wgt, double *leafval, uint32_t *__restrict__ leaf_of_doc) {
  const uint32_t leaf = blockIdx.x;
  const LeafSeg sg = segs[leaf];
  const uint32_t *ids = sg.buf == 1 ? ids1 : ids0;
  const uint32_t lane = lane_id();
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t base = 0; base < sg.n; base += 32) {
    const uint32_t i = base + lane;
    double v = 0.0, w = 0.0;
    if (i < sg.n) {
      const uint32_t d = sg.buf == 2 ? sg.lo + i : ids[sg.lo + i];
      leaf_of_doc[d] = leaf;
      v = lam[d];
      if (wgt) w = wgt[d];
    }
    const uint32_t cntk = min(32u, sg.n - base);
    for (uint32_t k = 0; k < cntk; ++k) {
      s1 += __shfl_sync(0xffffffffu, v, k);
      s2 += __shfl_sync(0xffffffffu, w, k);
    }
  }
  if (lane == 0) {
    if (wgt) leafval[leaf] = s2 >= DBL_EPSILON ? s1 / s2 : 0.0;   // rt.cc:200
    else leafval[leaf] = s1 / (double) sg.n;                      // rt.cc:178
  }
}

__global__ void update_scores_kernel(const uint32_t *__restrict__ leaf_of_doc,
                                     const double *__restrict__ leafval, double weight, size_t N,
                                     double *scores) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) scores[i] = fma(weight, leafval[leaf_of_doc[i]], scores[i]);   // mart.cc:466 (fused)
}

// scores[i] += weight * tree(doc_i) for an arbitrary tree expressed on this context's bins
// (Dart::update_modelscores, dart.cc:634-650).
struct DevTree { const int32_t *feature; const uint32_t *tidx; const int32_t *left, *right; const double *value; };

template <typename BinT>
__global__ void apply_tree_kernel(const uint4 *__restrict__ panels, size_t N, DevTree t, double weight,
                                  double *scores) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int32_t nd = 0;
  while (t.feature[nd] >= 0)
    nd = load_bin<BinT>(panels, N, (uint32_t) t.feature[nd], (uint32_t) i) <= t.tidx[nd] ? t.left[nd] : t.right[nd];
  scores[i] = fma(weight, t.value[nd], scores[i]);
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
