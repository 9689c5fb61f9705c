// quickrank_b200 — REFERENCE-mode (QR_HIST_REFERENCE) accumulation on sm_100a: every floating-point sum is formed in
// the order the reference's loops form it, so that every histogram cell, squares sum and leaf output carries the
// reference's own rounding (rtnode_histogram.cc:51-69, 183-203; rt.cc:165-207).  A sequentially rounded FP64 sum cannot
// be re-associated, but it can be spread over the machine:
//   * per-bin sums — one chain per (feature, bin).  hist_exact_list_kernel (small nodes): a warp per (feature, node)
//     walks the node's list 32 documents at a time, documents of a chunk that share a bin are added one after the other
//     in list order (__match_any_sync ranks them), cells in shared memory, next chunk's loads in flight.
//     hist_exact_walk_kernel (large nodes): a warp per (feature, bin) walks that bin's documents of the WHOLE dataset
//     in ascending document order — a per-feature list of documents sorted by (bin, document), made once at context
//     creation since the bins never change — and adds those that carry this round's mark of the node: 34 816 chains
//     (136 features x 256 bins) of N/256 additions instead of 136 chains of N.
//   * squares sum — ONE chain over the node's documents; its addends are non-negative, which makes the sequentially
//     rounded sum computable in parallel (ordered_squares_*): while the running sum stays inside one binade
//     [2^E, 2^(E+1)) it is an integer multiple of ulp = 2^(E-52), and adding a >= 0 in round-to-nearest-even moves that
//     integer by floor(a/ulp), +1 if the discarded part is above half an ulp or — exactly half — if the result would be
//     odd.  So each addend is a function (parity of the integer) -> increment, these functions compose associatively,
//     and a chunk of addends reduces to one pair (increment if even, increment if odd) without knowing the sum.  A last
//     sequential pass applies one pair per chunk; chunks in which the sum leaves its binade (a few dozen at most: the sum
//     only grows) are replayed addition by addition.  The result is bit-identical to the loop by construction — every
//     shortcut is checked at run time (binade at chunk entry, no overflow of the binade at chunk exit) and falls back to
//     the plain chain.
#pragma once

#include "qr_kernels.cuh"
#include "qr_task.cuh"

namespace qr {

// segment (in buffer `dst`, or identity when whole && src == 2) whose histogram is built
__device__ __forceinline__ uint32_t task_lcount(const NodeTask &t, const uint32_t *lcount, uint32_t task) {
  return t.whole ? 0u : (t.lc_known ? t.lcount : lcount[task]);
}

__device__ __forceinline__ void built_segment(const NodeTask &t, uint32_t lcount, uint32_t &begin,
                                              uint32_t &len) {
  if (t.whole) { begin = t.lo; len = t.n; }
  else if (t.build_left) { begin = t.lo; len = lcount; }
  else { begin = t.lo + lcount; len = t.n - lcount; }
}

// ------------------------------------------------------------------------------------------
// Small nodes: list order, one warp per (feature, task).  The slot must be zero on entry when the
// feature's cells do not fit the warp's shared-memory table (they are then accumulated in place).
// Ends with the sequential inclusive prefix over bins (rtnode_histogram.cc:59-62).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kExactWarps = 4;
constexpr uint32_t kExactSmemCells = 512;   // per warp: 8-bit bins always fit
constexpr uint32_t kListRanks = 4;          // documents of a bin added rank by rank before the per-bin chain takes over
constexpr int kListDepth = 4;               // 32-document chunks loaded together (the ids, then bins and pseudo-responses)

// fskip[f] != 0: the feature has a single occupied bin, so no threshold of it can leave documents on both sides
// (rt.cc:279 with a minimum leaf support >= 1) and its sums are never read (the node totals come from feature 0,
// rtnode.h:99-104): its cells stay zero, which the split scan rejects the same way (no document on either side).
template <typename BinT>
__global__ void __launch_bounds__(kExactWarps * 32)
hist_exact_list_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                       const uint4 *__restrict__ panels, size_t N, const uint32_t *__restrict__ ids0,
                       const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                       const uint32_t *__restrict__ thr_off, uint32_t F, const uint8_t *__restrict__ fskip,
                       unsigned long long *hsum, uint32_t *hcnt, uint32_t ncells) {
  __shared__ double s_sum[kExactWarps][kExactSmemCells];
  __shared__ uint32_t s_cnt[kExactWarps][kExactSmemCells];
  __shared__ double s_stage_v[kExactWarps][32];
  __shared__ uint32_t s_stage_b[kExactWarps][32], s_stage_p[kExactWarps][32];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t f = blockIdx.x * kExactWarps + warp;
  if (f >= F) return;                    // (only warp-level synchronisation below)
  const NodeTask t = tasks[blockIdx.y];
  if (t.walk) return;                    // large node: hist_exact_walk_kernel
  if (fskip != nullptr && fskip[f]) return;
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, blockIdx.y), seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  double *gsum = reinterpret_cast<double *>(hsum + (size_t) t.slotB * ncells) + thr_off[f];
  uint32_t *gcnt = hcnt + (size_t) t.slotB * ncells + thr_off[f];
  const uint32_t cells = thr_off[f + 1] - thr_off[f];
  const bool in_smem = cells <= kExactSmemCells;
  double *sum = in_smem ? s_sum[warp] : gsum;
  uint32_t *cnt = in_smem ? s_cnt[warp] : gcnt;
  if (in_smem) {
    for (uint32_t k = lane; k < cells; k += 32) { sum[k] = 0.0; cnt[k] = 0u; }
    __syncwarp();
  }
  constexpr int D = kListDepth;
  // two-stage pipeline: the ids of batch i + 2 and the bins / pseudo-responses of batch i + 1 are in flight while
  // batch i is added
  uint32_t d1[D], d2[D], b[D], nb[D];
  double v[D], nv[D];
  auto load_ids = [&](uint32_t base, uint32_t (&d)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const uint32_t k = base + (uint32_t) i * 32u + lane;
      d[i] = k < n ? (identity ? seg0 + k : ids[seg0 + k]) : 0xffffffffu;
    }
  };
  auto gather = [&](const uint32_t (&d)[D], uint32_t (&bb)[D], double (&vv)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const bool in = d[i] != 0xffffffffu;
      bb[i] = in ? load_bin<BinT>(panels, N, f, d[i]) : 0xffffffffu;   // inactive lanes share a bin no document can have
      vv[i] = in ? lam[d[i]] : 0.0;
    }
  };
  load_ids(0u, d1);
  gather(d1, b, v);
  load_ids(32u * D, d1);
  for (uint32_t base = 0; base < n; base += 32u * D) {
    gather(d1, nb, nv);
    load_ids(base + 64u * D, d2);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const bool act = b[i] != 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, b[i]);
      const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
      const uint32_t maxr = __reduce_max_sync(0xffffffffu, act ? (uint32_t) __popc(peers) : 0u);
      // the first kListRanks documents of every bin, one rank per step (list order within a bin = rank order)
      const uint32_t steps = min(maxr, kListRanks);
      for (uint32_t r = 0; r < steps; ++r) {
        if (act && rank == r) { sum[b[i]] += v[i]; cnt[b[i]] += 1u; }
        __syncwarp();
      }
      if (maxr > kListRanks) {   // crowded bins (skewed features): the rest of each as one register chain, the chunk
        s_stage_v[warp][lane] = v[i];   // parked in shared memory (broadcast reads instead of shuffles)
        s_stage_b[warp][lane] = b[i];
        s_stage_p[warp][lane] = peers;
        uint32_t todo = __ballot_sync(0xffffffffu, act && rank == 0u && (uint32_t) __popc(peers) > kListRanks);
        __syncwarp();
        while (todo) {
          const int leader = __ffs(todo) - 1;
          todo &= todo - 1u;
          const uint32_t bb = s_stage_b[warp][leader];
          uint32_t pm = s_stage_p[warp][leader];
          const uint32_t extra = (uint32_t) __popc(pm) - kListRanks;
#pragma unroll
          for (uint32_t r = 0; r < kListRanks; ++r) pm &= pm - 1u;   // (already added above)
          double acc = sum[bb];
          for (uint32_t m = pm; m; m &= m - 1u) acc += s_stage_v[warp][__ffs(m) - 1];
          __syncwarp();
          if (lane == 0) { sum[bb] = acc; cnt[bb] += extra; }
          __syncwarp();
        }
      }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) { b[i] = nb[i]; v[i] = nv[i]; d1[i] = d2[i]; }
  }
  __syncwarp();
  if (lane == 0) {
    double run = sum[0];
    uint32_t rc = cnt[0];
    for (uint32_t k = 1; k < cells; ++k) {
      run = sum[k] + run; rc += cnt[k];     // sumlbl[t] += sumlbl[t-1]
      sum[k] = run; cnt[k] = rc;
    }
  }
  __syncwarp();
  if (in_smem)
    for (uint32_t k = lane; k < cells; k += 32) { gsum[k] = sum[k]; gcnt[k] = cnt[k]; }
}

// ------------------------------------------------------------------------------------------
// Large nodes: one warp per histogram cell walks the bin's documents of the whole dataset in ascending
// document order (perm: per feature, the documents sorted by (bin, document); cell_pos[c] = first entry of
// cell c, absolute, cell_pos[ncells] = F*N) and adds those marked for one of the K nodes of its group (blockIdx.y):
// mark[doc] = tag << 4 | index in the walk list.  K == 0: the node is the dataset (root), no marks.
// Writes the RAW per-bin sums and counts; hist_exact_prefix_kernel makes them cumulative.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kWalkMax = 16;     // large nodes per launch: built segments are disjoint, so at most N / threshold
struct WalkList { uint32_t n; uint32_t idx[kWalkMax]; };

constexpr uint32_t kWalkWarps = 8;    // cells per block
constexpr int kWalkDepth = 8;         // 32-entry chunks per batch: one memory round trip per 256 entries
constexpr uint32_t kWalkBatch = 32u * kWalkDepth;

// One WARP per cell.  The additions of a cell are one sequential chain, so everything else is kept off it: the lanes
// load a batch of 256 list entries, gather their marks and pseudo-responses and park the pseudo-responses of the
// documents that belong to the node, compacted in list order, in shared memory; then every lane runs the same chain
// of unconditional additions over the parked addends (broadcast reads, no shuffles, no selects; the tail of the last
// group of 16 adds +0.0: the running sum starts at +0.0 and can therefore never be -0.0, so that leaves it bit for
// bit), while the gathers of the next batch and the list entries of the one after are in flight.  A cell costs one
// dependent FP64 addition per document of the node in it, plus the memory round trips of the whole cell.
template <int K>   // nodes accumulated per walk (1 or 2); K == 0: the node is the whole dataset
__global__ void __launch_bounds__(kWalkWarps * 32)
hist_exact_walk_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ WalkList wl,
                       const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ cell_pos,
                       const uint32_t *__restrict__ mark, uint32_t tag, const double *__restrict__ lam,
                       const uint32_t *__restrict__ thr_off, uint32_t F, const uint8_t *__restrict__ fskip,
                       unsigned long long *hsum, uint32_t *hcnt, uint32_t ncells) {
  constexpr bool WHOLE = K == 0;
  constexpr int KA = WHOLE ? 1 : K;
  constexpr int D = kWalkDepth;
  __shared__ double s_v[kWalkWarps][KA][kWalkBatch];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t c = blockIdx.x * kWalkWarps + warp;
  if (c >= ncells) return;
  const uint32_t g = blockIdx.y;
  double s[KA];
  uint32_t cn[KA];
#pragma unroll
  for (int k = 0; k < KA; ++k) { s[k] = 0.0; cn[k] = 0u; }
  const unsigned long long p0 = cell_pos[c];
  unsigned long long p1 = cell_pos[c + 1];
  if (fskip != nullptr) {   // single-bin features are not accumulated (see hist_exact_list_kernel)
    uint32_t lo = 0, hi = F - 1;
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (thr_off[mid] <= c) lo = mid; else hi = mid - 1; }
    if (fskip[lo]) p1 = p0;
  }
  uint32_t d1[D], d2[D], m[D];
  double v[D];
  auto load_ids = [&](unsigned long long pb, uint32_t (&d)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) { const unsigned long long p = pb + (unsigned long long) i * 32u + lane; d[i] = p < p1 ? perm[p] : 0xffffffffu; }
  };
  auto gather = [&](const uint32_t (&d)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const bool in = d[i] != 0xffffffffu;
      m[i] = (!WHOLE && in) ? mark[d[i]] : 0u;
      v[i] = in ? lam[d[i]] : 0.0;
    }
  };
  uint32_t nh[KA];   // addends parked for each node: only the documents that belong to it, in list order
  auto park = [&](const uint32_t (&d)[D]) {
#pragma unroll
    for (int k = 0; k < KA; ++k) nh[k] = 0u;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const bool in = d[i] != 0xffffffffu;
#pragma unroll
      for (int k = 0; k < KA; ++k) {
        const bool hit = in && (WHOLE || ((m[i] >> 4) == tag && (m[i] & 15u) == g * (uint32_t) KA + (uint32_t) k));
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (hit) s_v[warp][k][nh[k] + __popc(bal & ((1u << lane) - 1u))] = v[i];
        nh[k] += (uint32_t) __popc(bal);
      }
    }
#pragma unroll
    for (int k = 0; k < KA; ++k) cn[k] += nh[k];
  };
  load_ids(p0, d1);
  gather(d1);
  park(d1);
  load_ids(p0 + kWalkBatch, d1);
  __syncwarp();
  for (unsigned long long pb = p0; pb < p1; pb += kWalkBatch) {
    gather(d1);                               // batch b + 1
    load_ids(pb + 2ull * kWalkBatch, d2);     // batch b + 2
#pragma unroll
    for (int k = 0; k < KA; ++k) {            // batch b: the chain(s); past the parked addends +0.0, which changes nothing
#pragma unroll 1
      for (uint32_t j = 0; j < nh[k]; j += 16) {
        double tv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) tv[u] = j + u < nh[k] ? s_v[warp][k][j + u] : 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u) s[k] = s[k] + tv[u];
      }
    }
    __syncwarp();
    park(d1);
#pragma unroll
    for (int i = 0; i < D; ++i) d1[i] = d2[i];
    __syncwarp();
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < KA; ++k) {
      const uint32_t w = g * (uint32_t) KA + k;
      if (w < wl.n) {
        const NodeTask &t = tasks[wl.idx[w]];
        hsum[(size_t) t.slotB * ncells + c] = (unsigned long long) __double_as_longlong(s[k]);
        hcnt[(size_t) t.slotB * ncells + c] = cn[k];
      }
    }
  }
}

// marks the documents of the built segment of every large node of the round
__global__ void mark_walk_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ WalkList wl,
                                 const uint32_t *__restrict__ lcount, const uint32_t *__restrict__ ids0,
                                 const uint32_t *__restrict__ ids1, uint32_t *mark, uint32_t tag) {
  const uint32_t w = blockIdx.y;
  const NodeTask &t = tasks[wl.idx[w]];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, wl.idx[w]), seg0, n);
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    mark[ids[seg0 + i]] = (tag << 4) | w;
}

// sequential inclusive prefix over the bins of one feature (rtnode_histogram.cc:59-62); one thread per (feature, node)
__global__ void __launch_bounds__(64)
hist_exact_prefix_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ WalkList wl,
                         const uint32_t *__restrict__ thr_off, uint32_t F, unsigned long long *hsum, uint32_t *hcnt,
                         uint32_t ncells) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const NodeTask &t = tasks[wl.idx[blockIdx.y]];
  double *sum = reinterpret_cast<double *>(hsum + (size_t) t.slotB * ncells) + thr_off[f];
  uint32_t *cnt = hcnt + (size_t) t.slotB * ncells + thr_off[f];
  const uint32_t cells = thr_off[f + 1] - thr_off[f];
  double run = sum[0];
  uint32_t rc = cnt[0];
  uint32_t k = 1;
  for (; k + 8 <= cells; k += 8) {
    double x[8];
    uint32_t y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { x[u] = sum[k + u]; y[u] = cnt[k + u]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { run = x[u] + run; rc += y[u]; sum[k + u] = run; cnt[k + u] = rc; }
  }
  for (; k < cells; ++k) { run = sum[k] + run; rc += cnt[k]; sum[k] = run; cnt[k] = rc; }
}

// context creation: lower bound of every bin in a feature's sorted bin column -> absolute position in perm
__global__ void cell_pos_kernel(const uint32_t *__restrict__ sorted_bins, size_t N, uint32_t f, uint32_t cell0,
                                uint32_t cells, unsigned long long *cell_pos) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= cells) return;
  size_t lo = 0, hi = N;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (sorted_bins[mid] < b) lo = mid + 1; else hi = mid;
  }
  cell_pos[cell0 + b] = (unsigned long long) f * N + lo;
}
template <typename BinT>
__global__ void bin_column_kernel(const uint4 *__restrict__ panels, size_t N, uint32_t f, uint32_t *keys, uint32_t *iota) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { keys[i] = load_bin<BinT>(panels, N, f, (uint32_t) i); iota[i] = (uint32_t) i; }
}

// ------------------------------------------------------------------------------------------
// squares_sum_ (rtnode_histogram.cc:65-69 fused multiply-add chain, 199-203 multiply then add), in list order.
// Short lists: the chain itself, one warp per task.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kSqSerialMax = 4096;   // longer lists take the parallel scheme below

__device__ __forceinline__ double sq_step(double acc, double v, bool fused) {
  return fused ? fma(v, v, acc) : __dadd_rn(acc, __dmul_rn(v, v));
}

__global__ void squares_exact_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                                     const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                                     const uint32_t *__restrict__ ids1, double *partials, int all) {
  const NodeTask t = tasks[blockIdx.x];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, blockIdx.x), seg0, n);
  if (!all && n > kSqSerialMax) return;
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t lane = lane_id();
  const bool fused = t.fused_sq != 0u;
  double acc = 0.0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    double v = 0.0;
    if (i < n) v = lam[identity ? seg0 + i : ids[seg0 + i]];
    const uint32_t cntk = min(32u, n - base);
    if (cntk == 32u) {
#pragma unroll
      for (uint32_t k = 0; k < 32u; ++k) acc = sq_step(acc, __shfl_sync(0xffffffffu, v, k), fused);
    } else {
      for (uint32_t k = 0; k < cntk; ++k) acc = sq_step(acc, __shfl_sync(0xffffffffu, v, k), fused);
    }
  }
  if (lane == 0) partials[t.sq0] = acc;
}

// ---- the parallel scheme ---------------------------------------------------------------------
constexpr uint32_t kSqChunk = 256;        // addends per chunk: 8 per lane of a warp
struct SqPair { unsigned long long d0, d1; };   // increment of the integer if it is even / odd on entry
struct __align__(16) SqChunk { unsigned long long d0, d1; double sum; int32_t e; uint32_t ok; };

__device__ __forceinline__ uint32_t sq_chunks_of(uint32_t n) { return n > kSqSerialMax ? (n + kSqChunk - 1) / kSqChunk : 0u; }

// the addend of one document as an exact integer M (128 bits) times 2^x
__device__ __forceinline__ void sq_addend(double v, bool fused, unsigned long long &hi, unsigned long long &lo, int &x) {
  if (fused) {   // the exact square enters the fused multiply-add
    const unsigned long long bits = (unsigned long long) __double_as_longlong(v) & 0x7fffffffffffffffull;
    const int ef = (int) (bits >> 52);
    const unsigned long long m = ef ? ((bits & 0xfffffffffffffull) | 0x10000000000000ull) : bits;
    const int e = ef ? ef - 1075 : -1074;          // v = m * 2^e
    lo = m * m; hi = __umul64hi(m, m);
    x = 2 * e;
  } else {       // the rounded product is added
    const double a = __dmul_rn(v, v);
    const unsigned long long bits = (unsigned long long) __double_as_longlong(a);
    const int ef = (int) (bits >> 52);
    lo = ef ? ((bits & 0xfffffffffffffull) | 0x10000000000000ull) : bits;
    hi = 0ull;
    x = ef ? ef - 1075 : -1074;
  }
}

// the function of one addend for a running sum in [2^E, 2^(E+1)): returns false when it cannot be expressed
// (the addend alone would leave the binade)
__device__ __forceinline__ bool sq_element(double v, bool fused, int E, SqPair &out) {
  unsigned long long hi, lo;
  int x;
  sq_addend(v, fused, hi, lo, x);
  out.d0 = out.d1 = 0ull;
  if ((hi | lo) == 0ull) return true;
  const int sh = (E - 52) - x;       // ulp = 2^(E-52); addend / ulp = M / 2^sh
  if (sh <= 0) {
    if (hi != 0ull || sh < -11 || (lo >> (53 + sh)) != 0ull) return false;
    out.d0 = out.d1 = lo << (-sh);
    return true;
  }
  if (sh >= 128) return true;        // M < 2^106: far below half an ulp
  unsigned long long q, qh, half, sticky;
  if (sh >= 64) {
    const int s2 = sh - 64;
    q = s2 ? hi >> s2 : hi; qh = 0ull;
    if (s2 == 0) { half = lo >> 63; sticky = lo << 1; }
    else { half = (hi >> (s2 - 1)) & 1ull; sticky = (s2 > 1 ? hi << (65 - s2) : 0ull) | lo; }
  } else {
    q = (lo >> sh) | (hi << (64 - sh)); qh = hi >> sh;
    half = (lo >> (sh - 1)) & 1ull;
    sticky = sh > 1 ? lo << (65 - sh) : 0ull;
  }
  if (qh != 0ull || (q >> 53) != 0ull) return false;
  if (half && sticky) { out.d0 = out.d1 = q + 1ull; }
  else if (half) { out.d0 = q + (q & 1ull); out.d1 = q + ((q + 1ull) & 1ull); }   // tie: to even
  else { out.d0 = out.d1 = q; }
  return true;
}

// b after a
__device__ __forceinline__ SqPair sq_compose(const SqPair &a, const SqPair &b) {
  SqPair r;
  r.d0 = a.d0 + ((a.d0 & 1ull) ? b.d1 : b.d0);
  r.d1 = a.d1 + (((a.d1 + 1ull) & 1ull) ? b.d1 : b.d0);
  return r;
}

__device__ __forceinline__ double sq_load(const NodeTask &t, uint32_t seg0, const uint32_t *ids, bool identity,
                                          const double *lam, uint32_t i) {
  return lam[identity ? seg0 + i : ids[seg0 + i]];
}

// step 1: approximate sum of every chunk (any order: it only predicts the binade).  One warp per chunk.
__global__ void __launch_bounds__(128)
ordered_squares_sums_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint32_t *__restrict__ lcount,
                            const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                            const uint32_t *__restrict__ ids1, SqChunk *chunks, uint32_t total_chunks) {
  const uint32_t ch = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
  if (ch >= total_chunks) return;
  uint32_t task = 0;
  while (task + 1 < ntasks && tasks[task + 1].sq_chunk0 <= ch) ++task;
  const NodeTask &t = tasks[task];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, task), seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const uint32_t c0 = (ch - t.sq_chunk0) * kSqChunk;
  double a = 0.0;
#pragma unroll
  for (uint32_t j = 0; j < kSqChunk / 32; ++j) {
    const uint32_t i = c0 + j * 32 + lane;
    if (i < n) { const double v = sq_load(t, seg0, ids, identity, lam, i); a += v * v; }
  }
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) chunks[ch].sum = a;
}

// step 2: one block per task: exclusive prefix of the chunk sums -> the binade each chunk is expected to start in
__global__ void __launch_bounds__(256)
ordered_squares_binade_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount, SqChunk *chunks) {
  const NodeTask &t = tasks[blockIdx.x];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, blockIdx.x), seg0, n);
  const uint32_t nch = sq_chunks_of(n);
  if (nch == 0u) return;
  SqChunk *cs = chunks + t.sq_chunk0;
  __shared__ double s_w[8];
  __shared__ double s_carry;
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t b0 = 0; b0 < nch; b0 += 256) {
    const uint32_t i = b0 + threadIdx.x;
    const double v = i < nch ? cs[i].sum : 0.0;
    double inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, inc, o);
      if ((int) lane >= o) inc += u;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    double add = s_carry;
    for (uint32_t k = 0; k < warp; ++k) add += s_w[k];
    if (i < nch) {
      const double start = add + inc - v;
      const unsigned long long bits = (unsigned long long) __double_as_longlong(start);
      const int ef = (int) (bits >> 52);
      cs[i].e = (ef >= 64 && ef < 2047) ? ef - 1023 : INT32_MIN;   // 0, tiny or non-finite: replay the chunk
    }
    __syncthreads();
    if (threadIdx.x == 255) s_carry = add + inc;
    __syncthreads();
  }
}

// step 3: the (even, odd) increments of every chunk for its expected binade.  One warp per chunk, 8 consecutive
// addends per lane, composed in list order.
__global__ void __launch_bounds__(128)
ordered_squares_pairs_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint32_t *__restrict__ lcount,
                             const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                             const uint32_t *__restrict__ ids1, SqChunk *chunks, uint32_t total_chunks) {
  const uint32_t ch = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
  if (ch >= total_chunks) return;
  uint32_t task = 0;
  while (task + 1 < ntasks && tasks[task + 1].sq_chunk0 <= ch) ++task;
  const NodeTask &t = tasks[task];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, task), seg0, n);
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const bool fused = t.fused_sq != 0u;
  const int E = chunks[ch].e;
  if (E == INT32_MIN) { if (lane == 0) chunks[ch].ok = 0u; return; }
  constexpr uint32_t PER = kSqChunk / 32;
  const uint32_t c0 = (ch - t.sq_chunk0) * kSqChunk + lane * PER;
  double v[PER];
#pragma unroll
  for (uint32_t j = 0; j < PER; ++j) v[j] = c0 + j < n ? sq_load(t, seg0, ids, identity, lam, c0 + j) : 0.0;
  SqPair acc{0ull, 0ull};
  bool ok = true;
#pragma unroll
  for (uint32_t j = 0; j < PER; ++j) {
    SqPair e;
    ok = sq_element(v[j], fused, E, e) && ok;
    acc = sq_compose(acc, e);
  }
  for (int o = 1; o < 32; o <<= 1) {   // lane l absorbs lane l + o (the later addends)
    SqPair b;
    b.d0 = __shfl_down_sync(0xffffffffu, acc.d0, o);
    b.d1 = __shfl_down_sync(0xffffffffu, acc.d1, o);
    if ((lane & (2 * o - 1)) == 0) acc = sq_compose(acc, b);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    // an increment beyond 2^53 leaves the binade whatever the sum is: the chunk is replayed
    chunks[ch].d0 = acc.d0; chunks[ch].d1 = acc.d1;
    chunks[ch].ok = (ok && (acc.d0 >> 53) == 0ull && (acc.d1 >> 53) == 0ull) ? 1u : 0u;
  }
}

// step 4: the chain over chunks, one warp per task (every lane carries the same sum).  A chunk's pair is applied
// when the sum really is in the binade the pair was made for and stays in it; otherwise its additions are replayed.
__global__ void __launch_bounds__(32)
ordered_squares_resolve_kernel(const NodeTask *__restrict__ tasks, const uint32_t *__restrict__ lcount,
                               const double *__restrict__ lam, const uint32_t *__restrict__ ids0,
                               const uint32_t *__restrict__ ids1, const SqChunk *__restrict__ chunks, double *partials,
                               unsigned long long *replayed) {
  const NodeTask &t = tasks[blockIdx.x];
  uint32_t seg0, n;
  built_segment(t, task_lcount(t, lcount, blockIdx.x), seg0, n);
  const uint32_t nch = sq_chunks_of(n);
  if (nch == 0u) return;
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = (t.whole ? t.src : t.dst) == 1 ? ids1 : ids0;
  const bool fused = t.fused_sq != 0u;
  const SqChunk *cs = chunks + t.sq_chunk0;
  const uint32_t lane = lane_id();
  double acc = 0.0;
  uint32_t nreplay = 0;
  for (uint32_t b0 = 0; b0 < nch; b0 += 32) {
    SqChunk mine{0ull, 0ull, 0.0, INT32_MIN, 0u};
    if (b0 + lane < nch) mine = cs[b0 + lane];
    const uint32_t cnt = min(32u, nch - b0);
    for (uint32_t k = 0; k < cnt; ++k) {
      const unsigned long long d0 = __shfl_sync(0xffffffffu, mine.d0, k);
      const unsigned long long d1 = __shfl_sync(0xffffffffu, mine.d1, k);
      const int e = __shfl_sync(0xffffffffu, mine.e, k);
      const uint32_t ok = __shfl_sync(0xffffffffu, mine.ok, k);
      const unsigned long long bits = (unsigned long long) __double_as_longlong(acc);
      bool done = false;
      if (ok && (int) (bits >> 52) - 1023 == e) {
        const unsigned long long m = (bits & 0xfffffffffffffull) | 0x10000000000000ull;
        const unsigned long long m2 = m + ((m & 1ull) ? d1 : d0);
        if ((m2 >> 53) == 0ull) {
          acc = __longlong_as_double((long long) ((bits & 0x7ff0000000000000ull) | (m2 & 0xfffffffffffffull)));
          done = true;
        }
      }
      if (!done) {   // replay the chunk's additions (uniform across the warp)
        ++nreplay;
        const uint32_t c0 = (b0 + k) * kSqChunk;
        for (uint32_t base = c0; base < min(n, c0 + kSqChunk); base += 32) {
          const uint32_t i = base + lane;
          const double v = i < n ? sq_load(t, seg0, ids, identity, lam, i) : 0.0;
          const uint32_t cntk = min(32u, n - base);
          for (uint32_t j = 0; j < cntk; ++j) acc = sq_step(acc, __shfl_sync(0xffffffffu, v, j), fused);
        }
      }
    }
  }
  if (lane == 0) {
    partials[t.sq0] = acc;
    if (replayed) atomicAdd(replayed, (unsigned long long) nreplay);
  }
}

// ------------------------------------------------------------------------------------------
// Leaf outputs (RegressionTree::update_output, rt.cc:165-207): one warp per leaf, sums in list order; batches of 256
// documents parked in shared memory and added by unconditional chains, as in hist_exact_walk_kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
leaf_exact_kernel(const LeafSeg *__restrict__ segs, const uint32_t *__restrict__ ids0,
                  const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                  const double *__restrict__ wgt, double *leafval, uint32_t *__restrict__ leaf_of_doc) {
  constexpr int D = kWalkDepth;
  __shared__ double s_v[kWalkBatch], s_w[kWalkBatch];
  const uint32_t leaf = blockIdx.x;
  const LeafSeg sg = segs[leaf];
  const uint32_t *ids = sg.buf == 1 ? ids1 : ids0;
  const uint32_t lane = lane_id();
  double s1 = 0.0, s2 = 0.0;
  uint32_t d1[D], d2[D];
  double v[D], w[D];
  auto load_ids = [&](uint32_t base, uint32_t (&d)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const uint32_t k = base + (uint32_t) i * 32u + lane;
      d[i] = k < sg.n ? (sg.buf == 2 ? sg.lo + k : ids[sg.lo + k]) : 0xffffffffu;
    }
  };
  auto gather = [&](const uint32_t (&d)[D]) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const bool in = d[i] != 0xffffffffu;
      if (in) leaf_of_doc[d[i]] = leaf;
      v[i] = in ? lam[d[i]] : 0.0;
      w[i] = (in && wgt) ? wgt[d[i]] : 0.0;
    }
  };
  auto park = [&]() {
#pragma unroll
    for (int i = 0; i < D; ++i) { s_v[i * 32 + lane] = v[i]; s_w[i * 32 + lane] = w[i]; }
  };
  load_ids(0u, d1);
  gather(d1);
  park();
  load_ids(kWalkBatch, d1);
  __syncwarp();
  for (uint32_t base = 0; base < sg.n; base += kWalkBatch) {
    gather(d1);
    load_ids(base + 2u * kWalkBatch, d2);
#pragma unroll 1
    for (uint32_t j = 0; j < kWalkBatch; j += 16) {
      double tv[16], tw[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { tv[u] = s_v[j + u]; tw[u] = s_w[j + u]; }
#pragma unroll
      for (int u = 0; u < 16; ++u) { s1 = s1 + tv[u]; s2 = s2 + tw[u]; }   // (+0.0 past the end of the list: no effect)
    }
    __syncwarp();
    park();
#pragma unroll
    for (int i = 0; i < D; ++i) d1[i] = d2[i];
    __syncwarp();
  }
  if (lane == 0) {
    if (wgt) leafval[leaf] = s2 >= DBL_EPSILON ? s1 / s2 : 0.0;   // rt.cc:200
    else leafval[leaf] = s1 / (double) sg.n;                      // rt.cc:178
  }
}

}  // namespace qr
