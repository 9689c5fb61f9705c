// quickrank_b200 — one launch per growth round (single GPU, fixed-point mode).
//
// A growth round is partition -> histogram -> split scan.  As three kernels each phase pays a launch
// boundary and its own ramp-up/drain, and from tree ~150 of a LambdaMART run a tree needs ~32 rounds of
// mostly small work: the rounds are latency, not throughput ("Growth rounds").
// Here the three phases are chained INSIDE one launch through per-task counters:
//   * blocks take a ticket; tickets [0, P) are partition blocks (one-pass, decoupled look-back),
//     tickets [P, P + slices x panels) are histogram blocks;
//   * a histogram block first clears its shared memory, then waits until every partition block of its
//     task has published (part_done[task] == blocks of the task).  Waiting is deadlock-free: tickets are
//     handed out in start order, so every block a waiter depends on is already running;
//   * the last histogram block of a (task, panel) runs the split scan of that panel's features (one warp
//     per feature), and the last panel of a task reduces over features and publishes the result to the
//     polling host thread — no third kernel.
// STATUS: opt-in (QR_FUSED_ROUNDS=1; QR_FUSE_PARTITION=1 additionally chains the partition).  Correct
// (the GPU parity suite passes on it) but slower than the three separate kernels on B200: one launch
// configuration has to serve three very different roles (see qr_train.cu where the switch is read).
//
// The arithmetic is that of partition_onepass_kernel / hist_limb_kernel / finalize_kernel, which stay in
// use for the root, for REFERENCE mode, for oblivious trees and for multi-GPU runs (where an all-reduce
// sits between histogram and scan).
#pragma once

#include "qr_tree_kernels.cuh"

namespace qr {

constexpr uint32_t kRoundThreads = kHistThreads;          // 512: 16 warps
constexpr uint32_t kRoundWarps = kRoundThreads / 32;
constexpr uint32_t kRoundPartRounds = kPartItems / kRoundThreads;   // 4 documents per thread

struct RoundCounters {       // all zero between rounds
  uint32_t *part_done;       // [max_tasks] partition blocks of the task that have finished
  uint32_t *panel_done;      // [max_tasks * npanels] histogram slices of (task, panel) that have flushed
  uint32_t *task_done;       // [max_tasks] panels of the task whose features have been scanned
};

// FAST-mode split scan of one feature by one warp (rt.cc:257-292 on cumulative histograms):
// prefix over the built child's bins, derived = parent - built, best threshold of both children.
__device__ __noinline__ void round_scan_feature(const NodeTask &t, uint32_t task, uint32_t f, unsigned long long *hsum,
                                                   uint32_t *hcnt, uint32_t ncells, const uint32_t *__restrict__ thr_off,
                                                   uint32_t F, uint32_t minls, double inv, double *fbest_score,
                                                   uint32_t *fbest_t, uint32_t *fbest_lc, ulonglong2 *totals) {
  const uint32_t lane = lane_id();
  const uint32_t c0 = thr_off[f], cells = thr_off[f + 1] - c0;
  unsigned long long *Bs = hsum + (size_t) t.slotB * ncells + c0;
  uint32_t *Bc = hcnt + (size_t) t.slotB * ncells + c0;
  const unsigned long long *Ps = hsum + (size_t) t.slotP * ncells + c0;
  const uint32_t *Pc = hcnt + (size_t) t.slotP * ncells + c0;
  unsigned long long *Ds = hsum + (size_t) t.slotD * ncells + c0;
  uint32_t *Dc = hcnt + (size_t) t.slotD * ncells + c0;
  constexpr int CH = 9;   // 32-cell chunks held in registers
  double best[2] = {-1.0, -1.0};
  uint32_t best_t[2] = {0xffffffffu, 0xffffffffu}, best_lc[2] = {0u, 0u};
  unsigned long long tot_s[2] = {0ull, 0ull};
  uint32_t tot_c[2] = {0u, 0u};
  if (cells <= CH * 32) {
    // (the parent's bins are fetched when the derived child is formed, not held: this code shares the
    // histogram role's 64-register budget)
    unsigned long long bs[CH];
    uint32_t bc[CH];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const uint32_t k = ch * 32 + lane;
      const bool in = k < cells;
      // the histogram blocks of other SMs wrote these cells with atomics: read them from L2
      bs[ch] = in ? __ldcg(Bs + k) : 0ull;
      bc[ch] = in ? __ldcg(Bc + k) : 0u;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const unsigned long long pv = __shfl_up_sync(0xffffffffu, bs[ch], o);
        const uint32_t pcv = __shfl_up_sync(0xffffffffu, bc[ch], o);
        if ((int) lane >= o) { bs[ch] += pv; bc[ch] += pcv; }
      }
    }
    unsigned long long carry = 0;
    uint32_t carryc = 0;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      const unsigned long long tot = __shfl_sync(0xffffffffu, bs[ch], 31);
      const uint32_t totc = __shfl_sync(0xffffffffu, bc[ch], 31);
      bs[ch] += carry; bc[ch] += carryc;
      carry += tot; carryc += totc;
      const uint32_t k = ch * 32 + lane;
      if (k < cells) { Bs[k] = bs[ch]; Bc[k] = bc[ch]; }
    }
    tot_s[0] = carry; tot_c[0] = carryc;
    const unsigned long long plast = __ldcg(Ps + cells - 1);
    const uint32_t pclast = __ldcg(Pc + cells - 1);
    tot_s[1] = plast - carry; tot_c[1] = pclast - carryc;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {   // derived child = parent - built (rtnode_histogram.cc:79-85, 209-216)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          const uint32_t k = ch * 32 + lane;
          const bool in = k < cells;
          bs[ch] = (in ? __ldcg(Ps + k) : 0ull) - bs[ch];
          bc[ch] = (in ? __ldcg(Pc + k) : 0u) - bc[ch];
          if (in) { Ds[k] = bs[ch]; Dc[k] = bc[ch]; }
        }
      }
      const double s = (double) (long long) tot_s[pass] * inv;
      const uint32_t cn = tot_c[pass];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {   // rt.cc:272-291: strict '>' in ascending t, start value -1
        const uint32_t k = ch * 32 + lane;
        const uint32_t lc = bc[ch], rc = cn - lc;
        if (k < cells && lc >= minls && rc >= minls) {
          const double ls = (double) (long long) bs[ch] * inv;
          const double rs = s - ls;
          const double score = ls * ls / (double) lc + rs * rs / (double) rc;
          if (score > best[pass]) { best[pass] = score; best_t[pass] = k; best_lc[pass] = lc; }
        }
      }
    }
  } else {
    long long carry = 0;
    uint32_t carryc = 0;
    for (uint32_t base = 0; base < cells; base += 32) {
      const uint32_t k = base + lane;
      long long v = k < cells ? (long long) __ldcg(Bs + k) : 0;
      uint32_t cv = k < cells ? __ldcg(Bc + k) : 0u;
      for (int o = 1; o < 32; o <<= 1) {
        const long long pv = __shfl_up_sync(0xffffffffu, v, o);
        const uint32_t pcv = __shfl_up_sync(0xffffffffu, cv, o);
        if ((int) lane >= o) { v += pv; cv += pcv; }
      }
      v += carry; cv += carryc;
      if (k < cells) {
        Bs[k] = (unsigned long long) v; Bc[k] = cv;
        Ds[k] = __ldcg(Ps + k) - (unsigned long long) v; Dc[k] = __ldcg(Pc + k) - cv;
      }
      carry = __shfl_sync(0xffffffffu, v, 31);
      carryc = __shfl_sync(0xffffffffu, cv, 31);
    }
    __syncwarp();
    tot_s[0] = (unsigned long long) carry; tot_c[0] = carryc;
    tot_s[1] = __ldcg(Ps + cells - 1) - (unsigned long long) carry; tot_c[1] = __ldcg(Pc + cells - 1) - carryc;
    for (int pass = 0; pass < 2; ++pass) {
      const unsigned long long *S = pass ? Ds : Bs;
      const uint32_t *C = pass ? Dc : Bc;
      const double s = (double) (long long) tot_s[pass] * inv;
      const uint32_t cn = tot_c[pass];
      for (uint32_t k = lane; k < cells; k += 32) {
        const uint32_t lc = C[k], rc = cn - lc;
        if (lc >= minls && rc >= minls) {
          const double ls = (double) (long long) S[k] * inv;
          const double rs = s - ls;
          const double score = ls * ls / (double) lc + rs * rs / (double) rc;
          if (score > best[pass]) { best[pass] = score; best_t[pass] = k; best_lc[pass] = lc; }
        }
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    double b = best[pass];
    uint32_t bt = best_t[pass], bl = best_lc[pass];
    for (int o = 16; o > 0; o >>= 1) {   // arg-max, ties to the smaller t
      const double ob = __shfl_xor_sync(0xffffffffu, b, o);
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ob > b || (ob == b && ot < bt)) { b = ob; bt = ot; bl = ol; }
    }
    if (lane == 0) {
      // pass 0 scanned the built child, pass 1 the derived one; child 0 = left
      const int child = (pass == 0) == (t.build_left != 0) ? 0 : 1;
      const size_t o = ((size_t) task * 2 + child) * F + f;
      fbest_score[o] = b;
      fbest_t[o] = bt;
      fbest_lc[o] = bl;
      // node size and sum are read from feature 0's last bin (rtnode.h:99-104)
      if (f == 0) totals[(size_t) task * 2 + child] = make_ulonglong2((unsigned long long) tot_c[pass], tot_s[pass]);
    }
  }
}

template <typename BinT>
__global__ void __launch_bounds__(kRoundThreads, 2)
round_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, uint32_t part_blocks, uint32_t hist_slices,
             const uint4 *__restrict__ panels, size_t N, uint32_t *ids0, uint32_t *ids1,
             const long long *__restrict__ lamq, const uint32_t *__restrict__ thr_off, uint32_t F, uint32_t npanels,
             unsigned long long *hsum, uint32_t *hcnt, uint32_t ncells, ulonglong2 *sq_partials, uint32_t stride,
             unsigned long long *status, uint32_t *ticket, uint32_t ticket_base, uint32_t epoch, RoundCounters cnt,
             uint32_t minls, const int *__restrict__ qexp, double *fbest_score, uint32_t *fbest_t, uint32_t *fbest_lc,
             ulonglong2 *totals, SplitResult *res, volatile uint32_t *host_flags, uint32_t round_id,
             const __grid_constant__ TaskPack pack) {
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t s_vb, s_task, s_prefix, s_flag;
  __shared__ uint32_t wc[kRoundPartRounds][kRoundWarps];
  __shared__ uint32_t s_base[FPP + 1];
  __shared__ U128 s_sq[kRoundWarps];
  __shared__ double wb[kRoundWarps];
  __shared__ uint32_t wt[kRoundWarps], wt2[kRoundWarps], wl2[kRoundWarps];
  if (pack.n) tasks = pack.t;
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const uint32_t vb = s_vb;

  // ======================= partition role (rt.cc:325-334) =======================
  if (vb < part_blocks) {
    if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, vb, false);
    __syncthreads();
    const uint32_t task = s_task;
    const NodeTask t = tasks[task];
    const uint32_t lb = vb - t.part_blk0;
    const uint32_t nb = max(1u, (t.n + kPartItems - 1) / kPartItems);
    {   // clear this block's share of the histogram slot the task builds into
      const uint32_t chunk = (ncells + nb - 1) / nb;
      const uint32_t z0 = lb * chunk, z1 = min(ncells, z0 + chunk);
      unsigned long long *zs = hsum + (size_t) t.slotB * ncells;
      uint32_t *zc = hcnt + (size_t) t.slotB * ncells;
      for (uint32_t i = z0 + threadIdx.x; i < z1; i += kRoundThreads) { zs[i] = 0ull; zc[i] = 0u; }
    }
    const uint32_t *src = t.src == 1 ? ids1 : ids0;
    uint32_t *dst = t.dst == 1 ? ids1 : ids0;
    const uint32_t b0 = lb * kPartItems, e = min(t.n, b0 + kPartItems);
    uint32_t d[kRoundPartRounds], wr[kRoundPartRounds];
    uint32_t flags = 0;
#pragma unroll
    for (int r = 0; r < (int) kRoundPartRounds; ++r) {
      const uint32_t i = b0 + r * kRoundThreads + threadIdx.x;
      bool left = false;
      d[r] = 0;
      if (i < e) {
        d[r] = t.src == 2 ? t.lo + i : src[t.lo + i];
        left = load_bin<BinT>(panels, N, t.f, d[r]) <= t.t;
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, left);
      if (lane == 0) wc[r][warp] = __popc(bal);
      wr[r] = __popc(bal & ((1u << lane) - 1u));
      flags |= (left ? 1u : 0u) << r;
    }
    __syncthreads();
    if (warp == 0) {
      // warp-wide decoupled look-back: 32 predecessors of the same task per step
      uint32_t total = 0;
      for (int i = (int) lane; i < (int) (kRoundPartRounds * kRoundWarps); i += 32) total += wc[i / kRoundWarps][i % kRoundWarps];
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      volatile unsigned long long *st = status;
      const unsigned long long ep = (unsigned long long) epoch << 32;
      uint32_t prefix = 0;
      if (lb == 0) {
        if (lane == 0) st[vb] = ep | (2ull << 30) | total;
      } else {
        if (lane == 0) st[vb] = ep | (1ull << 30) | total;
        int hi = (int) vb - 1;
        const int first = (int) t.part_blk0;
        for (;;) {
          const int j = hi - (int) lane;
          unsigned long long v = 0;
          const bool in = j >= first;
          if (in) {
            do { v = st[j]; } while ((v >> 32) != epoch || ((v >> 30) & 3ull) == 0ull);
          }
          const uint32_t incl = __ballot_sync(0xffffffffu, in && ((v >> 30) & 3ull) == 2ull);
          const int stop = incl ? __ffs(incl) - 1 : 32;
          uint32_t part = (in && (int) lane <= stop) ? (uint32_t) (v & 0x3fffffffull) : 0u;
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          prefix += part;
          if (incl || hi - 32 < first) break;
          hi -= 32;
        }
        if (lane == 0) st[vb] = ep | (2ull << 30) | (prefix + total);
      }
      if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    uint32_t run = s_prefix;
    const uint32_t lc = t.lcount;
#pragma unroll
    for (int r = 0; r < (int) kRoundPartRounds; ++r) {
      uint32_t before = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < (int) kRoundWarps; ++w) { const uint32_t c = wc[r][w]; if (w < (int) warp) before += c; tot += c; }
      const uint32_t i = b0 + r * kRoundThreads + threadIdx.x;
      if (i < e) {
        const uint32_t lrank = run + before + wr[r];
        if ((flags >> r) & 1u) dst[t.lo + lrank] = d[r];
        else dst[t.lo + lc + (i - lrank)] = d[r];
      }
      run += tot;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(cnt.part_done + task, 1u);
    return;
  }

  // ======================= histogram role (rtnode_histogram.cc:51-58) =======================
  const uint32_t hidx = vb - part_blocks;
  if (hidx >= hist_slices * npanels) return;
  const uint32_t slice = hidx / npanels, p = hidx % npanels;
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, slice, true);
  __syncthreads();
  const uint32_t task = s_task;
  const NodeTask t = tasks[task];
  uint32_t seg0, seglen;
  built_segment(t, t.lcount, seg0, seglen);
  const uint32_t begin = (slice - t.hist_blk0) * t.hist_dpb;
  const uint32_t end = min(seglen, begin + t.hist_dpb);
  const bool has_docs = begin < seglen;
  const uint32_t f0 = p * FPP;
  const uint32_t nf = min(FPP, F - f0);
  const uint32_t cell0 = thr_off[f0];
  const uint32_t scells = FPP * stride;
  if (threadIdx.x <= FPP) s_base[threadIdx.x] = thr_off[f0 + min(threadIdx.x, nf)] - cell0;
  if (has_docs) {
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    const uint32_t nz = scells * 3u / 4u;
    for (uint32_t i = threadIdx.x; i < nz; i += kRoundThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  // wait for the task's partition (its blocks hold lower tickets, so they are running or done)
  if (part_blocks != 0) {   // (0: the partition ran as its own kernel before this launch)
    if (threadIdx.x == 0) {
      const uint32_t need = max(1u, (t.n + kPartItems - 1) / kPartItems);
      volatile uint32_t *pd = cnt.part_done + task;
      while (*pd < need) __nanosleep(64);
    }
    __syncthreads();
    __threadfence();
  }
  unsigned long long *gs = hsum + (size_t) t.slotB * ncells + cell0;
  uint32_t *gc = hcnt + (size_t) t.slotB * ncells + cell0;
  U128 sq{0ull, 0ull};
  if (has_docs) {
    // the id list was just written by other SMs: read it from L2
    const uint32_t *ids = (t.dst == 1 ? ids1 : ids0) + seg0;
    const uint4 *prow = panels + (size_t) p * N;
    const uint32_t rot = lane & (FPP - 1);
    const uint32_t sel = xor_permute_selector<BinT>(rot);
    unsigned char *rbp = smem_raw + rot * 4u;
    const uint32_t hi_off = scells * 4u, cnt_off = scells * 8u;
    uint32_t i = begin + threadIdx.x;
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
    long long q0 = 0, q1 = 0;
    bool v0 = i < end, v1 = i + kRoundThreads < end;
    if (v0) { const uint32_t d = __ldcg(ids + i); c0 = prow[d]; q0 = lamq[d]; }
    if (v1) { const uint32_t d = __ldcg(ids + i + kRoundThreads); c1 = prow[d]; q1 = lamq[d]; }
    bool w0 = i + 2 * kRoundThreads < end, w1 = i + 3 * kRoundThreads < end;
    uint32_t nd0 = 0, nd1 = 0;
    if (w0) nd0 = __ldcg(ids + i + 2 * kRoundThreads);
    if (w1) nd1 = __ldcg(ids + i + 3 * kRoundThreads);
    while (v0) {
      uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
      long long nq0 = 0, nq1 = 0;
      if (w0) { n0 = prow[nd0]; nq0 = lamq[nd0]; }
      if (w1) { n1 = prow[nd1]; nq1 = lamq[nd1]; }
      i += 2 * kRoundThreads;
      const bool z0 = i + 2 * kRoundThreads < end, z1 = i + 3 * kRoundThreads < end;
      if (z0) nd0 = __ldcg(ids + i + 2 * kRoundThreads);
      if (z1) nd1 = __ldcg(ids + i + 3 * kRoundThreads);
      if (p == 0) {   // squares_sum_ (rtnode_histogram.cc:65-69) as an exact integer
        const unsigned long long a0 = (unsigned long long) (q0 < 0 ? -q0 : q0);
        const unsigned long long a1 = (unsigned long long) (q1 < 0 ? -q1 : q1);
        u128_add(sq, a0 * a0, __umul64hi(a0, a0));
        u128_add(sq, a1 * a1, __umul64hi(a1, a1));
      }
      const uint4 x0 = xor_permute<BinT>(c0, rot, sel), x1 = xor_permute<BinT>(c1, rot, sel);
      hist_add_row_smem<BinT, true>(x0, q0, 1u, rbp, hi_off, cnt_off);
      hist_add_row_smem<BinT, true>(x1, q1, v1 ? 1u : 0u, rbp, hi_off, cnt_off);
      c0 = n0; c1 = n1; q0 = nq0; q1 = nq1; v0 = w0; v1 = w1; w0 = z0; w1 = z1;
    }
  }
  if (p == 0) {
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xffffffffu, sq.lo, o);
      const unsigned long long oh = __shfl_xor_sync(0xffffffffu, sq.hi, o);
      u128_add(sq, ol, oh);
    }
    if (lane == 0) s_sq[warp] = sq;
  }
  __syncthreads();
  if (p == 0 && threadIdx.x == 0) {
    U128 tot = s_sq[0];
    for (int w = 1; w < (int) kRoundWarps; ++w) u128_add(tot, s_sq[w].lo, s_sq[w].hi);
    sq_partials[slice] = make_ulonglong2(tot.lo, tot.hi);
  }
  if (has_docs) {
    const uint32_t *s_lo = reinterpret_cast<const uint32_t *>(smem_raw);
    const uint32_t *s_hi = s_lo + scells, *s_cnt = s_hi + scells;
    for (uint32_t i = threadIdx.x; i < scells; i += kRoundThreads) {
      const uint32_t slot = i & (FPP - 1), bin = i / FPP;
      const long long v = ((long long) (int32_t) s_hi[i] << 32) + (long long) s_lo[i];
      const uint32_t cn = s_cnt[i];
      if (slot < nf && bin < s_base[slot + 1] - s_base[slot]) {
        if (v != 0) atomicAdd(gs + s_base[slot] + bin, (unsigned long long) v);
        if (cn) atomicAdd(gc + s_base[slot] + bin, cn);
      }
    }
  }
  // ---- last slice of this (task, panel): split scan of the panel's features ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_flag = atomicAdd(cnt.panel_done + (size_t) task * npanels + p, 1u) == t.hist_nblk - 1u ? 1u : 0u;
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const double inv = ldexp(1.0, -*qexp);
  for (uint32_t w = warp; w < nf; w += kRoundWarps)
    round_scan_feature(t, task, f0 + w, hsum, hcnt, ncells, thr_off, F, minls, inv, fbest_score, fbest_t, fbest_lc, totals);
  // ---- last panel of the task: arg-max over features, node statistics, publication ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt.panel_done[(size_t) task * npanels + p] = 0u;
    s_flag = atomicAdd(cnt.task_done + task, 1u) == npanels - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  U128 tot{0ull, 0ull};
  for (uint32_t i = 0; i < t.hist_nblk; ++i) { const ulonglong2 v = __ldcg(sq_partials + t.hist_blk0 + i); u128_add(tot, v.x, v.y); }
  const double inv2 = ldexp(1.0, -2 * *qexp);
  const double sqB = ((double) tot.hi * 18446744073709551616.0 + (double) tot.lo) * inv2;
  for (int child = 0; child < 2; ++child) {
    const double *fs = fbest_score + ((size_t) task * 2 + child) * F;
    const uint32_t *ft = fbest_t + ((size_t) task * 2 + child) * F;
    const uint32_t *fl = fbest_lc + ((size_t) task * 2 + child) * F;
    double best = -1.0;
    uint32_t bf = 0xffffffffu, bt = 0xffffffffu, blc = 0;
    for (uint32_t ff = threadIdx.x; ff < F; ff += kRoundThreads) {
      const double sc = __ldcg(fs + ff);
      if (sc > best) { best = sc; bf = ff; bt = __ldcg(ft + ff); blc = __ldcg(fl + ff); }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const uint32_t of = __shfl_xor_sync(0xffffffffu, bf, o);
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, bt, o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, blc, o);
      if (ob > best || (ob == best && of < bf)) { best = ob; bf = of; bt = ot; blc = ol; }
    }
    if (lane == 0) { wb[warp] = best; wt[warp] = bf; wt2[warp] = bt; wl2[warp] = blc; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int) kRoundWarps; ++w)
        if (wb[w] > best || (wb[w] == best && wt[w] < bf)) { best = wb[w]; bf = wt[w]; bt = wt2[w]; blc = wl2[w]; }
      const bool built = (child == 0) == (t.build_left != 0);
      const ulonglong2 tv = __ldcg(totals + (size_t) task * 2 + child);
      SplitResult r;
      r.n = tv.x;
      r.sum = (double) (long long) tv.y * inv;
      r.squares = built ? sqB : t.parent_squares - sqB;          // rtnode_histogram.cc:86,207
      r.deviance = r.squares - r.sum * r.sum / (double) r.n;      // rtnode.h:106
      r.score = best;
      r.valid = best != -1.0;
      r.feature = bf;
      r.threshold_idx = r.valid ? bt : 0xffffffffu;
      r.lcount = r.valid ? blc : 0;
      r.pad = 0;
      res[(size_t) task * 2 + child] = r;
      if (child == 1) {
        cnt.task_done[task] = 0u;   // ready for the next round
        cnt.part_done[task] = 0u;
        __threadfence_system();
        host_flags[task] = round_id;
      }
    }
    __syncthreads();
  }
}

}  // namespace qr
