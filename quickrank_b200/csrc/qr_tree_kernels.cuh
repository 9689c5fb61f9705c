// quickrank_b200 — device kernels of tree growth (sm_100a): histograms, split scan, partition,
// leaf fit.  Every kernel works on a BATCH of node expansions described by NodeTask records, so
// one launch serves all the frontier nodes a growth round expands (qr_train.cu explains why that
// reproduces the reference's one-node-at-a-time heap order exactly).
#pragma once

#include "qr_kernels.cuh"
#include "qr_task.cuh"
#include "qr_grow.cuh"

namespace qr {

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
__device__ __forceinline__ uint32_t task_lcount(const NodeTask &t, const uint32_t *lcount, uint32_t task) {
  return t.whole ? 0u : (t.lc_known ? t.lcount : lcount[task]);
}

__device__ __forceinline__ void built_segment(const NodeTask &t, uint32_t lcount, uint32_t &begin,
                                              uint32_t &len) {
  if (t.whole) { begin = t.lo; len = t.n; }
  else if (t.build_left) { begin = t.lo; len = lcount; }
  else { begin = t.lo + lcount; len = t.n - lcount; }
}

// Clears the histogram slot each task builds into.  For the root refresh the per-bin counts never
// change from tree to tree ("count doesn't change, so no need to re-compute",
// rtnode_histogram.cc:149): they are copied from the table made at init instead of being recounted.
__global__ void prep_slots_kernel(const NodeTask *__restrict__ tasks, unsigned long long *hsum,
                                  uint32_t *hcnt, uint32_t ncells, const uint32_t *__restrict__ root_cnt) {
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

// Single-pass variant.  On one GPU the host knows every task's left count (lc_known) and the list order
// is preserved; in sharded training it does not (the split was chosen on all-reduced histograms), the
// right side is written from the end of the segment and the count is published by the last block:
// blocks take a ticket, count their own lefts, publish the count and obtain the number of lefts
// in the preceding blocks of the same task by decoupled look-back; the list order is preserved.
// Each block also clears its share of the histogram slot the task is about to build into.
// status word: epoch << 32 | flag << 30 | count   (flag 1: block aggregate, 2: inclusive prefix)
template <typename BinT>
__global__ void __launch_bounds__(256)
partition_onepass_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint4 *__restrict__ panels,
                         size_t N, const uint32_t *__restrict__ ids0, const uint32_t *__restrict__ ids1,
                         uint32_t *out0, uint32_t *out1, unsigned long long *status, uint32_t *ticket,
                         uint32_t ticket_base, uint32_t epoch, unsigned long long *hsum, uint32_t *hcnt,
                         uint32_t ncells, const RoundHdr *__restrict__ hdr, const __grid_constant__ TaskPack pack,
                         const long long *__restrict__ lamq, long long *__restrict__ lamq_c,
                         uint32_t *lcount_out, uint32_t *lcount_host) {
  if (pack.n) tasks = pack.t;
  __shared__ uint32_t s_vb, s_task, s_prefix;
  __shared__ uint32_t wc[8][8];
  // device-driven growth (qr_grow.cuh): the grid is an upper bound, the header has the real counts
  const uint32_t nblocks = hdr ? hdr->part_blocks : 0xffffffffu;
  if (hdr) ntasks = hdr->ntasks;
  if (threadIdx.x == 0) {
    s_vb = atomicAdd(ticket, 1u) - ticket_base;
    s_task = s_vb < nblocks ? find_task_by(tasks, ntasks, s_vb, false) : 0u;
  }
  __syncthreads();
  const uint32_t vb = s_vb;
  if (vb >= nblocks) return;
  const NodeTask t = tasks[s_task];
  const uint32_t lb = vb - t.part_blk0;
  const uint32_t nb = max(1u, (t.n + kPartItems - 1) / kPartItems);
  if (t.slotB >= 0) {
    const uint32_t chunk = (ncells + nb - 1) / nb;
    const uint32_t z0 = lb * chunk, z1 = min(ncells, z0 + chunk);
    unsigned long long *zs = hsum + (size_t) build_slot(t) * ncells;
    uint32_t *zc = hcnt + (size_t) build_slot(t) * ncells;
    for (uint32_t i = z0 + threadIdx.x; i < z1; i += 256) { zs[i] = 0ull; zc[i] = 0u; }
  }
  const uint32_t *src = t.src == 1 ? ids1 : ids0;
  uint32_t *dst = t.dst == 1 ? out1 : out0;
  const uint32_t b0 = lb * kPartItems, e = min(t.n, b0 + kPartItems);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t d[8], wr[8];
  uint32_t flags = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t i = b0 + r * 256 + threadIdx.x;
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
  // the fixed-point pseudo-responses of the documents that go to the child whose histogram is built are
  // copied next to the child's id list, so that the histogram kernel reads them coalesced instead of
  // gathering 8 bytes per document per panel; issued here, the loads complete during the look-back
  const bool compact = lamq_c != nullptr && t.slotB >= 0;
  long long lq[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint32_t i = b0 + r * 256 + threadIdx.x;
    const bool built = (((flags >> r) & 1u) != 0u) == (t.build_left != 0u);
    lq[r] = (compact && i < e && built) ? lamq[d[r]] : 0ll;
  }
  __syncthreads();
  if (warp == 0) {
    // warp-wide decoupled look-back: 32 predecessors of the same task per step
    uint32_t total = 0;
    for (int i = (int) lane; i < 64; i += 32) total += wc[i >> 3][i & 7];
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    volatile unsigned long long *st = status;
    const unsigned long long ep = (unsigned long long) epoch << 32;
    uint32_t prefix = 0;
    if (lb == 0) {
      if (lane == 0) st[vb] = ep | (2ull << 30) | total;
    } else {
      if (lane == 0) st[vb] = ep | (1ull << 30) | total;
      int hi = (int) vb - 1;                       // newest predecessor not yet accounted for
      const int first = (int) t.part_blk0;
      for (;;) {
        const int j = hi - (int) lane;
        unsigned long long v = 0;
        const bool in = j >= first;
        if (in) {
          do { v = st[j]; } while ((v >> 32) != epoch || ((v >> 30) & 3ull) == 0ull);
        }
        const uint32_t incl = __ballot_sync(0xffffffffu, in && ((v >> 30) & 3ull) == 2ull);
        const int stop = incl ? __ffs(incl) - 1 : 32;   // first lane (closest predecessor) with an inclusive prefix
        uint32_t part = (in && (int) lane <= stop) ? (uint32_t) (v & 0x3fffffffull) : 0u;
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        prefix += part;
        if (incl || hi - 32 < first) break;
        hi -= 32;
      }
      if (lane == 0) st[vb] = ep | (2ull << 30) | (prefix + total);
    }
    if (lane == 0) s_prefix = prefix;
    // sharded training: the local left count is not known beforehand; the task's last block publishes it
    // for the histogram kernel (device) and for the host's node records (mapped host memory)
    if (lane == 0 && !t.lc_known && lb == nb - 1) {
      lcount_out[s_task] = prefix + total;
      if (lcount_host) lcount_host[s_task] = prefix + total;
    }
  }
  __syncthreads();
  uint32_t run = s_prefix;
  const uint32_t lc = t.lcount;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = wc[r][w]; if (w < (int) warp) before += c; tot += c; }
    const uint32_t i = b0 + r * 256 + threadIdx.x;
    if (i < e) {
      const uint32_t lrank = run + before + wr[r];
      const bool left = (flags >> r) & 1u;
      // unknown left count: the right side is filled downwards from the end of the segment (it comes out in
      // descending order: fixed-point histograms do not depend on the order, the FP64 leaf sums only in
      // their last bits and deterministically)
      const uint32_t rpos = t.lc_known ? t.lo + lc + (i - lrank) : t.lo + t.n - 1u - (i - lrank);
      const uint32_t pos = left ? t.lo + lrank : rpos;
      dst[pos] = d[r];
      if (compact && left == (t.build_left != 0u)) lamq_c[pos] = lq[r];
    }
    run += tot;
  }
}

// ------------------------------------------------------------------------------------------
// FAST histograms: RTNodeHistogram::update / RTNodeHistogram(parent, sampleids, ...) scatter loops
// (rtnode_histogram.cc:51-58, 183-191) in 64-bit fixed point.  Shared memory has no native 64-bit
// add, so each cell is two 32-bit limbs updated with native shared atomics: the low limb's atomic
// returns the old value, which tells this very addition whether it carried into the high limb.
// One block per (document slice, panel).  Integer sums are order-independent: the result is
// deterministic and identical for any slicing (and any number of GPUs).
//
// Shared-memory layout and lane schedule (measured with scripts/hist_mb2.cu: the update loop runs
// at the shared-atomic issue rate, 16 lanes per clock per SM):
//  * cells are BIN-major: word index = bin * FPP + slot (slot = feature within the panel), three
//    word arrays (low limb | high limb | count);
//  * at step j lane L updates slot j ^ (L mod FPP): the 32 lanes of a warp touch every slot of the
//    panel twice per step, so no two lanes of a half-warp share a bank (bank = (bin & 1) * 16 +
//    slot for 8-bit bins) whatever the bins are, and at most two lanes can meet on one address;
//  * the row is permuted once per document (element j <- element j ^ rot) so that every extract
//    below has a compile-time position; there are no branches in the update loop;
//  * the rows of iteration t+1 are already in flight while iteration t updates shared memory.
// ------------------------------------------------------------------------------------------
// element j of the result = element j ^ r of v (elements of sizeof(BinT) bytes)
template <typename BinT>
__device__ __forceinline__ uint4 xor_permute(uint4 v, uint32_t r, uint32_t sel) {
  constexpr uint32_t WB = sizeof(BinT) == 1 ? 4u : 2u;   // bit of r that swaps neighbouring words
  if (r & WB) { uint32_t t = v.x; v.x = v.y; v.y = t; t = v.z; v.z = v.w; v.w = t; }
  if (r & (WB << 1)) { uint32_t t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
  uint4 o;
  o.x = __byte_perm(v.x, 0, sel); o.y = __byte_perm(v.y, 0, sel);
  o.z = __byte_perm(v.z, 0, sel); o.w = __byte_perm(v.w, 0, sel);
  return o;
}
template <typename BinT>
__device__ __forceinline__ uint32_t xor_permute_selector(uint32_t r) {
  if (sizeof(BinT) == 1) return 0x3210u ^ (0x1111u * (r & 3u));
  return (r & 1u) ? 0x1032u : 0x3210u;
}

// one document's (permuted) panel row into the block's limb histogram; cinc = 1 for a real
// document, 0 for the padding document of a thread's last, half-filled iteration (whose q is 0)
template <typename BinT, bool COUNT>
__device__ __forceinline__ void hist_add_row_smem(const uint4 &x, long long q, uint32_t cinc, unsigned char *rbp,
                                                  uint32_t hi_off, uint32_t cnt_off) {
  constexpr int FPP = kPanelBytes / sizeof(BinT);
  constexpr int H = FPP < 8 ? FPP : 8;
  constexpr int SH = sizeof(BinT) == 1 ? 6 : 5;            // log2(4 * FPP): bytes per bin row
  const uint32_t qlo = (uint32_t) q;
  const uint32_t qhi = (uint32_t) (q >> 32);
  const uint32_t rb = (uint32_t) (uintptr_t) rbp;           // only the low bits matter (xor below)
#pragma unroll
  for (int h0 = 0; h0 < FPP; h0 += H) {
    unsigned char *addr[H];
    uint32_t old[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const uint32_t xb = extract_bin<BinT>(x, h0 + j);
      addr[j] = rbp + ((xb << SH) + ((rb ^ (uint32_t) ((h0 + j) * 4)) - rb));
      old[j] = atomicAdd(reinterpret_cast<uint32_t *>(addr[j]), qlo);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const uint32_t carry = (old[j] + qlo) < old[j];
      atomicAdd(reinterpret_cast<uint32_t *>(addr[j] + hi_off), qhi + carry);
      if (COUNT) atomicAdd(reinterpret_cast<uint32_t *>(addr[j] + cnt_off), cinc);
    }
  }
}

// One block per SM (kHistThreads threads, one 16-feature limb histogram): every resident histogram is
// flushed with global atomics at the end of its block, so fewer, fatter blocks cut that cost; the
// update loop itself is bound by the shared-atomic issue rate, not by occupancy (scripts/hist_mb2.cu).
constexpr uint32_t kHistThreads = 512;

template <typename BinT, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(kHistThreads, 2)
hist_limb_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const uint32_t *__restrict__ lcount,
                 const uint4 *__restrict__ panels, size_t N, const uint32_t *__restrict__ ids0,
                 const uint32_t *__restrict__ ids1, const long long *__restrict__ lamq,
                 const uint32_t *__restrict__ thr_off, uint32_t F, unsigned long long *hsum,
                 uint32_t *hcnt, uint32_t ncells, ulonglong2 *sq_partials, uint32_t stride,
                 const RoundHdr *__restrict__ hdr, const __grid_constant__ TaskPack pack,
                 const long long *__restrict__ lamq_c) {
  if (pack.n) tasks = pack.t;
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint32_t s_base[FPP + 1];
  __shared__ uint32_t s_task;
  __shared__ U128 s_sq[kHistThreads / 32];
  if (hdr) {   // device-driven growth: upper-bound grid
    if (blockIdx.x >= hdr->hist_slices) return;
    ntasks = hdr->ntasks;
  }
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, blockIdx.x, true);
  __syncthreads();
  const NodeTask t = tasks[s_task];
  uint32_t seg0, seglen;
  built_segment(t, task_lcount(t, lcount, s_task), seg0, seglen);
  const uint32_t begin = (blockIdx.x - t.hist_blk0) * t.hist_dpb;
  const uint32_t p = blockIdx.y;
  if (begin >= seglen) {
    if (p == 0 && threadIdx.x == 0) sq_partials[blockIdx.x] = make_ulonglong2(0ull, 0ull);
    return;
  }
  const uint32_t end = min(seglen, begin + t.hist_dpb);

  const uint32_t f0 = p * FPP;
  const uint32_t nf = min(FPP, F - f0);
  const uint32_t cell0 = thr_off[f0];
  const uint32_t scells = SMEM ? FPP * stride : 0u;   // shared-memory cells, bin-major
  if (threadIdx.x <= FPP) s_base[threadIdx.x] = thr_off[f0 + min(threadIdx.x, nf)] - cell0;
  if (SMEM) {
    uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
    const uint32_t nz = scells * (COUNT ? 3u : 2u) / 4u;   // scells is a multiple of 8
    for (uint32_t i = threadIdx.x; i < nz; i += kHistThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  unsigned long long *gs = hsum + (size_t) build_slot(t) * ncells + cell0;
  uint32_t *gc = hcnt + (size_t) build_slot(t) * ncells + cell0;
  const bool identity = t.whole && t.src == 2;
  const uint32_t *ids = ((t.whole ? t.src : t.dst) == 1 ? ids1 : ids0) + seg0;
  const uint4 *prow = panels + (size_t) p * N;
  U128 sq{0ull, 0ull};
  if (SMEM) {
    const uint32_t rot = lane_id() & (FPP - 1);
    const uint32_t sel = xor_permute_selector<BinT>(rot);
    unsigned char *rbp = smem_raw + rot * 4u;
    const uint32_t hi_off = scells * 4u, cnt_off = scells * 8u;
    // pseudo-responses: gathered by document id, or — when the partition left a compacted copy next to
    // the id list — read by list position (coalesced)
    const bool byq = lamq_c != nullptr && !identity;
    const long long *lq = byq ? lamq_c + seg0 : lamq;
    uint32_t i = begin + threadIdx.x;
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
    long long q0 = 0, q1 = 0;
    bool v0 = i < end, v1 = i + kHistThreads < end;
    if (v0) { const uint32_t d = identity ? seg0 + i : ids[i]; c0 = prow[d]; q0 = lq[byq ? i : d]; }
    if (v1) { const uint32_t d = identity ? seg0 + i + kHistThreads : ids[i + kHistThreads]; c1 = prow[d]; q1 = lq[byq ? i + kHistThreads : d]; }
    bool w0 = i + 2 * kHistThreads < end, w1 = i + 3 * kHistThreads < end;
    uint32_t nd0 = 0, nd1 = 0;   // documents of the NEXT iteration
    if (w0) nd0 = identity ? seg0 + i + 2 * kHistThreads : ids[i + 2 * kHistThreads];
    if (w1) nd1 = identity ? seg0 + i + 3 * kHistThreads : ids[i + 3 * kHistThreads];
    while (v0) {
      uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
      long long nq0 = 0, nq1 = 0;
      if (w0) { n0 = prow[nd0]; nq0 = lq[byq ? i + 2 * kHistThreads : nd0]; }
      if (w1) { n1 = prow[nd1]; nq1 = lq[byq ? i + 3 * kHistThreads : nd1]; }
      i += 2 * kHistThreads;
      const bool z0 = i + 2 * kHistThreads < end, z1 = i + 3 * kHistThreads < end;
      if (z0) nd0 = identity ? seg0 + i + 2 * kHistThreads : ids[i + 2 * kHistThreads];
      if (z1) nd1 = identity ? seg0 + i + 3 * kHistThreads : ids[i + 3 * kHistThreads];
      if (p == 0) {   // squares_sum_ (rtnode_histogram.cc:65-69) as an exact integer
        const unsigned long long a0 = (unsigned long long) (q0 < 0 ? -q0 : q0);
        const unsigned long long a1 = (unsigned long long) (q1 < 0 ? -q1 : q1);
        u128_add(sq, a0 * a0, __umul64hi(a0, a0));
        u128_add(sq, a1 * a1, __umul64hi(a1, a1));
      }
      const uint4 x0 = xor_permute<BinT>(c0, rot, sel), x1 = xor_permute<BinT>(c1, rot, sel);
      hist_add_row_smem<BinT, COUNT>(x0, q0, 1u, rbp, hi_off, cnt_off);
      hist_add_row_smem<BinT, COUNT>(x1, q1, v1 ? 1u : 0u, rbp, hi_off, cnt_off);
      c0 = n0; c1 = n1; q0 = nq0; q1 = nq1; v0 = w0; v1 = w1; w0 = z0; w1 = z1;
    }
  } else {
    const uint32_t rot = lane_id() & (FPP - 1);
    const uint32_t rotb = rot * (uint32_t) sizeof(BinT);
    for (uint32_t i = begin + threadIdx.x; i < end; i += kHistThreads) {
      const uint32_t d = identity ? seg0 + i : ids[i];
      const uint4 row = rotate_bytes(prow[d], rotb);
      const long long q = lamq[d];
      if (p == 0) {
        const unsigned long long a = (unsigned long long) (q < 0 ? -q : q);
        u128_add(sq, a * a, __umul64hi(a, a));
      }
#pragma unroll
      for (int j = 0; j < (int) FPP; ++j) {
        const uint32_t slot = (j + rot) & (FPP - 1);
        if (slot < nf) {
          const uint32_t c = s_base[slot] + extract_bin<BinT>(row, j);
          atomicAdd(gs + c, (unsigned long long) q);
          if (COUNT) atomicAdd(gc + c, 1u);
        }
      }
    }
  }
  if (p == 0) {   // block total of the squares: integer, so any reduction shape gives the same value
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xffffffffu, sq.lo, o);
      const unsigned long long oh = __shfl_xor_sync(0xffffffffu, sq.hi, o);
      u128_add(sq, ol, oh);
    }
    if (lane_id() == 0) s_sq[threadIdx.x >> 5] = sq;
  }
  if (SMEM || p == 0) __syncthreads();
  if (p == 0 && threadIdx.x == 0) {
    U128 tot = s_sq[0];
    for (int w = 1; w < (int) kHistThreads / 32; ++w) u128_add(tot, s_sq[w].lo, s_sq[w].hi);
    sq_partials[blockIdx.x] = make_ulonglong2(tot.lo, tot.hi);
  }
  if (SMEM) {
    // flush: consecutive threads read consecutive shared cells (cell = bin * FPP + slot)
    const uint32_t *s_lo = reinterpret_cast<const uint32_t *>(smem_raw);
    const uint32_t *s_hi = s_lo + scells, *s_cnt = s_hi + scells;
    for (uint32_t i = threadIdx.x; i < scells; i += kHistThreads) {
      const uint32_t slot = i & (FPP - 1), bin = i / FPP;
      const long long v = ((long long) (int32_t) s_hi[i] << 32) + (long long) s_lo[i];
      const uint32_t cn = COUNT ? s_cnt[i] : 0u;
      // padding slots of the last panel collect the zero bins of their all-zero columns: dropped
      if (slot < nf && bin < s_base[slot + 1] - s_base[slot]) {
        if (v != 0) atomicAdd(gs + s_base[slot] + bin, (unsigned long long) v);
        if (COUNT && cn) atomicAdd(gc + s_base[slot] + bin, cn);
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
  built_segment(t, task_lcount(t, lcount, blockIdx.y), seg0, n);
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
  built_segment(t, task_lcount(t, lcount, blockIdx.x), seg0, n);
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
  built_segment(t, task_lcount(t, lcount, blockIdx.y), seg0, n);
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

constexpr uint32_t kFinWarps = 9;   // warps per finalize block
// Each feature is shared by kFinParts warps: every one of them forms the whole cumulative histogram
// (cheap), but evaluates the split score — two FP64 divisions per cell and child — only for its own
// third of the bins.  A (task, feature) is ~2300 dependent instructions for a single warp, and small
// growth rounds have far fewer (task, feature) pairs than the GPU has schedulers, so the serial
// chain, not throughput, sets the kernel's duration.
constexpr uint32_t kFinParts = 3;
__host__ __device__ inline uint32_t fin_blocks(uint32_t F) { return (F * kFinParts + kFinWarps - 1) / kFinWarps; }

template <bool EXACT, bool PEER = false>
__global__ void __launch_bounds__(kFinWarps * 32)
finalize_kernel(const NodeTask *__restrict__ tasks, unsigned long long *hsum, uint32_t *hcnt,
                uint32_t ncells, const uint32_t *__restrict__ thr_off, uint32_t F, uint32_t minls,
                const int *__restrict__ qexp, double *fbest_score, uint32_t *fbest_t, uint32_t *fbest_lc,
                ulonglong2 *totals, const ulonglong2 *__restrict__ sq128,
                const double *__restrict__ sq_exact, uint32_t *task_done, SplitResult *res,
                volatile uint32_t *host_flags, uint32_t round_id, const RoundHdr *__restrict__ hdr,
                const __grid_constant__ TaskPack pack, const __grid_constant__ PeerView pv) {
  if (pack.n) tasks = pack.t;
  const uint32_t task = blockIdx.y;
  if (hdr && task >= hdr->ntasks) return;   // device-driven growth: upper-bound grid
  const NodeTask t = tasks[task];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // Sharded training, fused exchange (all-reduce + split scan in one kernel): every rank accumulated the
  // built child's LOCAL histogram in a staging slot; this kernel is stream-ordered after that, so its
  // first block tells the peers "my staging slots of this round are complete", every block waits for the
  // same word from all peers, and the loads below add the W staging slots (NVLink loads from the peers'
  // pools) instead of reading one.  Integer sums: every rank obtains the same totals in any order.
  // Nothing is written to a peer, and staging slots alternate between two sets by round, so no second
  // barrier is needed (a rank can be at most one round ahead of its slowest peer).
  const int W = (PEER && !EXACT && pv.world > 1 && t.stage1) ? pv.world : 1;   // PEER = false: compiled out
  if (W > 1) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < (uint32_t) W && threadIdx.x != (uint32_t) pv.rank)
      st_flag(pv.peer_flags[threadIdx.x] + pv.rank, pv.epoch);
    if (threadIdx.x < (uint32_t) W && threadIdx.x != (uint32_t) pv.rank) wait_flag(pv.flags + threadIdx.x, pv.epoch);
    __syncthreads();
  }
  const uint32_t gw = blockIdx.x * kFinWarps + warp;
  const uint32_t f = gw / kFinParts, part = gw % kFinParts;
  const double inv = EXACT ? 1.0 : ldexp(1.0, -*qexp);
  const int nchild = t.whole ? 1 : 2;
  __shared__ double wb[2][kFinWarps];
  __shared__ uint32_t wt[2][kFinWarps], wt2[2][kFinWarps], wl2[2][kFinWarps];
  __shared__ uint32_t s_last;

  // Features with at most kFinChunks * 32 cells are processed entirely in registers: every load is
  // issued up front.  The kFinParts warps of a feature sit in the same block (kFinWarps is a multiple
  // of kFinParts): they all read the raw bins of the built child before any of them overwrites its
  // share with the cumulative values, hence the block barrier between loading and storing.
  static_assert(kFinWarps % kFinParts == 0, "the warps sharing a feature must share a block");
  constexpr int kFinChunks = 9;
  static_assert(kFinChunks % kFinParts == 0, "chunks are dealt to the parts in equal contiguous runs");
  constexpr int kOwn = kFinChunks / kFinParts;
  const bool active = f < F;
  const uint32_t c0 = active ? thr_off[f] : 0u, cells = active ? thr_off[f + 1] - c0 : 0u;
  unsigned long long *Bs = hsum + (size_t) t.slotB * ncells + c0;
  uint32_t *Bc = hcnt + (size_t) t.slotB * ncells + c0;
  // raw (not yet cumulative) bins of the built child: slotB itself, or the staging slot(s)
  const size_t roff = (size_t) build_slot(t) * ncells + c0;
  const unsigned long long *Rs = hsum + roff;
  const uint32_t *Rc = hcnt + roff;
  const bool two = nchild == 2;
  const unsigned long long *Ps = two ? hsum + (size_t) t.slotP * ncells + c0 : Bs;
  const uint32_t *Pc = two ? hcnt + (size_t) t.slotP * ncells + c0 : Bc;
  unsigned long long *Ds = two ? hsum + (size_t) t.slotD * ncells + c0 : Bs;
  uint32_t *Dc = two ? hcnt + (size_t) t.slotD * ncells + c0 : Bc;
  const bool regpath = active && cells <= kFinChunks * 32;
  unsigned long long bs[kFinChunks], ps[kFinChunks];
  uint32_t bc[kFinChunks], pc[kFinChunks];
  const uint32_t lastk = cells - 1;
  unsigned long long plast = 0ull, blast = 0ull;
  uint32_t pclast = 0u, bclast = 0u;
  if (regpath) {
#pragma unroll
    for (int ch = 0; ch < kFinChunks; ++ch) {
      const uint32_t k = ch * 32 + lane;
      const bool own = (uint32_t) (ch / kOwn) == part;
      const bool in = k < cells && (!EXACT || own);   // FAST: the prefix needs every bin of the built child
      bs[ch] = in ? Rs[k] : 0ull;
      bc[ch] = in ? Rc[k] : 0u;
      ps[ch] = (in && two && own) ? Ps[k] : 0ull;
      pc[ch] = (in && two && own) ? Pc[k] : 0u;
    }
    if (W > 1) {
      // the peers' staging slots, two peers at a time (all loads of a pair are in flight together)
      for (int q0 = 0; q0 < W - 1; q0 += 2) {
        const int pa = q0 + (q0 >= pv.rank ? 1 : 0);
        const bool has_b = q0 + 1 < W - 1;
        const int pb = has_b ? q0 + 1 + (q0 + 1 >= pv.rank ? 1 : 0) : pa;
        const volatile unsigned long long *sa = pv.sum[pa] + roff, *sb = pv.sum[pb] + roff;
        const volatile uint32_t *ca = pv.cnt[pa] + roff, *cb = pv.cnt[pb] + roff;
        unsigned long long va[kFinChunks], vb[kFinChunks];
        uint32_t na[kFinChunks], nb[kFinChunks];
#pragma unroll
        for (int ch = 0; ch < kFinChunks; ++ch) {
          const uint32_t k = ch * 32 + lane;
          const bool in = k < cells;
          va[ch] = in ? sa[k] : 0ull;
          vb[ch] = (in && has_b) ? sb[k] : 0ull;
          na[ch] = (in && pv.with_counts) ? ca[k] : 0u;
          nb[ch] = (in && has_b && pv.with_counts) ? cb[k] : 0u;
        }
#pragma unroll
        for (int ch = 0; ch < kFinChunks; ++ch) { bs[ch] += va[ch] + vb[ch]; bc[ch] += na[ch] + nb[ch]; }
      }
    }
    // last bins (node totals): parent's from memory, built child's from memory (EXACT) or the prefix
    plast = two ? Ps[lastk] : 0ull;
    pclast = two ? Pc[lastk] : 0u;
    blast = EXACT ? Bs[lastk] : 0ull;
    bclast = EXACT ? Bc[lastk] : 0u;
  }
  __syncthreads();
  if (active) {
    if (regpath) {
      // chunk ch holds cells ch * 32 + lane (coalesced).  The nine 5-step warp scans are independent
      // (their shuffles overlap); the running carry between chunks is added afterwards
      if (!EXACT) {   // inclusive prefix over bins (rtnode_histogram.cc:59-62), exact in fixed point
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
          for (int ch = 0; ch < kFinChunks; ++ch) {
            const unsigned long long pv = __shfl_up_sync(0xffffffffu, bs[ch], o);
            const uint32_t pcv = __shfl_up_sync(0xffffffffu, bc[ch], o);
            if ((int) lane >= o) { bs[ch] += pv; bc[ch] += pcv; }
          }
        }
        unsigned long long carry = 0;
        uint32_t carryc = 0;
#pragma unroll
        for (int ch = 0; ch < kFinChunks; ++ch) {
          const unsigned long long tot = __shfl_sync(0xffffffffu, bs[ch], 31);
          const uint32_t totc = __shfl_sync(0xffffffffu, bc[ch], 31);
          bs[ch] += carry; bc[ch] += carryc;
          carry += tot; carryc += totc;
          const uint32_t k = ch * 32 + lane;
          if (k < cells && (uint32_t) (ch / kOwn) == part) { Bs[k] = bs[ch]; Bc[k] = bc[ch]; }
        }
        blast = carry; bclast = carryc;   // bins past the last one are empty: the running total is the node total
      }
      for (int pass = 0; pass < nchild; ++pass) {
        if (pass == 1) {   // derived child = parent - built (rtnode_histogram.cc:79-85, 209-216)
#pragma unroll
          for (int ch = 0; ch < kFinChunks; ++ch) {
            const uint32_t k = ch * 32 + lane;
            if (EXACT) {
              const double pv = __longlong_as_double((long long) ps[ch]);
              const double bv = __longlong_as_double((long long) bs[ch]);
              bs[ch] = (unsigned long long) __double_as_longlong(pv - bv);   // rtnode_histogram.cc:82
            } else {
              bs[ch] = ps[ch] - bs[ch];
            }
            bc[ch] = pc[ch] - bc[ch];
            if (k < cells && (uint32_t) (ch / kOwn) == part) { Ds[k] = bs[ch]; Dc[k] = bc[ch]; }
          }
        }
        // node totals = last bin of the child being scanned
        unsigned long long sraw = blast;
        uint32_t cn = bclast;
        if (pass == 1) {
          if (EXACT) sraw = (unsigned long long) __double_as_longlong(__longlong_as_double((long long) plast) -
                                                                      __longlong_as_double((long long) blast));
          else sraw = plast - blast;
          cn = pclast - bclast;
        }
        const double s = cell_value(EXACT, sraw, inv);
        // split scan (rt.cc:272-291): strict '>' in ascending t, start value -1
        double best = -1.0;
        uint32_t best_t = 0xffffffffu, best_lc = 0;
#pragma unroll
        for (int ch = 0; ch < kFinChunks; ++ch) {
          const uint32_t k = ch * 32 + lane;
          const uint32_t lc = bc[ch], rc = cn - lc;
          if ((uint32_t) (ch / kOwn) == part && k < cells && lc >= minls && rc >= minls) {
            const double ls = cell_value(EXACT, bs[ch], inv);
            const double rs = s - ls;
            const double score = ls * ls / (double) lc + rs * rs / (double) rc;
            if (score > best) { best = score; best_t = k; best_lc = lc; }
          }
        }
        for (int o = 16; o > 0; o >>= 1) {   // arg-max, ties to the smaller t
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const uint32_t ot = __shfl_xor_sync(0xffffffffu, best_t, o);
          const uint32_t ol = __shfl_xor_sync(0xffffffffu, best_lc, o);
          if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; best_lc = ol; }
        }
        if (lane == 0) {
          // pass 0 scanned the built child, pass 1 the derived one; child 0 = left
          const int child = t.whole ? 0 : ((pass == 0) == (t.build_left != 0) ? 0 : 1);
          const size_t o = (((size_t) task * 2 + child) * F + f) * kFinParts + part;
          fbest_score[o] = best;
          fbest_t[o] = best_t;
          fbest_lc[o] = best_lc;
          // node size and sum are read from feature 0's last bin (rtnode.h:99-104)
          if (f == 0 && part == 0) totals[(size_t) task * 2 + child] = make_ulonglong2((unsigned long long) cn, sraw);
        }
      }
    } else if (part != 0) {
      // wide features (more than kFinChunks * 32 bins) are handled whole by part 0
      if (lane == 0)
        for (int child = 0; child < nchild; ++child) {
          const size_t o = (((size_t) task * 2 + child) * F + f) * kFinParts + part;
          fbest_score[o] = -1.0; fbest_t[o] = 0xffffffffu; fbest_lc[o] = 0;
        }
    } else {
      if (!EXACT) {
        long long carry = 0;
        uint32_t carryc = 0;
        for (uint32_t base = 0; base < cells; base += 32) {
          const uint32_t k = base + lane;
          long long v = k < cells ? (long long) Rs[k] : 0;
          uint32_t cv = k < cells ? Rc[k] : 0u;
          if (W > 1 && k < cells) {
            for (int p = 0; p < W; ++p) {
              if (p == pv.rank) continue;
              v += (long long) *(const volatile unsigned long long *) (pv.sum[p] + roff + k);
              if (pv.with_counts) cv += *(const volatile uint32_t *) (pv.cnt[p] + roff + k);
            }
          }
          for (int o = 1; o < 32; o <<= 1) {
            const long long pv = __shfl_up_sync(0xffffffffu, v, o);
            const uint32_t pcv = __shfl_up_sync(0xffffffffu, cv, o);
            if ((int) lane >= o) { v += pv; cv += pcv; }
          }
          v += carry; cv += carryc;
          if (k < cells) { Bs[k] = (unsigned long long) v; Bc[k] = cv; }
          carry = __shfl_sync(0xffffffffu, v, 31);
          carryc = __shfl_sync(0xffffffffu, cv, 31);
        }
      }
      __syncwarp();
      for (int pass = 0; pass < nchild; ++pass) {
        unsigned long long *S = Bs;
        uint32_t *C = Bc;
        if (pass == 1) {
          S = Ds;
          C = Dc;
          for (uint32_t k = lane; k < cells; k += 32) {
            if (EXACT) {
              const double pv = __longlong_as_double((long long) Ps[k]);
              const double bv = __longlong_as_double((long long) Bs[k]);
              S[k] = (unsigned long long) __double_as_longlong(pv - bv);
            } else {
              S[k] = Ps[k] - Bs[k];
            }
            C[k] = Pc[k] - Bc[k];
          }
          __syncwarp();
        }
        const unsigned long long sraw = S[cells - 1];
        const double s = cell_value(EXACT, sraw, inv);
        const uint32_t cn = C[cells - 1];
        double best = -1.0;
        uint32_t best_t = 0xffffffffu, best_lc = 0;
        for (uint32_t k = lane; k < cells; k += 32) {
          const uint32_t lc = C[k], rc = cn - lc;
          if (lc >= minls && rc >= minls) {
            const double ls = cell_value(EXACT, S[k], inv);
            const double rs = s - ls;
            const double score = ls * ls / (double) lc + rs * rs / (double) rc;
            if (score > best) { best = score; best_t = k; best_lc = lc; }
          }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const uint32_t ot = __shfl_xor_sync(0xffffffffu, best_t, o);
          const uint32_t ol = __shfl_xor_sync(0xffffffffu, best_lc, o);
          if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; best_lc = ol; }
        }
        if (lane == 0) {
          const int child = t.whole ? 0 : ((pass == 0) == (t.build_left != 0) ? 0 : 1);
          const size_t o = (((size_t) task * 2 + child) * F + f) * kFinParts;
          fbest_score[o] = best;
          fbest_t[o] = best_t;
          fbest_lc[o] = best_lc;
          if (f == 0) totals[(size_t) task * 2 + child] = make_ulonglong2((unsigned long long) cn, sraw);
        }
      }
    }
  }

  // The last block of a task to finish reduces the per-feature winners: arg-max over features
  // (first maximum wins: rt.cc:297-306 with GCC's static schedule) and the node statistics of
  // RTNode(sampleids, hist) (rtnode.h:97-107).
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(task_done + task, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // every load of this tail is issued before anything waits on one: squares partials, node totals and
  // the per-feature winners of BOTH children travel together (the tail is a chain of L2 round trips)
  const uint32_t FP = F * kFinParts;
  double sqB = 0.0;
  if (W > 1) {
    // exact squares of the built child over all ranks: warp 0 folds the W x hist_nblk 128-bit partials
    if (warp == 0) {
      U128 tot{0ull, 0ull};
      const uint32_t items = (uint32_t) W * t.hist_nblk;
      for (uint32_t it = lane; it < items; it += 32) {
        const uint32_t p = it / t.hist_nblk, i = it - p * t.hist_nblk;
        const volatile unsigned long long *v = reinterpret_cast<const volatile unsigned long long *>(pv.sq[p] + t.hist_blk0 + i);
        const unsigned long long lo = v[0], hi = v[1];
        u128_add(tot, lo, hi);
      }
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ol = __shfl_xor_sync(0xffffffffu, tot.lo, o);
        const unsigned long long oh = __shfl_xor_sync(0xffffffffu, tot.hi, o);
        u128_add(tot, ol, oh);
      }
      const double inv2 = ldexp(1.0, -2 * *qexp);
      sqB = ((double) tot.hi * 18446744073709551616.0 + (double) tot.lo) * inv2;
    }
  } else if (threadIdx.x < 2) {
    if (EXACT) {
      sqB = sq_exact[t.sq0];
    } else {
      U128 tot{0ull, 0ull};
      for (uint32_t i = 0; i < t.hist_nblk; ++i) { const ulonglong2 v = sq128[t.hist_blk0 + i]; u128_add(tot, v.x, v.y); }
      const double inv2 = ldexp(1.0, -2 * *qexp);
      sqB = ((double) tot.hi * 18446744073709551616.0 + (double) tot.lo) * inv2;
    }
  }
  ulonglong2 tv = make_ulonglong2(0ull, 0ull);
  if ((int) threadIdx.x < nchild) {
    const volatile ulonglong2 *tp = totals + (size_t) task * 2 + threadIdx.x;
    tv.x = tp->x; tv.y = tp->y;
  }
  // entries are ordered by (feature, part) = (feature, ascending threshold range): the first maximum
  // in that order is the reference's winner
  double best[2] = {-1.0, -1.0};
  uint32_t bf[2] = {0xffffffffu, 0xffffffffu}, bt[2] = {0xffffffffu, 0xffffffffu}, blc[2] = {0u, 0u};
  for (uint32_t ff = threadIdx.x; ff < FP; ff += kFinWarps * 32) {
#pragma unroll
    for (int child = 0; child < 2; ++child) {
      if (child < nchild) {
        const size_t o = ((size_t) task * 2 + child) * FP + ff;
        const double sc = ((const volatile double *) fbest_score)[o];
        const uint32_t tt = ((const volatile uint32_t *) fbest_t)[o];
        const uint32_t ll = ((const volatile uint32_t *) fbest_lc)[o];
        if (sc > best[child]) { best[child] = sc; bf[child] = ff; bt[child] = tt; blc[child] = ll; }
      }
    }
  }
#pragma unroll
  for (int child = 0; child < 2; ++child) {
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best[child], o);
      const uint32_t of = __shfl_xor_sync(0xffffffffu, bf[child], o);
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, bt[child], o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, blc[child], o);
      if (ob > best[child] || (ob == best[child] && of < bf[child])) { best[child] = ob; bf[child] = of; bt[child] = ot; blc[child] = ol; }
    }
    if (lane == 0) { wb[child][warp] = best[child]; wt[child][warp] = bf[child]; wt2[child][warp] = bt[child]; wl2[child][warp] = blc[child]; }
  }
  __syncthreads();
  if ((int) threadIdx.x < nchild) {   // thread c finishes child c
    const int child = (int) threadIdx.x;
    double b = wb[child][0];
    uint32_t f1 = wt[child][0], t1 = wt2[child][0], l1 = wl2[child][0];
    for (int w = 1; w < (int) kFinWarps; ++w)
      if (wb[child][w] > b || (wb[child][w] == b && wt[child][w] < f1)) { b = wb[child][w]; f1 = wt[child][w]; t1 = wt2[child][w]; l1 = wl2[child][w]; }
    const bool built = t.whole || ((child == 0) == (t.build_left != 0));
    SplitResult r;
    r.n = tv.x;
    r.sum = cell_value(EXACT, tv.y, inv);
    r.squares = built ? sqB : t.parent_squares - sqB;          // rtnode_histogram.cc:86,207
    r.deviance = r.squares - r.sum * r.sum / (double) r.n;      // rtnode.h:106
    r.score = b;
    r.valid = b != -1.0;
    r.feature = f1 == 0xffffffffu ? f1 : f1 / kFinParts;
    r.threshold_idx = r.valid ? t1 : 0xffffffffu;
    r.lcount = r.valid ? l1 : 0;
    r.pad = 0;
    res[(size_t) task * 2 + child] = r;
    if (host_flags) __threadfence_system();   // res lives in mapped host memory
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    task_done[task] = 0u;   // ready for the next round
    if (host_flags) host_flags[task] = round_id;   // publish to the polling host thread
  }
}

// ------------------------------------------------------------------------------------------
// Leaf outputs (RegressionTree::update_output, rt.cc:165-207) and score update
// (Mart::update_modelscores, mart.cc:459-468).
// ------------------------------------------------------------------------------------------


// FAST: per (leaf, chunk) partial sums with a fixed tree shape; also writes the doc -> leaf map.
__global__ void __launch_bounds__(256)
leaf_partial_kernel(const LeafSeg *__restrict__ segs, uint32_t nleaves, const uint32_t *__restrict__ ids0,
                    const uint32_t *__restrict__ ids1, const double *__restrict__ lam,
                    const double *__restrict__ wgt, double2 *partials, uint32_t *__restrict__ leaf_of_doc,
                    const RoundHdr *__restrict__ hdr) {
  __shared__ uint32_t s_leaf;
  __shared__ double p1[256], p2[256];
  if (hdr) {   // device-driven growth: upper-bound grid, counts in the header
    if (blockIdx.x >= hdr->leaf_blocks) return;
    nleaves = hdr->nleaves;
  }
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
                                  double *leafval, const RoundHdr *__restrict__ hdr) {
  if (hdr) nleaves = hdr->nleaves;
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

// The same for a packed set of trees: one pass over the documents, trees in array order.
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
