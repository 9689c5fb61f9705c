// quickrank_b200 — FAST-mode (64-bit fixed-point) tree growth on sm_100a: route -> histogram -> split scan.
//
// The reference keeps a list of sample ids per node and cuts it in two at every split (rt.cc:325-334).
// Here a node is not a list: every document carries the id of the frontier node it currently sits in
// (node_of_doc, 2 bytes per document, L2-resident), and a growth round is three kernels over that array:
//   route_kernel      every document of a node being split moves to its child (node_of_doc[d] = child id);
//                     the documents that land in the child whose histogram is BUILT (the smaller one) are
//                     appended, with their fixed-point pseudo-response, to that task's compact list.  The
//                     append order is arbitrary: everything accumulated from the list is an integer sum.
//   hist_limb_kernel  per (slice of the compact list, 16-feature panel): limb histograms in shared memory
//                     (rtnode_histogram.cc:51-58 in fixed point), flushed with global atomics.
//   scan_kernel       per (feature, task): inclusive prefix over bins (rtnode_histogram.cc:59-62), sibling =
//                     parent - built (:79-85), split score of every threshold for both children
//                     (rt.cc:257-292), arg-max; the task's last block reduces over features and publishes the
//                     result to the polling host thread through mapped memory.
// There is no stable partition, hence no prefix over blocks and no look-back chain: the round's critical
// path is three short kernels.  Leaf outputs (rt.cc:186-207) are one deterministic pass over node_of_doc.
// REFERENCE mode (bit-exact accumulation order) keeps the list-based kernels of qr_tree_kernels.cuh.
#pragma once

#include "qr_kernels.cuh"
#include "qr_task.cuh"

namespace qr {

// QR_KTRACE=1 (development aid): every block stamps %globaltimer at a few points of a kernel; the host prints
// the distribution for one growth round (qr_tree_host.cuh).  ktrace == nullptr: one predicated-off store.
constexpr uint32_t kTraceStamps = 16;
__device__ __forceinline__ void kstamp(unsigned long long *ktrace, uint32_t block, uint32_t i) {
  if (ktrace != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ktrace[(size_t) block * kTraceStamps + i] = t;
  }
}

// Programmatic dependent launch (route -> histogram): the histogram kernel is launched while the route kernel
// still runs; its blocks become resident as SMs free up, clear their shared memory and look their task up,
// then wait here until the route kernel has completed and its writes are visible.  Without the launch
// attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// ------------------------------------------------------------------------------------------
// Route (replaces the sample-id partition of RegressionTree::split, rt.cc:325-334; the test
// bin(f, doc) <= t equals the reference's float test because thresholds ascend).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kRouteThreads = 512;
constexpr uint32_t kRouteDocs = 2048;     // documents per block: 4 consecutive documents per thread (fewer, fatter
                                          // blocks: the launch ramp of ~1000 small blocks was half the kernel)
constexpr uint32_t kRouteTasks = 256;     // node expansions per launch (wider rounds take several launches)
constexpr uint32_t kNoTask = 0xffffu;

struct RouteTask { uint32_t f, t, child0, region0; };   // child0: bit 31 = build_left, bit 30 = a histogram is built

// zero_slots: sharded training accumulates into staging slots the peers read, which are cleared here; on one
// GPU the split scan clears the raw slot as it consumes it (scan_pub_kernel).
template <typename BinT>
__global__ void __launch_bounds__(kRouteThreads, 4)   // 4 blocks per SM: 1M documents are one wave
route_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const __grid_constant__ TaskPack pack,
             const uint4 *__restrict__ panels, size_t N, uint16_t *node, uint32_t node_min, uint32_t node_range,
             const long long *__restrict__ lamq, uint32_t *__restrict__ cids, long long *__restrict__ clamq,
             uint32_t *counts, uint32_t *counts_clear, uint32_t nclear, unsigned long long *hsum, uint32_t *hcnt,
             uint32_t ncells, int zero_slots, unsigned long long *ktrace) {
  if (pack.n) tasks = pack.t;
  pdl_launch_dependents();
  kstamp(ktrace, blockIdx.x, 0);
  extern __shared__ uint16_t s_lut[];     // [node_range]: node id - node_min -> task of this launch
  __shared__ RouteTask s_t[kRouteTasks];
  __shared__ uint32_t s_cnt[kRouteTasks], s_base[kRouteTasks];
  __shared__ uint32_t s_doc[kRouteDocs], s_tr[kRouteDocs];
  __shared__ long long s_q[kRouteDocs];
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  // this thread's 4 documents: the load is in flight while the task tables are set up
  const size_t d0 = (size_t) blockIdx.x * kRouteDocs + (size_t) tid * 4u;
  ushort4 nn = make_ushort4(0xffffu, 0xffffu, 0xffffu, 0xffffu);
  if (d0 < N) nn = *reinterpret_cast<const ushort4 *>(node + d0);   // the array is padded to a multiple of 4
  for (uint32_t i = tid; i < node_range; i += kRouteThreads) s_lut[i] = (uint16_t) kNoTask;
  for (uint32_t j = tid; j < ntasks; j += kRouteThreads) {
    const NodeTask &t = tasks[j];
    s_t[j] = RouteTask{t.f, t.t, t.child0 | (t.build_left ? 0x80000000u : 0u) | (t.slotB >= 0 ? 0x40000000u : 0u), t.region0};
    s_cnt[j] = 0u;
  }
  if (tid == 0) s_total = 0u;
  // the counters of the NEXT round (the other half of the double buffer) are cleared here
  if (blockIdx.x == 0) for (uint32_t i = tid; i < nclear; i += kRouteThreads) counts_clear[i] = 0u;
  __syncthreads();
  for (uint32_t j = tid; j < ntasks; j += kRouteThreads) s_lut[tasks[j].node - node_min] = (uint16_t) j;
  if (zero_slots) {   // every block clears its share of the slots the round's histograms are accumulated into
    const unsigned long long total = (unsigned long long) ntasks * ncells;
    const unsigned long long per = (total + gridDim.x - 1) / gridDim.x;
    const unsigned long long z0 = (unsigned long long) blockIdx.x * per;
    const unsigned long long z1 = z0 + per < total ? z0 + per : total;
    if (z0 < z1) {
      for (uint32_t j = (uint32_t) (z0 / ncells); j <= (uint32_t) ((z1 - 1) / ncells); ++j) {
        const NodeTask &t = tasks[j];
        if (t.slotB < 0) continue;
        const unsigned long long b = (unsigned long long) j * ncells;
        const uint32_t lo = (uint32_t) ((z0 > b ? z0 : b) - b);
        const uint32_t hi = (uint32_t) ((z1 < b + ncells ? z1 : b + ncells) - b);
        unsigned long long *zs = hsum + (size_t) build_slot(t) * ncells;
        uint32_t *zc = hcnt + (size_t) build_slot(t) * ncells;
        for (uint32_t i = lo + tid; i < hi; i += kRouteThreads) { zs[i] = 0ull; zc[i] = 0u; }
      }
    }
  }
  __syncthreads();

  kstamp(ktrace, blockIdx.x, 1);
  uint32_t v[4] = {nn.x, nn.y, nn.z, nn.w};
  uint32_t tk[4];
  uint32_t bins[4];
  long long q[4];
  // all dependent loads of the 4 documents are issued together: the split bin, and the pseudo-response of a
  // document whose task builds a histogram (it is only needed if the document lands in the built child)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t d = d0 + i;
    const uint32_t idx = v[i] - node_min;
    tk[i] = kNoTask;
    bins[i] = 0u;
    q[i] = 0ll;
    if (d < N && idx < node_range) {
      tk[i] = s_lut[idx];
      if (tk[i] != kNoTask) {
        bins[i] = load_bin<BinT>(panels, N, s_t[tk[i]].f, (uint32_t) d);
        if (s_t[tk[i]].child0 & 0x40000000u) q[i] = lamq[d];
      }
    }
  }
  bool changed = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bool built = false;
    if (tk[i] != kNoTask) {
      const RouteTask rt = s_t[tk[i]];
      const bool left = bins[i] <= rt.t;
      v[i] = (rt.child0 & 0x3fffffffu) + (left ? 0u : 1u);
      changed = true;
      built = ((rt.child0 >> 30) & 1u) != 0u && left == ((rt.child0 >> 31) != 0u);
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, built);
    if (bal) {
      const int leader = __ffs(bal) - 1;
      uint32_t base = 0;
      if ((int) lane == leader) base = atomicAdd(&s_total, (uint32_t) __popc(bal));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (built) {
        const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
        s_doc[pos] = (uint32_t) (d0 + i);
        s_tr[pos] = tk[i];
        s_q[pos] = q[i];
      }
    }
  }
  if (changed) *reinterpret_cast<ushort4 *>(node + d0) = make_ushort4((unsigned short) v[0], (unsigned short) v[1],
                                                                      (unsigned short) v[2], (unsigned short) v[3]);
  __syncthreads();
  kstamp(ktrace, blockIdx.x, 2);
  const uint32_t total = s_total;
  if (total == 0u) return;
  if (ntasks == 1u) {
    if (tid == 0) s_cnt[0] = total;
  } else {
    // rank of every staged document within its task (warp-aggregated shared atomics)
    for (uint32_t i0 = 0; i0 < total; i0 += kRouteThreads) {
      const uint32_t i = i0 + tid;
      const bool in = i < total;
      const uint32_t key = in ? s_tr[i] : 0xffffffffu;
      const uint32_t peers = __match_any_sync(0xffffffffu, key);
      const int leader = __ffs(peers) - 1;
      uint32_t b = 0;
      if (in && (int) lane == leader) b = atomicAdd(&s_cnt[key], (uint32_t) __popc(peers));
      b = __shfl_sync(0xffffffffu, b, leader);
      if (in) s_tr[i] = key | ((b + __popc(peers & ((1u << lane) - 1u))) << 10);
    }
  }
  __syncthreads();
  for (uint32_t j = tid; j < ntasks; j += kRouteThreads) s_base[j] = s_cnt[j] ? atomicAdd(counts + j, s_cnt[j]) : 0u;
  __syncthreads();
  kstamp(ktrace, blockIdx.x, 3);
  for (uint32_t i = tid; i < total; i += kRouteThreads) {
    const uint32_t tr = ntasks == 1u ? (i << 10) : s_tr[i];
    const uint32_t j = tr & 1023u, r = tr >> 10;
    const size_t pos = (size_t) s_t[j].region0 + s_base[j] + r;
    cids[pos] = s_doc[i];
    clamq[pos] = s_q[i];
  }
  kstamp(ktrace, blockIdx.x, 4);
}

// ------------------------------------------------------------------------------------------
// FAST histograms: RTNodeHistogram::update / RTNodeHistogram(parent, sampleids, ...) scatter loops
// (rtnode_histogram.cc:51-58, 183-191) in 64-bit fixed point.  Shared memory has no native 64-bit
// add, so each cell is two 32-bit limbs updated with native shared atomics: the low limb's atomic
// returns the old value, which tells this very addition whether it carried into the high limb.
// One block per (document slice, panel).  Integer sums are order-independent: the result is
// deterministic and identical for any slicing, any list order and any number of GPUs.
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

// a histogram cell as a double: FAST cells are fixed-point integers, REFERENCE cells are the doubles themselves
__device__ __forceinline__ double cell_value(bool exact, unsigned long long raw, double inv) {
  return exact ? __longlong_as_double((long long) raw) : (double) (long long) raw * inv;
}

template <bool EXACT>
__device__ __forceinline__ void scan_consider(uint32_t lc, uint32_t cn, unsigned long long sraw, double s, double inv,
                                              uint32_t minls, uint32_t k, double &best, uint32_t &bt, uint32_t &bl) {
  const uint32_t rc = cn - lc;
  if (lc >= minls && rc >= minls) {
    const double ls = cell_value(EXACT, sraw, inv);
    const double rs = s - ls;
    const double score = ls * ls / (double) lc + rs * rs / (double) rc;
    if (score > best) { best = score; bt = k; bl = lc; }   // strict '>' in ascending t, start value -1 (rt.cc:272-291)
  }
}

// Cheap estimate of how good a split is, used to select the few cells whose exact FP64 score (two divisions,
// rt.cc:283-286) is then evaluated.  score = ls^2/lc + rs^2/rc = tot^2/n + G/n with G = (ls*n - tot*lc)^2 / (lc*rc):
// the first term is the same for every threshold of a node and usually dwarfs the second, so the estimate is of G
// (from the raw fixed-point sums: the power-of-two scale does not change the order), good to ~1e-6 relative even
// when the gain is a tiny fraction of the score.  Cells failing the leaf support give -1; an empty side gives
// +inf, which makes the caller evaluate every cell exactly (only possible with a minimum leaf support of 0).
__device__ __forceinline__ double approx_gain(uint32_t lc, uint32_t cn, unsigned long long sraw, double tot, double cnd,
                                              uint32_t minls) {
  const uint32_t rc = cn - lc;
  const double d = (double) (long long) sraw * cnd - tot * (double) lc;
  const double g = d * d * (double) __frcp_rn((float) lc * (float) rc);
  if (!(lc >= minls && rc >= minls)) return -1.0;
  if (lc == 0u || rc == 0u) return __longlong_as_double(0x7ff0000000000000ll);
  return g;
}

// What the split scan hands to the polling host thread on one GPU (mapped pinned memory): per (task, child) the
// best split over all features and the node statistics, as three 16-byte records, each written with ONE
// 16-byte store and carrying the round id in its last word: no system-scope fence and no separate flag (a fence
// per publishing block cost 2-7 us).  The host clears the tag of every record it consumes, so a record whose
// tag matches was written in this round.
struct __align__(16) PubRec { double v; uint32_t a; uint32_t tag; };
struct __align__(16) ChildOut {
  PubRec split;     // v = best score (-1: none), a = left count at the best split
  PubRec where;     // v = node sum, a = feature << 16 | threshold index
  PubRec stats;     // v = squares sum of the BUILT child of the task, a = node size
};
struct __align__(16) DevCand { double score; uint32_t lc; uint32_t t; };     // per-feature winners (device memory)
__device__ __forceinline__ void publish16(PubRec *dst, double v, uint32_t a, uint32_t tag) {
  const unsigned long long bits = (unsigned long long) __double_as_longlong(v);
  *reinterpret_cast<uint4 *>(dst) = make_uint4((uint32_t) bits, (uint32_t) (bits >> 32), a, tag);
}
struct ScanOut {
  ChildOut *out;                    // [task][child] mapped host memory, child 0 = left
  DevCand *cand;                    // [task][child][feature] device scratch
  double2 *node;                    // [task][child] (sum, n) read from feature 0 (rtnode.h:99-104), device scratch
  double *sq_built;                 // [task] device scratch
  uint32_t *done;                   // [task] features scanned so far (device counter, zero between rounds)
  ulonglong2 *sq_acc;               // [task] 128-bit accumulator of the exact squares (device)
  const uint32_t *root_cnt;         // per-bin document counts of the whole dataset (root refresh)
  const int *qexp;
  uint32_t round_id, minls;
  unsigned long long *ktrace;
};

// the exact split score of rt.cc:283-286
__device__ __forceinline__ double exact_score(uint32_t lc, uint32_t rc, unsigned long long sraw, double s, double inv) {
  const double ls = (double) (long long) sraw * inv;
  const double rs = s - ls;
  return ls * ls / (double) lc + rs * rs / (double) rc;
}

// One block per SM (kHistThreads threads, one 16-feature limb histogram): every resident histogram is
// flushed with global atomics at the end of its block, so fewer, fatter blocks cut that cost; the
// update loop itself is bound by the shared-atomic issue rate, not by occupancy (scripts/hist_mb2.cu).
constexpr uint32_t kHistThreads = 512;

// whole == 1: the node is the whole (local) dataset in document order, pseudo-responses read from lamq;
// else: the task's compact list (cids / clamq from region0, `counts[task]` entries) written by route_kernel.
// ACC (one GPU): the exact squares go to one 128-bit accumulator per task instead of per-slice partials
template <typename BinT, bool SMEM, bool COUNT, bool ACC>
__global__ void __launch_bounds__(kHistThreads, 2)
hist_limb_kernel(const NodeTask *__restrict__ tasks, uint32_t ntasks, const __grid_constant__ TaskPack pack,
                 const uint32_t *__restrict__ counts, const uint4 *__restrict__ panels, size_t N,
                 const uint32_t *__restrict__ cids, const long long *__restrict__ clamq,
                 const long long *__restrict__ lamq, const uint32_t *__restrict__ thr_off, uint32_t F,
                 unsigned long long *hsum, uint32_t *hcnt, uint32_t ncells, ulonglong2 *sq_partials,
                 uint32_t stride, ulonglong2 *sq_acc, unsigned long long *ktrace, unsigned long long *kspan,
                 const uint4 *__restrict__ rows, uint32_t npanels) {
  if (pack.n) tasks = pack.t;
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  extern __shared__ __align__(1024) unsigned char hist_smem[];
  __shared__ uint32_t s_base[FPP + 1];
  __shared__ uint32_t s_task;
  __shared__ U128 s_sq[kHistThreads / 32];
  const uint32_t kblock = blockIdx.y * gridDim.x + blockIdx.x;
  kstamp(ktrace, kblock, 0);
  // profiling (bench.py roofline): first block start .. last block end of this launch on the device's own
  // clock; CUDA events around a launch on an idle stream also time the launch itself
  if (kspan != nullptr && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(kspan, t0);
  }
  if (threadIdx.x == 0) s_task = find_task_by(tasks, ntasks, blockIdx.x, true);
  __syncthreads();
  const uint32_t task = s_task;
  const NodeTask &t = tasks[task];
  const bool identity = t.whole != 0u;
  const uint32_t dpb = t.hist_dpb;
  const uint32_t begin = (blockIdx.x - t.hist_blk0) * dpb;
  const uint32_t p = blockIdx.y;
  const uint32_t f0 = p * FPP;
  const uint32_t nf = min(FPP, F - f0);
  const uint32_t cell0 = thr_off[f0];
  const uint32_t scells = SMEM ? FPP * stride : 0u;   // shared-memory cells, bin-major
  if (threadIdx.x <= FPP) s_base[threadIdx.x] = thr_off[f0 + min(threadIdx.x, nf)] - cell0;
  if (SMEM) {
    uint4 *z = reinterpret_cast<uint4 *>(hist_smem);
    const uint32_t nz = scells * (COUNT ? 3u : 2u) / 4u;   // scells is a multiple of 8
    for (uint32_t i = threadIdx.x; i < nz; i += kHistThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  pdl_wait();   // everything below reads what the route kernel wrote
  const uint32_t seglen = identity ? t.n : counts[task];
  const bool empty = begin >= seglen;
  if (empty) {
    if (!ACC && p == 0 && threadIdx.x == 0) sq_partials[blockIdx.x] = make_ulonglong2(0ull, 0ull);
    if (kspan != nullptr && threadIdx.x == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicMax(kspan + 1, t1);
    }
    return;
  }
  const uint32_t end = min(seglen, begin + dpb);
  __syncthreads();

  kstamp(ktrace, kblock, 1);
  unsigned long long *gs = hsum + (size_t) build_slot(t) * ncells + cell0;
  uint32_t *gc = hcnt + (size_t) build_slot(t) * ncells + cell0;
  const uint32_t *ids = cids + t.region0;
  const long long *lq = identity ? lamq : clamq + t.region0;
  // a gathered list reads the document-major copy when there is one (document d's row of panel p at d * npanels + p)
  const bool by_doc = !identity && rows != nullptr;
  const uint4 *prow = by_doc ? rows + p : panels + (size_t) p * N;
  const uint32_t rstride = by_doc ? npanels : 1u;
  U128 sq{0ull, 0ull};
  if (SMEM) {
    const uint32_t rot = lane_id() & (FPP - 1);
    const uint32_t sel = xor_permute_selector<BinT>(rot);
    unsigned char *rbp = hist_smem + rot * 4u;
    const uint32_t hi_off = scells * 4u, cnt_off = scells * 8u;
    uint32_t i = begin + threadIdx.x;
    uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = c0;
    long long q0 = 0, q1 = 0;
    bool v0 = i < end, v1 = i + kHistThreads < end;
    if (v0) { c0 = prow[(size_t) (identity ? i : ids[i]) * rstride]; q0 = lq[i]; }
    if (v1) { c1 = prow[(size_t) (identity ? i + kHistThreads : ids[i + kHistThreads]) * rstride]; q1 = lq[i + kHistThreads]; }
    bool w0 = i + 2 * kHistThreads < end, w1 = i + 3 * kHistThreads < end;
    uint32_t nd0 = 0, nd1 = 0;   // documents of the NEXT iteration
    if (w0) nd0 = identity ? i + 2 * kHistThreads : ids[i + 2 * kHistThreads];
    if (w1) nd1 = identity ? i + 3 * kHistThreads : ids[i + 3 * kHistThreads];
    while (v0) {
      uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = n0;
      long long nq0 = 0, nq1 = 0;
      if (w0) { n0 = prow[(size_t) nd0 * rstride]; nq0 = lq[i + 2 * kHistThreads]; }
      if (w1) { n1 = prow[(size_t) nd1 * rstride]; nq1 = lq[i + 3 * kHistThreads]; }
      i += 2 * kHistThreads;
      const bool z0 = i + 2 * kHistThreads < end, z1 = i + 3 * kHistThreads < end;
      if (z0) nd0 = identity ? i + 2 * kHistThreads : ids[i + 2 * kHistThreads];
      if (z1) nd1 = identity ? i + 3 * kHistThreads : ids[i + 3 * kHistThreads];
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
      const uint32_t d = identity ? i : ids[i];
      const uint4 row = rotate_bytes(prow[(size_t) d * rstride], rotb);
      const long long q = lq[i];
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
  kstamp(ktrace, kblock, 2);
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
    if (ACC) {
      // 128-bit add as two 64-bit atomics: the low word's returning atomic tells this addition its carry
      if (tot.lo | tot.hi) {
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(sq_acc + task);
        const unsigned long long old = atomicAdd(acc, tot.lo);
        atomicAdd(acc + 1, tot.hi + ((old + tot.lo) < old ? 1ull : 0ull));
      }
    } else {
      sq_partials[blockIdx.x] = make_ulonglong2(tot.lo, tot.hi);
    }
  }
  if (SMEM) {
    // flush: consecutive threads read consecutive shared cells (cell = bin * FPP + slot)
    const uint32_t *s_lo = reinterpret_cast<const uint32_t *>(hist_smem);
    const uint32_t *s_hi = s_lo + scells, *s_cnt = s_hi + scells;
    for (uint32_t i = threadIdx.x; i < scells; i += kHistThreads) {
      const uint32_t slot = i & (FPP - 1), bin = i / FPP;
      const long long v = ((long long) (int32_t) s_hi[i] << 32) + (long long) s_lo[i];
      // padding slots of the last panel collect the zero bins of their all-zero columns: dropped
      if (slot < nf && bin < s_base[slot + 1] - s_base[slot]) {
        if (v != 0) atomicAdd(gs + s_base[slot] + bin, (unsigned long long) v);
      }
    }
    if (COUNT) {
      // counts: two neighbouring 32-bit cells per 64-bit reduction (an SM issues ~0.8 reductions per clock:
      // the flush, not the update loop, is what a small slice costs).  Pairs are formed on the absolute index
      // into the count pool, whose base is 8-byte aligned; a count never carries out of its 32 bits.
      const size_t abs0 = (size_t) build_slot(t) * ncells + cell0;
      const uint32_t pcells = s_base[nf];
      const size_t q0 = abs0 >> 1, q1 = (abs0 + pcells + 1) >> 1;
      for (size_t q = q0 + threadIdx.x; q < q1; q += kHistThreads) {
        unsigned long long pair = 0ull;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const size_t a = 2 * q + h;
          if (a >= abs0 && a < abs0 + pcells) {
            const uint32_t rel = (uint32_t) (a - abs0);
            uint32_t slot = 0;
#pragma unroll
            for (uint32_t u = 1; u < FPP; ++u) slot += (u < nf && s_base[u] <= rel) ? 1u : 0u;
            const uint32_t bin = rel - s_base[slot];
            pair |= (unsigned long long) s_cnt[bin * FPP + slot] << (32 * h);
          }
        }
        if (pair) atomicAdd(reinterpret_cast<unsigned long long *>(hcnt) + q, pair);
      }
    }
  }
  kstamp(ktrace, kblock, 3);
  if (kspan != nullptr) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicMax(kspan + 1, t1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Split scan, one GPU: one block per (feature, task), one bin per thread, launched with programmatic
// stream serialization right behind the histogram kernel (its blocks are resident and have fetched the
// parent's bins when the histogram kernel ends).  From the RAW bins the round's histogram blocks
// accumulated in the task's raw slot: inclusive prefix over bins (rtnode_histogram.cc:59-62) -> slotB,
// sibling = parent - built (:79-85, 209-216) -> slotD, best threshold of both children (rt.cc:257-292),
// published straight to the host, which reduces over features (ScanOut).  The raw cells are cleared as they
// are consumed: the raw slot is ready for the next round without a clearing pass.
//
// The exact FP64 score (two divisions per cell and child; FP64 is slow on this part) is evaluated only for
// the cells approx_gain() cannot rule out: every cell that can be the maximum of the exact formula, or tie
// with it, is among them; candidates are compared in ascending threshold order with a strict '>'
// (rt.cc:272-291), ties across threads go to the smaller threshold.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPubThreads = 128;    // 4 warps; kPubCPT consecutive bins per thread: one pass covers 384 thresholds.
constexpr uint32_t kPubWarps = kPubThreads / 32;   // Small blocks: 8+ are resident per SM, so a round of up to ~8 node
constexpr int kPubCPT = 3;                         // expansions (136 features each) is ONE wave of blocks.

// PEER (sharded training, small rounds): every rank accumulated the built child's LOCAL histogram in a staging
// slot; after its own histogram kernel has completed, the first block tells the peers "my staging slots of this
// round are complete", every block waits for the same word from all peers, and the loads add the W staging slots
// (NVLink loads from the peers' pools).  Integer sums: every rank obtains the same totals and takes the same
// decision.  Nothing is written to a peer and the staging slots alternate between two sets by round, so one flag
// barrier per round is enough.
template <bool COUNT, bool PEER>
__global__ void __launch_bounds__(kPubThreads, PEER ? 4 : 6)
scan_pub_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ TaskPack pack, unsigned long long *hsum,
                uint32_t *hcnt, uint32_t ncells, const uint32_t *__restrict__ thr_off, uint32_t F,
                const __grid_constant__ ScanOut out, const ulonglong2 *__restrict__ sq128, uint32_t *host_err,
                const __grid_constant__ PeerView pv) {
  if (pack.n) tasks = pack.t;
  // feature-sliced exchange: this rank's blocks cover its own features [f_lo, f_hi) only
  const bool sliced = PEER && pv.f_hi != 0u;
  const uint32_t f_first = sliced ? pv.f_lo : 0u, nf_scan = sliced ? pv.f_hi - pv.f_lo : F;
  const uint32_t f = f_first + blockIdx.x, task = blockIdx.y;
  const NodeTask &t = tasks[task];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t kb = blockIdx.y * gridDim.x + blockIdx.x;
  kstamp(out.ktrace, kb, 0);
  __shared__ unsigned long long s_ws[kPubWarps];
  __shared__ uint32_t s_wc[kPubWarps];
  __shared__ double s_g[2][kPubWarps], s_bs[2][kPubWarps];
  __shared__ uint32_t s_bt[2][kPubWarps], s_bl[2][kPubWarps];

  const bool two = t.whole == 0u;
  const uint32_t minls = out.minls;
  const uint32_t c0 = thr_off[f], cells = thr_off[f + 1] - c0;
  unsigned long long *Bs = hsum + (size_t) t.slotB * ncells + c0;
  uint32_t *Bc = hcnt + (size_t) t.slotB * ncells + c0;
  const size_t roff = (size_t) build_slot(t) * ncells + c0;
  unsigned long long *Rs = hsum + roff;
  uint32_t *Rc = hcnt + roff;
  const uint32_t *Cc = COUNT ? Rc : out.root_cnt + c0;    // root refresh: the counts never change (rtnode_histogram.cc:149)
  const unsigned long long *Ps = two ? hsum + (size_t) t.slotP * ncells + c0 : Bs;
  const uint32_t *Pc = two ? hcnt + (size_t) t.slotP * ncells + c0 : Bc;
  unsigned long long *Ds = two ? hsum + (size_t) t.slotD * ncells + c0 : Bs;
  uint32_t *Dc = two ? hcnt + (size_t) t.slotD * ncells + c0 : Bc;
  constexpr uint32_t kChunk = kPubThreads * kPubCPT;
  const uint32_t nchunks = (cells + kChunk - 1) / kChunk;
  const uint32_t k0 = tid * kPubCPT;                       // this thread's first bin within a chunk

  // what does not depend on the histogram kernel is fetched before waiting for it
  unsigned long long s[kPubCPT], ps[kPubCPT];
  uint32_t c[kPubCPT], pc[kPubCPT];
#pragma unroll
  for (int i = 0; i < kPubCPT; ++i) {
    const bool in = k0 + i < cells;
    ps[i] = (in && two) ? Ps[k0 + i] : 0ull;
    pc[i] = (in && two) ? Pc[k0 + i] : 0u;
  }
  const unsigned long long plast = two ? Ps[cells - 1] : 0ull;
  const uint32_t pclast = two ? Pc[cells - 1] : 0u;
  const double inv = ldexp(1.0, -*out.qexp);
  pdl_wait();
  kstamp(out.ktrace, kb, 1);
  const int W = (PEER && pv.world > 1 && t.stage1) ? pv.world : 1;   // PEER = false: compiled out
  const bool own_raw = out.sq_acc != nullptr;   // one GPU: the raw slot is this kernel's to clear
  if (W > 1) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid < (uint32_t) W && tid != (uint32_t) pv.rank)
      st_flag(pv.peer_flags[tid] + pv.rank, pv.epoch);
    if (tid < (uint32_t) W && tid != (uint32_t) pv.rank) wait_flag_or_report(pv.flags + tid, pv.epoch, host_err);
    __syncthreads();
    kstamp(out.ktrace, kb, 4);
  }
  // the bins were accumulated with atomics by blocks on other SMs (or GPUs): read them from L2
  if (W > 1) {
    // The W copies of the chunk are read COALESCED — lane l of load i takes cell i * kPubThreads + tid, so a warp's
    // load is 256 contiguous bytes of a peer's pool: an NVLink read moves whole 32-byte sectors and system-scope
    // loads are not cached, so the scan's own layout (three consecutive cells per thread: a stride of 24 bytes
    // between lanes) asked every sector three times over and was bound by the number of remote requests in flight
    // per SM (24-28 us at 8 GPUs) — added up in that layout and handed to the scan's layout through shared memory.
    // Every peer's loads are in flight together: left to itself ptxas adds each peer's values as they arrive to
    // save registers, which serialises the NVLink round trips; the empty asm statements take every loaded value as
    // an operand, so all loads are issued before the first of them is waited for.
    __shared__ unsigned long long x_s[kPubThreads * kPubCPT];
    __shared__ uint32_t x_c[kPubThreads * kPubCPT];
    unsigned long long ts[kPubCPT];
    uint32_t tc[kPubCPT];
    unsigned long long rs[kMaxPeers][kPubCPT];
    uint32_t rc[kMaxPeers][kPubCPT];
    const bool wc = COUNT && pv.with_counts;
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      const uint32_t j = (uint32_t) i * kPubThreads + tid;
      ts[i] = j < cells ? __ldcg(Rs + j) : 0ull;
      tc[i] = j < cells ? __ldcg(Cc + j) : 0u;
    }
#pragma unroll
    for (int pr = 0; pr < kMaxPeers; ++pr) {
      const bool on = pr < W && pr != pv.rank;
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i) {
        const uint32_t j = (uint32_t) i * kPubThreads + tid;
        rs[pr][i] = 0ull; rc[pr][i] = 0u;
        if (on && j < cells) {
          asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(rs[pr][i]) : "l"(pv.sum[pr] + roff + j) : "memory");
          if (wc) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(rc[pr][i]) : "l"(pv.cnt[pr] + roff + j) : "memory");
        }
      }
    }
    static_assert(kPubCPT == 3, "the operand lists below name three cells per thread");
#pragma unroll
    for (int pr = 0; pr < kMaxPeers; ++pr)
      asm volatile("" : "+l"(rs[pr][0]), "+l"(rs[pr][1]), "+l"(rs[pr][2]), "+r"(rc[pr][0]), "+r"(rc[pr][1]), "+r"(rc[pr][2]));
#pragma unroll
    for (int pr = 0; pr < kMaxPeers; ++pr)
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i) { ts[i] += rs[pr][i]; tc[i] += rc[pr][i]; }
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      const uint32_t j = (uint32_t) i * kPubThreads + tid;
      x_s[j] = ts[i]; x_c[j] = tc[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) { s[i] = x_s[k0 + i]; c[i] = x_c[k0 + i]; }   // (cells beyond the feature's are 0)
    kstamp(out.ktrace, kb, 5);
  } else {
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      const bool in = k0 + i < cells;
      s[i] = in ? __ldcg(Rs + k0 + i) : 0ull;
      c[i] = in ? __ldcg(Cc + k0 + i) : 0u;
    }
  }
  if (f == f_first && warp == 0) {
    // exact squares of the built child
    U128 tot{0ull, 0ull};
    if (own_raw) {   // one GPU: every p == 0 histogram slice has added its part to the accumulator
      if (lane == 0) {
        unsigned long long *acc = reinterpret_cast<unsigned long long *>(out.sq_acc + task);
        tot.lo = __ldcg(acc); tot.hi = __ldcg(acc + 1);
        acc[0] = 0ull; acc[1] = 0ull;
      }
    } else {         // sharded: the slices' partials of every rank (or the all-reduced total in the first slice)
      const uint32_t items = (uint32_t) W * t.hist_nblk;
      for (uint32_t it = lane; it < items; it += 32) {
        const uint32_t pr = it / t.hist_nblk, i = it - pr * t.hist_nblk;
        const ulonglong2 *src = (W > 1 ? pv.sq[pr] : sq128) + t.hist_blk0 + i;
        const volatile unsigned long long *v = reinterpret_cast<const volatile unsigned long long *>(src);
        const unsigned long long lo = v[0], hi = v[1];
        u128_add(tot, lo, hi);
      }
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ol = __shfl_xor_sync(0xffffffffu, tot.lo, o);
        const unsigned long long oh = __shfl_xor_sync(0xffffffffu, tot.hi, o);
        u128_add(tot, ol, oh);
      }
    }
    if (lane == 0) {
      const double inv2 = ldexp(1.0, -2 * *out.qexp);
      out.sq_built[task] = ((double) tot.hi * 18446744073709551616.0 + (double) tot.lo) * inv2;
    }
  }

  // phase 1: inclusive prefix over the bins of the built child -> slotB; the raw cells are cleared
  unsigned long long carry_s = 0ull;
  uint32_t carry_c = 0u;
  for (uint32_t ch = 0; ch < nchunks; ++ch) {
    const uint32_t kk = ch * kChunk + k0;
    if (ch > 0) {
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i) {
        const bool in = kk + i < cells;
        s[i] = in ? __ldcg(Rs + kk + i) : 0ull;
        c[i] = in ? __ldcg(Cc + kk + i) : 0u;
        if (W > 1 && in) {
          for (int pr = 0; pr < W; ++pr) {
            if (pr == pv.rank) continue;
            s[i] += *reinterpret_cast<const volatile unsigned long long *>(pv.sum[pr] + roff + kk + i);
            if (COUNT && pv.with_counts) c[i] += *reinterpret_cast<const volatile uint32_t *>(pv.cnt[pr] + roff + kk + i);
          }
        }
      }
    }
    if (own_raw) {
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i)
        if (kk + i < cells) { Rs[kk + i] = 0ull; if (COUNT) Rc[kk + i] = 0u; }
    }
#pragma unroll
    for (int i = 1; i < kPubCPT; ++i) { s[i] += s[i - 1]; c[i] += c[i - 1]; }
    unsigned long long run_s = s[kPubCPT - 1];
    uint32_t run_c = c[kPubCPT - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long us = __shfl_up_sync(0xffffffffu, run_s, o);
      const uint32_t uc = __shfl_up_sync(0xffffffffu, run_c, o);
      if ((int) lane >= o) { run_s += us; run_c += uc; }
    }
    if (lane == 31) { s_ws[warp] = run_s; s_wc[warp] = run_c; }
    __syncthreads();
    unsigned long long add_s = carry_s + run_s - s[kPubCPT - 1], tot_s = 0ull;   // exclusive offset of this thread's run
    uint32_t add_c = carry_c + run_c - c[kPubCPT - 1], tot_c = 0u;
#pragma unroll
    for (uint32_t w = 0; w < kPubWarps; ++w) {
      const unsigned long long ws = s_ws[w];
      const uint32_t wc = s_wc[w];
      if (w < warp) { add_s += ws; add_c += wc; }
      tot_s += ws; tot_c += wc;
    }
    carry_s += tot_s; carry_c += tot_c;
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      s[i] += add_s; c[i] += add_c;
      if (kk + i < cells) { Bs[kk + i] = s[i]; Bc[kk + i] = c[i]; }
    }
    __syncthreads();
  }
  // node totals = the cumulative value of the last bin (rtnode.h:99-104)
  const unsigned long long tot_raw0 = carry_s, tot_raw1 = plast - carry_s;
  const uint32_t tot_cn0 = carry_c, tot_cn1 = pclast - carry_c;
  const double ts0 = (double) (long long) tot_raw0 * inv, ts1 = (double) (long long) tot_raw1 * inv;
  const double tr0 = (double) (long long) tot_raw0, tr1 = (double) (long long) tot_raw1;
  const double cnd0 = (double) tot_cn0, cnd1 = (double) tot_cn1;

  // phase 2: sibling = parent - built -> slotD; gain estimate of every threshold of both children
  double g0[kPubCPT], g1[kPubCPT], m0 = -1.0, m1 = -1.0;
  for (uint32_t ch = 0; ch < nchunks; ++ch) {
    const uint32_t kk = ch * kChunk + k0;
    if (nchunks > 1) {   // (a single chunk is still in registers)
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i) {
        const bool in = kk + i < cells;
        s[i] = in ? Bs[kk + i] : 0ull; c[i] = in ? Bc[kk + i] : 0u;
        ps[i] = (in && two) ? Ps[kk + i] : 0ull; pc[i] = (in && two) ? Pc[kk + i] : 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      g0[i] = g1[i] = -1.0;
      if (kk + i < cells) {
        g0[i] = approx_gain(c[i], tot_cn0, s[i], tr0, cnd0, minls);
        m0 = fmax(m0, g0[i]);
        if (two) {
          ps[i] -= s[i]; pc[i] -= c[i];                      // the sibling's cumulative bins
          Ds[kk + i] = ps[i]; Dc[kk + i] = pc[i];
          g1[i] = approx_gain(pc[i], tot_cn1, ps[i], tr1, cnd1, minls);
          m1 = fmax(m1, g1[i]);
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
  }
  if (lane == 0) { s_g[0][warp] = m0; s_g[1][warp] = m1; }
  __syncthreads();
#pragma unroll
  for (uint32_t w = 0; w < kPubWarps; ++w) { m0 = fmax(m0, s_g[0][w]); m1 = fmax(m1, s_g[1][w]); }
  // a cell can hold the maximum of the exact formula only if its G is within the estimate's error (relative) and
  // the formula's own FP64 rounding noise (~4e-16 of tot^2 + G, in G's units) of the best G
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double th0 = m0 < inf ? m0 - 1e-5 * m0 - 1e-14 * tr0 * tr0 : -inf;
  const double th1 = m1 < inf ? m1 - 1e-5 * m1 - 1e-14 * tr1 * tr1 : -inf;

  // phase 3: exact score of the candidates, ascending threshold order
  double best0 = -1.0, best1 = -1.0;
  uint32_t bt0 = 0xffffffffu, bt1 = 0xffffffffu, bl0 = 0u, bl1 = 0u;
  for (uint32_t ch = 0; ch < nchunks; ++ch) {
    const uint32_t kk = ch * kChunk + k0;
    if (nchunks > 1) {
#pragma unroll
      for (int i = 0; i < kPubCPT; ++i) {
        g0[i] = g1[i] = -1.0;
        if (kk + i < cells) {
          s[i] = Bs[kk + i]; c[i] = Bc[kk + i];
          g0[i] = approx_gain(c[i], tot_cn0, s[i], tr0, cnd0, minls);
          if (two) { ps[i] = Ds[kk + i]; pc[i] = Dc[kk + i]; g1[i] = approx_gain(pc[i], tot_cn1, ps[i], tr1, cnd1, minls); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kPubCPT; ++i) {
      if (g0[i] >= 0.0 && g0[i] >= th0) {
        const double sc = exact_score(c[i], tot_cn0 - c[i], s[i], ts0, inv);
        if (sc > best0) { best0 = sc; bt0 = kk + i; bl0 = c[i]; }
      }
      if (two && g1[i] >= 0.0 && g1[i] >= th1) {
        const double sc = exact_score(pc[i], tot_cn1 - pc[i], ps[i], ts1, inv);
        if (sc > best1) { best1 = sc; bt1 = kk + i; bl1 = pc[i]; }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {   // arg-max, ties to the smaller t
    const double ob0 = __shfl_xor_sync(0xffffffffu, best0, o), ob1 = __shfl_xor_sync(0xffffffffu, best1, o);
    const uint32_t ot0 = __shfl_xor_sync(0xffffffffu, bt0, o), ot1 = __shfl_xor_sync(0xffffffffu, bt1, o);
    const uint32_t ol0 = __shfl_xor_sync(0xffffffffu, bl0, o), ol1 = __shfl_xor_sync(0xffffffffu, bl1, o);
    if (ob0 > best0 || (ob0 == best0 && ot0 < bt0)) { best0 = ob0; bt0 = ot0; bl0 = ol0; }
    if (ob1 > best1 || (ob1 == best1 && ot1 < bt1)) { best1 = ob1; bt1 = ot1; bl1 = ol1; }
  }
  if (lane == 0) {
    s_bs[0][warp] = best0; s_bt[0][warp] = bt0; s_bl[0][warp] = bl0;
    s_bs[1][warp] = best1; s_bt[1][warp] = bt1; s_bl[1][warp] = bl1;
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (uint32_t w = 1; w < kPubWarps; ++w) {
      if (s_bs[0][w] > best0 || (s_bs[0][w] == best0 && s_bt[0][w] < bt0)) { best0 = s_bs[0][w]; bt0 = s_bt[0][w]; bl0 = s_bl[0][w]; }
      if (s_bs[1][w] > best1 || (s_bs[1][w] == best1 && s_bt[1][w] < bt1)) { best1 = s_bs[1][w]; bt1 = s_bt[1][w]; bl1 = s_bl[1][w]; }
    }
    // pass 0 scanned the built child, pass 1 the derived one; child 0 = left
    const int child0 = two ? (t.build_left != 0u ? 0 : 1) : 0;
    DevCand r;
    r.score = best0; r.lc = bl0; r.t = bt0;
    out.cand[((size_t) task * 2 + child0) * F + f] = r;
    if (two) {
      r.score = best1; r.lc = bl1; r.t = bt1;
      out.cand[((size_t) task * 2 + (1 - child0)) * F + f] = r;
    }
    if (f == f_first) {   // (integer sums: every feature's last cumulative bin holds the same node totals)
      out.node[(size_t) task * 2 + child0] = make_double2(ts0, (double) tot_cn0);
      if (two) out.node[(size_t) task * 2 + (1 - child0)] = make_double2(ts1, (double) tot_cn1);
    }
    kstamp(out.ktrace, kb, 2);
    // the last feature of the task to get here takes the first maximum over features (rt.cc:297-306)
    __threadfence();
    s_bt[0][0] = (atomicAdd(out.done + task, 1u) == nf_scan - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_bt[0][0] == 0u) return;
  __threadfence();
  __syncthreads();
  const int nchild = two ? 2 : 1;
  double fb[2] = {-1.0, -1.0};
  uint32_t ff[2] = {0xffffffffu, 0xffffffffu}, ft[2] = {0xffffffffu, 0xffffffffu}, fl[2] = {0u, 0u};
  for (uint32_t g = f_first + tid; g < f_first + nf_scan; g += kPubThreads) {   // ascending g per thread: its first maximum
#pragma unroll
    for (int child = 0; child < 2; ++child) {
      if (child < nchild) {
        const DevCand *q = out.cand + ((size_t) task * 2 + child) * F + g;
        const double sc = __ldcg(&q->score);
        const uint32_t tt = __ldcg(&q->t), ll = __ldcg(&q->lc);
        if (sc > fb[child]) { fb[child] = sc; ff[child] = g; ft[child] = tt; fl[child] = ll; }
      }
    }
  }
#pragma unroll
  for (int child = 0; child < 2; ++child) {
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, fb[child], o);
      const uint32_t of = __shfl_xor_sync(0xffffffffu, ff[child], o);
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, ft[child], o);
      const uint32_t ol = __shfl_xor_sync(0xffffffffu, fl[child], o);
      if (ob > fb[child] || (ob == fb[child] && of < ff[child])) { fb[child] = ob; ff[child] = of; ft[child] = ot; fl[child] = ol; }
    }
  }
  __shared__ uint32_t s_bf[2][kPubWarps];
  if (lane == 0) {
    s_bs[0][warp] = fb[0]; s_bf[0][warp] = ff[0]; s_bt[0][warp] = ft[0]; s_bl[0][warp] = fl[0];
    s_bs[1][warp] = fb[1]; s_bf[1][warp] = ff[1]; s_bt[1][warp] = ft[1]; s_bl[1][warp] = fl[1];
  }
  __syncthreads();
  double b = -1.0;
  uint32_t f1 = 0xffffffffu, t1 = 0xffffffffu, l1 = 0u;
  if ((int) tid < nchild) {   // thread c finishes child c
    const int child = (int) tid;
    const double *sb = child == 0 ? s_bs[0] : s_bs[1];
    const uint32_t *sf = child == 0 ? s_bf[0] : s_bf[1], *st = child == 0 ? s_bt[0] : s_bt[1], *sl = child == 0 ? s_bl[0] : s_bl[1];
    b = sb[0];
    f1 = sf[0]; t1 = st[0]; l1 = sl[0];
#pragma unroll
    for (uint32_t w = 1; w < kPubWarps; ++w)
      if (sb[w] > b || (sb[w] == b && sf[w] < f1)) { b = sb[w]; f1 = sf[w]; t1 = st[w]; l1 = sl[w]; }
  }
  if (PEER && sliced) {
    // The winners over this rank's features go to every rank's mailbox (records of four tagged words, see
    // qr_task.cuh); the ranks' winners are then compared in rank order = ascending feature order with a strict '>'
    // (rt.cc:297-306), so every rank publishes the same split.
    __shared__ double s_xb[2][kMaxPeers];
    __shared__ uint32_t s_xa[2][kMaxPeers], s_xl[2][kMaxPeers];
    __syncthreads();
    if ((int) tid < nchild) { s_xb[tid][pv.rank] = b; s_xa[tid][pv.rank] = (f1 << 16) | (t1 & 0xffffu); s_xl[tid][pv.rank] = l1; }
    __syncthreads();
    const uint32_t Wn = (uint32_t) pv.world, ep = pv.epoch;
    if (tid < (uint32_t) nchild * Wn) {
      const int child = (int) (tid / Wn), p = (int) (tid % Wn);
      if (p != pv.rank) {
        const unsigned long long bits = (unsigned long long) __double_as_longlong(s_xb[child][pv.rank]);
        volatile unsigned long long *dst = pv.mail[p] + mail_index(ep, pv.rank, pv.mail_tasks, task, child);
        dst[0] = (bits << 32) | ep;
        dst[1] = (bits & 0xffffffff00000000ull) | ep;
        dst[2] = ((unsigned long long) s_xl[child][pv.rank] << 32) | ep;
        dst[3] = ((unsigned long long) s_xa[child][pv.rank] << 32) | ep;
        const volatile unsigned long long *src = pv.mail[pv.rank] + mail_index(ep, p, pv.mail_tasks, task, child);
        unsigned long long w0, w1, w2, w3;
        const long long t0 = clock64();
        for (;;) {
          w0 = src[0]; w1 = src[1]; w2 = src[2]; w3 = src[3];
          if ((uint32_t) w0 == ep && (uint32_t) w1 == ep && (uint32_t) w2 == ep && (uint32_t) w3 == ep) break;
          if (clock64() - t0 > 40000000000ll) {   // ~20 s: a peer never arrived
            if (host_err) { *reinterpret_cast<volatile uint32_t *>(host_err) = 1u; __threadfence_system(); }
            break;
          }
        }
        s_xb[child][p] = __longlong_as_double((long long) ((w0 >> 32) | (w1 & 0xffffffff00000000ull)));
        s_xl[child][p] = (uint32_t) (w2 >> 32);
        s_xa[child][p] = (uint32_t) (w3 >> 32);
      }
    }
    __syncthreads();
    kstamp(out.ktrace, kb, 6);
    if ((int) tid < nchild) {
      b = s_xb[tid][0];
      uint32_t a = s_xa[tid][0];
      l1 = s_xl[tid][0];
      for (uint32_t p = 1; p < Wn; ++p)
        if (s_xb[tid][p] > b) { b = s_xb[tid][p]; a = s_xa[tid][p]; l1 = s_xl[tid][p]; }
      f1 = a >> 16; t1 = a & 0xffffu;
    }
  }
  if ((int) tid < nchild) {
    const int child = (int) tid;
    const double2 nd = __ldcg(out.node + (size_t) task * 2 + child);
    const double sq = __ldcg(out.sq_built + task);
    ChildOut *o = out.out + (size_t) task * 2 + child;
    publish16(&o->split, b, l1, out.round_id);
    publish16(&o->where, nd.x, (f1 << 16) | (t1 & 0xffffu), out.round_id);
    publish16(&o->stats, sq, (uint32_t) nd.y, out.round_id);
  }
  if (tid == 0) {
    out.done[task] = 0u;   // ready for the next round
    kstamp(out.ktrace, kb, 3);
  }
}

constexpr uint32_t kScanThreads = 288;   // 9 warps: one pass covers a feature of up to 288 thresholds
constexpr uint32_t kScanWarps = kScanThreads / 32;

// ------------------------------------------------------------------------------------------
// Split scan: cumulative histogram of the built child, sibling = parent - built
// (rtnode_histogram.cc:59-62, 79-85, 209-216) and the score of every (feature, threshold) for
// both children (rt.cc:257-292).  One block per (feature, task), one bin per thread: the dependent
// chain is a block-wide prefix (5 shuffle steps + one shared-memory hop), four FP64 divisions and a
// block arg-max.  Per-feature winners go to cand_*[task][child][f], child 0 = left; the task's last
// block takes the first maximum over features (rt.cc:297-306), fills the node statistics of
// RTNode(sampleids, hist) (rtnode.h:97-107) and publishes them to mapped host memory.
//
// PEER (sharded training, fused exchange): every rank accumulated the built child's LOCAL histogram
// in a staging slot; this kernel is stream-ordered after that, so its first block tells the peers
// "my staging slots of this round are complete", every block waits for the same word from all
// peers, and the loads add the W staging slots (NVLink loads from the peers' pools).  Integer sums:
// every rank obtains the same totals.  Nothing is written to a peer and the staging slots alternate
// between two sets by round, so one flag barrier per round is enough.
// ------------------------------------------------------------------------------------------

template <bool PEER, bool EXACT>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const NodeTask *__restrict__ tasks, const __grid_constant__ TaskPack pack, unsigned long long *hsum,
            uint32_t *hcnt, uint32_t ncells, const uint32_t *__restrict__ thr_off, uint32_t F, uint32_t minls,
            const int *__restrict__ qexp, double *cand_score, uint32_t *cand_t, uint32_t *cand_lc,
            ulonglong2 *totals, double *sq_built, const ulonglong2 *__restrict__ sq128,
            const double *__restrict__ sq_exact, uint32_t *task_done,
            SplitResult *res, volatile uint32_t *host_flags, uint32_t round_id, uint32_t *host_err,
            const __grid_constant__ PeerView pv) {
  if (pack.n) tasks = pack.t;
  const uint32_t f = blockIdx.x, task = blockIdx.y;
  const NodeTask &t = tasks[task];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int W = (PEER && !EXACT && pv.world > 1 && t.stage1) ? pv.world : 1;   // PEER = false: compiled out
  if (W > 1) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid < (uint32_t) W && tid != (uint32_t) pv.rank)
      st_flag(pv.peer_flags[tid] + pv.rank, pv.epoch);
    if (tid < (uint32_t) W && tid != (uint32_t) pv.rank) wait_flag_or_report(pv.flags + tid, pv.epoch, host_err);
    __syncthreads();
  }
  __shared__ unsigned long long s_ws[kScanWarps];
  __shared__ uint32_t s_wc[kScanWarps];
  __shared__ double s_bs[2][kScanWarps];
  __shared__ uint32_t s_bt[2][kScanWarps], s_bl[2][kScanWarps], s_bf[2][kScanWarps];
  __shared__ uint32_t s_last;

  const double inv = EXACT ? 1.0 : ldexp(1.0, -*qexp);
  const bool two = t.whole == 0u;
  const int nchild = two ? 2 : 1;
  const uint32_t c0 = thr_off[f], cells = thr_off[f + 1] - c0;
  unsigned long long *Bs = hsum + (size_t) t.slotB * ncells + c0;
  uint32_t *Bc = hcnt + (size_t) t.slotB * ncells + c0;
  const size_t roff = (size_t) build_slot(t) * ncells + c0;   // raw bins: slotB itself, or the staging slot(s)
  const unsigned long long *Rs = hsum + roff;
  const uint32_t *Rc = hcnt + roff;
  const unsigned long long *Ps = two ? hsum + (size_t) t.slotP * ncells + c0 : Bs;
  const uint32_t *Pc = two ? hcnt + (size_t) t.slotP * ncells + c0 : Bc;
  unsigned long long *Ds = two ? hsum + (size_t) t.slotD * ncells + c0 : Bs;
  uint32_t *Dc = two ? hcnt + (size_t) t.slotD * ncells + c0 : Bc;
  const uint32_t nchunks = (cells + kScanThreads - 1) / kScanThreads;

  // every independent load is issued before anything waits on one
  const uint32_t k0 = tid;
  const bool in0 = k0 < cells;
  unsigned long long s = in0 ? Rs[k0] : 0ull;
  uint32_t c = in0 ? Rc[k0] : 0u;
  unsigned long long p_s = (in0 && two) ? Ps[k0] : 0ull;
  uint32_t p_c = (in0 && two) ? Pc[k0] : 0u;
  const unsigned long long plast = two ? Ps[cells - 1] : 0ull;
  const uint32_t pclast = two ? Pc[cells - 1] : 0u;
  // REFERENCE mode: hist_exact_kernel left the bins cumulative already (its sequential prefix is the reference's)
  const unsigned long long blast = EXACT ? Bs[cells - 1] : 0ull;
  const uint32_t bclast = EXACT ? Bc[cells - 1] : 0u;

  // exact squares of the built child: the slices' 128-bit partials (of every rank), folded by warp 0 of the
  // task's first block while the others scan
  if (EXACT) {
    if (f == 0 && tid == 0) sq_built[task] = sq_exact[t.sq0];
  } else if (f == 0 && warp == 0) {
    U128 tot{0ull, 0ull};
    const uint32_t items = (uint32_t) W * t.hist_nblk;
    for (uint32_t it = lane; it < items; it += 32) {
      const uint32_t pr = it / t.hist_nblk, i = it - pr * t.hist_nblk;
      const ulonglong2 *src = (W > 1 ? pv.sq[pr] : sq128) + t.hist_blk0 + i;
      const volatile unsigned long long *v = reinterpret_cast<const volatile unsigned long long *>(src);
      const unsigned long long lo = v[0], hi = v[1];
      u128_add(tot, lo, hi);
    }
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ol = __shfl_xor_sync(0xffffffffu, tot.lo, o);
      const unsigned long long oh = __shfl_xor_sync(0xffffffffu, tot.hi, o);
      u128_add(tot, ol, oh);
    }
    if (lane == 0) {
      const double inv2 = ldexp(1.0, -2 * *qexp);
      sq_built[task] = ((double) tot.hi * 18446744073709551616.0 + (double) tot.lo) * inv2;
    }
  }

  // phase 1: inclusive prefix over the bins of the built child, written back in place (slotB)
  unsigned long long carry_s = blast;
  uint32_t carry_c = bclast;
  for (uint32_t ch = 0; ch < (EXACT ? 0u : nchunks); ++ch) {
    const uint32_t k = ch * kScanThreads + tid;
    const bool in = k < cells;
    if (ch > 0) { s = in ? Rs[k] : 0ull; c = in ? Rc[k] : 0u; }
    if (W > 1 && in) {
      for (int pr = 0; pr < W; ++pr) {
        if (pr == pv.rank) continue;
        s += *reinterpret_cast<const volatile unsigned long long *>(pv.sum[pr] + roff + k);
        if (pv.with_counts) c += *reinterpret_cast<const volatile uint32_t *>(pv.cnt[pr] + roff + k);
      }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long us = __shfl_up_sync(0xffffffffu, s, o);
      const uint32_t uc = __shfl_up_sync(0xffffffffu, c, o);
      if ((int) lane >= o) { s += us; c += uc; }
    }
    if (lane == 31) { s_ws[warp] = s; s_wc[warp] = c; }
    __syncthreads();
    unsigned long long add_s = carry_s, tot_s = 0ull;
    uint32_t add_c = carry_c, tot_c = 0u;
#pragma unroll
    for (uint32_t w = 0; w < kScanWarps; ++w) {
      const unsigned long long ws = s_ws[w];
      const uint32_t wc = s_wc[w];
      if (w < warp) { add_s += ws; add_c += wc; }
      tot_s += ws; tot_c += wc;
    }
    s += add_s; c += add_c;
    carry_s += tot_s; carry_c += tot_c;
    if (in) { Bs[k] = s; Bc[k] = c; }
    __syncthreads();
  }
  // node totals = the cumulative value of the last bin (rtnode.h:99-104)
  const unsigned long long tot_raw[2] = {
      carry_s, EXACT ? (unsigned long long) __double_as_longlong(__longlong_as_double((long long) plast) -
                                                                 __longlong_as_double((long long) carry_s))
                     : plast - carry_s};
  const uint32_t tot_cn[2] = {carry_c, pclast - carry_c};
  const double tot_s[2] = {cell_value(EXACT, tot_raw[0], inv), cell_value(EXACT, tot_raw[1], inv)};

  // phase 2: sibling = parent - built, split score of every threshold of both children
  double best[2] = {-1.0, -1.0};
  uint32_t bt[2] = {0xffffffffu, 0xffffffffu}, bl[2] = {0u, 0u};
  for (uint32_t ch = 0; ch < nchunks; ++ch) {
    const uint32_t k = ch * kScanThreads + tid;
    const bool in = k < cells;
    if (nchunks > 1) {   // (a single chunk is still in registers)
      s = in ? Bs[k] : 0ull; c = in ? Bc[k] : 0u;
      p_s = (in && two) ? Ps[k] : 0ull; p_c = (in && two) ? Pc[k] : 0u;
    }
    if (in) {
      scan_consider<EXACT>(c, tot_cn[0], s, tot_s[0], inv, minls, k, best[0], bt[0], bl[0]);
      if (two) {
        const unsigned long long ds = EXACT ? (unsigned long long) __double_as_longlong(   // rtnode_histogram.cc:82
                                                  __longlong_as_double((long long) p_s) - __longlong_as_double((long long) s))
                                            : p_s - s;
        const uint32_t dc = p_c - c;
        Ds[k] = ds; Dc[k] = dc;
        scan_consider<EXACT>(dc, tot_cn[1], ds, tot_s[1], inv, minls, k, best[1], bt[1], bl[1]);
      }
    }
  }
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    if (pass < nchild) {
      for (int o = 16; o > 0; o >>= 1) {   // arg-max, ties to the smaller t
        const double ob = __shfl_xor_sync(0xffffffffu, best[pass], o);
        const uint32_t ot = __shfl_xor_sync(0xffffffffu, bt[pass], o);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, bl[pass], o);
        if (ob > best[pass] || (ob == best[pass] && ot < bt[pass])) { best[pass] = ob; bt[pass] = ot; bl[pass] = ol; }
      }
      if (lane == 0) { s_bs[pass][warp] = best[pass]; s_bt[pass][warp] = bt[pass]; s_bl[pass][warp] = bl[pass]; }
    }
  }
  __syncthreads();
  if ((int) tid < nchild) {
    const int pass = (int) tid;
    const double *sb = pass == 0 ? s_bs[0] : s_bs[1];
    const uint32_t *st = pass == 0 ? s_bt[0] : s_bt[1], *sl = pass == 0 ? s_bl[0] : s_bl[1];
    double b = sb[0];
    uint32_t t1 = st[0], l1 = sl[0];
    for (uint32_t w = 1; w < kScanWarps; ++w)
      if (sb[w] > b || (sb[w] == b && st[w] < t1)) { b = sb[w]; t1 = st[w]; l1 = sl[w]; }
    // pass 0 scanned the built child, pass 1 the derived one; child 0 = left
    const int child = two ? ((pass == 0) == (t.build_left != 0u) ? 0 : 1) : 0;
    const size_t o = ((size_t) task * 2 + child) * F + f;
    cand_score[o] = b; cand_t[o] = t1; cand_lc[o] = l1;
    // node size and sum are read from feature 0's last bin (rtnode.h:99-104)
    if (f == 0) totals[(size_t) task * 2 + child] = make_ulonglong2((unsigned long long) (pass == 0 ? tot_cn[0] : tot_cn[1]),
                                                                   pass == 0 ? tot_raw[0] : tot_raw[1]);
  }

  // the last block of the task to get here reduces over features
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(task_done + task, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double fb[2] = {-1.0, -1.0};
  uint32_t ff[2] = {0xffffffffu, 0xffffffffu}, ft[2] = {0xffffffffu, 0xffffffffu}, fl[2] = {0u, 0u};
  for (uint32_t g = tid; g < F; g += kScanThreads) {
#pragma unroll
    for (int child = 0; child < 2; ++child) {
      if (child < nchild) {
        const size_t o = ((size_t) task * 2 + child) * F + g;
        const double sc = __ldcg(cand_score + o);
        const uint32_t tt = __ldcg(cand_t + o), ll = __ldcg(cand_lc + o);
        if (sc > fb[child]) { fb[child] = sc; ff[child] = g; ft[child] = tt; fl[child] = ll; }   // ascending g: first maximum
      }
    }
  }
  ulonglong2 tv = make_ulonglong2(0ull, 0ull);
  double sqB = 0.0;
  if ((int) tid < nchild) {
    tv = __ldcg(totals + (size_t) task * 2 + tid);
    sqB = __ldcg(sq_built + task);
  }
#pragma unroll
  for (int child = 0; child < 2; ++child) {
    if (child < nchild) {
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, fb[child], o);
        const uint32_t of = __shfl_xor_sync(0xffffffffu, ff[child], o);
        const uint32_t ot = __shfl_xor_sync(0xffffffffu, ft[child], o);
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, fl[child], o);
        if (ob > fb[child] || (ob == fb[child] && of < ff[child])) { fb[child] = ob; ff[child] = of; ft[child] = ot; fl[child] = ol; }
      }
      if (lane == 0) { s_bs[child][warp] = fb[child]; s_bf[child][warp] = ff[child]; s_bt[child][warp] = ft[child]; s_bl[child][warp] = fl[child]; }
    }
  }
  __syncthreads();
  if ((int) tid < nchild) {   // thread c finishes child c
    const int child = (int) tid;
    double b = s_bs[child][0];
    uint32_t f1 = s_bf[child][0], t1 = s_bt[child][0], l1 = s_bl[child][0];
    for (uint32_t w = 1; w < kScanWarps; ++w)
      if (s_bs[child][w] > b || (s_bs[child][w] == b && s_bf[child][w] < f1)) { b = s_bs[child][w]; f1 = s_bf[child][w]; t1 = s_bt[child][w]; l1 = s_bl[child][w]; }
    const bool built = !two || ((child == 0) == (t.build_left != 0u));
    SplitResult r;
    r.n = tv.x;
    r.sum = cell_value(EXACT, tv.y, inv);
    r.squares = built ? sqB : t.parent_squares - sqB;          // rtnode_histogram.cc:86,207
    r.deviance = r.squares - r.sum * r.sum / (double) r.n;      // rtnode.h:106
    r.score = b;
    r.valid = b != -1.0;
    r.feature = f1;
    r.threshold_idx = r.valid ? t1 : 0xffffffffu;
    r.lcount = r.valid ? l1 : 0;
    r.pad = 0;
    res[(size_t) task * 2 + child] = r;
    if (host_flags) __threadfence_system();   // res lives in mapped host memory
  }
  __syncthreads();
  if (tid == 0) {
    task_done[task] = 0u;   // ready for the next round
    if (host_flags) host_flags[task] = round_id;   // publish to the polling host thread
  }
}

// ------------------------------------------------------------------------------------------
// Leaf outputs (RegressionTree::update_output, rt.cc:165-207): per-leaf sums of the pseudo-responses
// and of the Newton weights in ONE pass over node_of_doc (a warp takes 32 consecutive documents per step: the
// documents of each leaf present are summed by the first of them and added to the warp's per-leaf accumulator in
// shared memory; warps, then blocks are summed).  Also writes the doc -> leaf map used by the score update
// (mart.cc:459-468).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kLeafWarpDocs = 256;   // documents per warp

// The sums are exact integers — the fixed-point pseudo-responses the histograms are made of (lamq, scale qexp[0]) and
// the weights at their own scale (qexp[1]) — so a leaf's output is the same for any slicing of the documents over
// blocks and over GPUs: sharded runs grow bit-identical models.
__global__ void leaf_node_kernel(const uint16_t *__restrict__ node, const uint16_t *__restrict__ leaf_lut,
                                 uint32_t nnodes, uint32_t nl, const long long *__restrict__ lamq,
                                 const double *__restrict__ wgt, const int *__restrict__ qexp, size_t N,
                                 longlong2 *partials, uint32_t *__restrict__ leaf_of_doc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  longlong2 *tab = reinterpret_cast<longlong2 *>(smem_raw);                 // [nwarps][nl] per-warp leaf accumulators
  longlong2 *stage = tab + (size_t) nwarps * nl;                            // [nwarps][32] the step's values
  uint16_t *lut = reinterpret_cast<uint16_t *>(stage + (size_t) nwarps * 32);   // [nnodes]
  for (uint32_t i = threadIdx.x; i < nwarps * nl; i += blockDim.x) tab[i] = make_longlong2(0ll, 0ll);
  for (uint32_t i = threadIdx.x; i < nnodes; i += blockDim.x) lut[i] = leaf_lut[i];
  longlong2 *mine = tab + (size_t) warp * nl;
  longlong2 *sv = stage + (size_t) warp * 32;
  const size_t base = ((size_t) blockIdx.x * nwarps + warp) * kLeafWarpDocs;
  constexpr uint32_t kIt = kLeafWarpDocs / 32;
  const int wexp = wgt ? qexp[1] : 0;
  // every load of the warp's documents is issued up front (one memory round trip instead of one per step)
  uint32_t nid[kIt];
  long long l1[kIt];
  double l2[kIt];
#pragma unroll
  for (uint32_t it = 0; it < kIt; ++it) {
    const size_t d = base + it * 32 + lane;
    nid[it] = d < N ? node[d] : 0xffffffffu;
    l1[it] = d < N ? lamq[d] : 0ll;
    l2[it] = (d < N && wgt) ? wgt[d] : 0.0;
  }
  __syncthreads();   // (the node -> leaf table)
#pragma unroll
  for (uint32_t it = 0; it < kIt; ++it) {
    const size_t d = base + it * 32 + lane;
    uint32_t leaf = 0xffffffffu;
    if (d < N) {
      leaf = lut[nid[it]];
      leaf_of_doc[d] = leaf;
    }
    sv[lane] = make_longlong2(l1[it], __double2ll_rn(ldexp(l2[it], wexp)));
    // the documents of each leaf present in this step are summed by the first of them
    const uint32_t peers = __match_any_sync(0xffffffffu, leaf);
    __syncwarp();
    if (d < N && (peers & ((1u << lane) - 1u)) == 0u) {
      long long a = 0, b = 0;
      for (uint32_t m = peers; m; m &= m - 1u) {
        const longlong2 x = sv[__ffs(m) - 1];
        a += x.x; b += x.y;
      }
      longlong2 acc = mine[leaf];
      acc.x += a; acc.y += b;
      mine[leaf] = acc;
    }
    __syncwarp();
  }
  __syncthreads();
  for (uint32_t l = threadIdx.x; l < nl; l += blockDim.x) {
    long long a = 0, b = 0;
    for (uint32_t w = 0; w < nwarps; ++w) { const longlong2 x = tab[(size_t) w * nl + l]; a += x.x; b += x.y; }
    partials[(size_t) blockIdx.x * nl + l] = make_longlong2(a, b);
  }
}

// one warp per leaf: block partials summed (integers: any order)
__global__ void leaf_reduce_kernel(const longlong2 *__restrict__ partials, uint32_t nblocks, uint32_t nl,
                                   const unsigned long long *__restrict__ leafn, bool newton, const int *__restrict__ qexp,
                                   longlong2 *leafsum, double *leafval) {
  const uint32_t l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (l >= nl) return;
  long long a = 0, b = 0;
  for (uint32_t k = lane; k < nblocks; k += 32) { const longlong2 x = partials[(size_t) k * nl + l]; a += x.x; b += x.y; }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    leafsum[l] = make_longlong2(a, b);
    leafval[l] = leaf_value_of(make_longlong2(a, b), leafn[l], newton, qexp);
  }
}

}  // namespace qr
