// quickrank_b200 — device-side controller of leaf-wise growth (single GPU).
//
// qr_tree_host.cuh replays RegressionTree::fit's heap (rt.cc:49-84) on the HOST between growth
// rounds: every round then costs a device->host->device round trip (results out, next tasks in,
// three launches) during which the GPU idles.  Here the same replay runs on the DEVICE, by one
// thread, at the end of the round's last finalize block ("grow_step"): it ingests the round's split
// results, advances the heap replay exactly as the host version does, and writes the next round's
// NodeTask records and a RoundHdr straight into device memory.  The host merely keeps a couple of
// rounds of (partition, histogram, finalize) launches queued ahead with upper-bound grids; blocks
// beyond the header's counts exit at once.  Progress and the "tree complete" flag reach the host
// through mapped pinned memory, so no stream synchronisation happens inside a tree.
//
// STATUS: opt-in (QR_DEVICE_GROWTH=1).  Correct and tested, but on B200 the replay kernel takes ~20 us
// per round — no faster than the host round trip it removes — so the host-driven rounds stay the default.
//
// The decisions are the host version's, statement for statement (same heap sift rules, same
// candidate order, same slot allocation order), so both paths grow identical trees
// (tests/test_gpu_parity.py::test_device_growth_equals_host_growth).
#pragma once

#include <cfloat>

#include "qr_internal.cuh"
#include "qr_task.cuh"

namespace qr {

struct LeafSeg { uint32_t lo, n; uint32_t buf; uint32_t blk0; };  // buf 2 = identity (unsplit root)
constexpr uint32_t kLeafItems = 4096;   // documents per block of the FAST leaf pass

struct RoundHdr {
  uint32_t ntasks;        // node expansions of the coming round (0: nothing left to do)
  uint32_t part_blocks;   // flat partition blocks of the round
  uint32_t hist_slices;   // flat histogram slices of the round
  uint32_t tasks_done;    // finalize: tasks whose result is published (the last one runs grow_step)
  uint32_t steps;         // grow_steps completed for this tree
  uint32_t done;          // tree complete
  uint32_t error;         // 1: histogram pool exhausted, 2: node table full, 3: slice table full
  uint32_t nleaves;       // leaves of the finished tree = entries of segs[]
  uint32_t leaf_blocks;   // blocks of leaf_partial_kernel
  uint32_t pad;
};

struct DevNode {          // mirror of HostNode
  uint32_t lo, n;
  int32_t buf, hist, left, right;
  uint32_t expanded, pushed;
  SplitResult res;
};

struct GrowOut {          // mapped pinned memory: what the host needs without synchronising the stream
  volatile uint32_t steps;     // grow_steps completed for the current tree
  volatile uint32_t done;
  volatile uint32_t error;
  volatile uint32_t nnodes, nleaves, nsplits, nrounds;
  volatile double rho, sigma, beta;
};

struct GrowState {
  // configuration
  uint32_t nleaves, max_tasks, max_nodes, nslots, want_slices, max_slices, exact, min_dpb;
  double n_global;
  // replay state
  uint32_t nnodes, heap_size, taken, root_done;
  int32_t nfree;
  double rho, sigma, beta;
  uint32_t nsplits, nrounds;
  // arrays
  DevNode *nodes;       // [max_nodes]
  double *heap_key;     // [max_nodes + 1], entry 0 is the DBL_MAX sentinel of maxheap.h
  int32_t *heap_val;    // [max_nodes + 1]
  int32_t *free_slots;  // [nslots] stack
  int32_t *S;           // [max_tasks] nodes expanded by the coming round
  double *cand_key;     // [max_nodes]
  int32_t *cand_val;    // [max_nodes]
  int32_t *stack;       // [max_nodes] DFS stack for the leaf order
};

__device__ inline int32_t grow_alloc_slot(GrowState *g) { return g->nfree > 0 ? g->free_slots[--g->nfree] : -1; }
__device__ inline void grow_release_slot(GrowState *g, int32_t &s) {
  if (s >= 0) g->free_slots[g->nfree++] = s;
  s = -1;
}
// MaxHeap<RTNode*>::push / pop (maxheap.h:58-86)
__device__ inline void grow_heap_push(GrowState *g, double key, int32_t val) {
  uint32_t p = ++g->heap_size;
  while (key > g->heap_key[p >> 1]) { g->heap_key[p] = g->heap_key[p >> 1]; g->heap_val[p] = g->heap_val[p >> 1]; p >>= 1; }
  g->heap_key[p] = key; g->heap_val[p] = val;
}
__device__ inline void grow_heap_pop(GrowState *g) {
  const double lk = g->heap_key[g->heap_size];
  const int32_t lv = g->heap_val[g->heap_size];
  --g->heap_size;
  uint32_t child, p = 1;
  while ((p << 1) <= g->heap_size) {
    child = p << 1;
    if (child < g->heap_size && g->heap_key[child + 1] > g->heap_key[child]) ++child;
    if (lk < g->heap_key[child]) { g->heap_key[p] = g->heap_key[child]; g->heap_val[p] = g->heap_val[child]; }
    else break;
    p = child;
  }
  g->heap_key[p] = lk; g->heap_val[p] = lv;
}
__device__ inline bool grow_can_split(const GrowState *g, int32_t i) {
  const SplitResult &r = g->nodes[i].res;
  return r.deviance > 0.0 && r.valid;   // rt.cc:212, 312
}
__device__ inline void grow_push_children(GrowState *g, int32_t i) {
  DevNode &nd = g->nodes[i];
  nd.pushed = 1;
  grow_heap_push(g, g->nodes[nd.left].res.deviance, nd.left);     // rt.cc:59-60, 72-73
  grow_heap_push(g, g->nodes[nd.right].res.deviance, nd.right);
  g->rho += (double) g->nodes[nd.left].res.n / g->n_global;
  g->sigma += (double) nd.res.n / g->n_global;
  g->nsplits++;
}

// starts a tree: empty heap, all slots free, the root node and its whole-node histogram task
__global__ void grow_init_kernel(GrowState *g, RoundHdr *hdr, NodeTask *tasks, uint32_t N, uint32_t root_dpb,
                                 uint32_t *ticket) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  g->nnodes = 1; g->heap_size = 0; g->taken = 0; g->root_done = 0;
  g->heap_key[0] = DBL_MAX; g->heap_val[0] = -1;
  g->nfree = 0;
  for (int32_t i = (int32_t) g->nslots - 1; i >= 0; --i) g->free_slots[g->nfree++] = i;
  g->rho = g->sigma = 0.0; g->beta = 1.0; g->nsplits = 0; g->nrounds = 0;
  DevNode root{};
  root.lo = 0; root.n = N; root.buf = 2; root.left = root.right = -1;
  root.hist = grow_alloc_slot(g);
  g->nodes[0] = root;
  NodeTask t{};
  t.lo = 0; t.n = N; t.src = 2; t.dst = 0; t.whole = 1; t.build_left = 1;
  t.slotP = -1; t.slotB = root.hist; t.slotD = -1;
  t.hist_dpb = root_dpb;
  t.hist_nblk = max(1u, (N + root_dpb - 1) / root_dpb);
  tasks[0] = t;
  RoundHdr h{};
  h.ntasks = 1; h.part_blocks = 0; h.hist_slices = t.hist_nblk;
  *hdr = h;
  *ticket = 0;   // (out->steps / done are reset by the host before this kernel is enqueued)
}

// One step of the replay of RegressionTree::fit (rt.cc:49-84); see fit_leafwise in qr_tree_host.cuh.
// One step of the replay of RegressionTree::fit (rt.cc:49-84); see fit_leafwise in qr_tree_host.cuh.
// One block.  The sequential part (heap replay, choice of the next expansion set) runs in thread 0
// on a compact copy of the state in SHARED memory; everything that touches the wide records in
// global memory (ingesting the round's 2k split results, writing the next k task records, updating
// the node table) is done by the block's threads in parallel.  (A first version ran the whole step
// as single-thread code over global memory: ~45 us per step, slower than the host round trip.)
//
// The header and the task records are double-buffered: round r reads copy r & 1, and the step that
// ends round r writes copy (r + 1) & 1, so a late-scheduled block of round r never sees round r + 1.
constexpr uint32_t kGrowThreads = 256;

__host__ __device__ inline size_t grow_smem_bytes(uint32_t max_nodes, uint32_t nslots, uint32_t max_tasks) {
  // s_dev, s_hkey, s_ckey (double) | s_rn, s_nn, s_left, s_hist, s_hval, s_cval, s_leaf (4 B) | s_free | per task 5 x 4 B | flags
  return (size_t) (max_nodes + 2) * (3 * 8 + 7 * 4 + 1) + (size_t) nslots * 4 + (size_t) max_tasks * 6 * 4 + 64;
}

struct GrowShared {
  double *dev, *hkey, *ckey;
  uint32_t *rn, *nn;
  int32_t *left, *hist, *hval, *cval, *leaf, *free_slots, *S, *slotB, *slotD;
  uint32_t *part0, *hist0, *nblk;
  uint8_t *flags;   // bit 0: can split, bit 1: expanded, bit 2: pushed
};

__device__ inline GrowShared grow_carve(unsigned char *smem, uint32_t M, uint32_t nslots, uint32_t mt) {
  GrowShared s;
  const uint32_t m = M + 2;
  double *d = reinterpret_cast<double *>(smem);
  s.dev = d; s.hkey = d + m; s.ckey = d + 2 * m;
  uint32_t *u = reinterpret_cast<uint32_t *>(d + 3 * m);
  s.rn = u; s.nn = u + m;
  int32_t *i = reinterpret_cast<int32_t *>(u + 2 * m);
  s.left = i; s.hist = i + m; s.hval = i + 2 * m; s.cval = i + 3 * m; s.leaf = i + 4 * m;
  s.free_slots = i + 5 * m;
  s.S = s.free_slots + nslots; s.slotB = s.S + mt; s.slotD = s.slotB + mt;
  s.part0 = reinterpret_cast<uint32_t *>(s.slotD + mt); s.hist0 = s.part0 + mt; s.nblk = s.hist0 + mt;
  s.flags = reinterpret_cast<uint8_t *>(s.nblk + mt);
  return s;
}

__device__ inline void sh_heap_push(GrowShared &s, uint32_t &size, double key, int32_t val) {
  uint32_t p = ++size;
  while (key > s.hkey[p >> 1]) { s.hkey[p] = s.hkey[p >> 1]; s.hval[p] = s.hval[p >> 1]; p >>= 1; }
  s.hkey[p] = key; s.hval[p] = val;
}
__device__ inline void sh_heap_pop(GrowShared &s, uint32_t &size) {
  const double lk = s.hkey[size];
  const int32_t lv = s.hval[size];
  --size;
  uint32_t child, p = 1;
  while ((p << 1) <= size) {
    child = p << 1;
    if (child < size && s.hkey[child + 1] > s.hkey[child]) ++child;
    if (lk < s.hkey[child]) { s.hkey[p] = s.hkey[child]; s.hval[p] = s.hval[child]; }
    else break;
    p = child;
  }
  s.hkey[p] = lk; s.hval[p] = lv;
}

__global__ void __launch_bounds__(kGrowThreads)
grow_step_kernel(GrowState *g, RoundHdr *hdr, RoundHdr *next_hdr, const NodeTask *__restrict__ tasks,
                 NodeTask *next_tasks, const SplitResult *__restrict__ res, uint32_t *ticket, LeafSeg *segs,
                 GrowOut *out) {
  extern __shared__ __align__(16) unsigned char grow_smem[];
  __shared__ uint32_t s_ns, s_done, s_error, s_nl, s_dpb;
  const uint32_t k = hdr->ntasks;
  if (k == 0) return;                      // queued beyond the end of the tree
  const uint32_t tid = threadIdx.x;
  DevNode *nodes = g->nodes;
  const uint32_t M = g->max_nodes, mt = g->max_tasks;
  GrowShared s = grow_carve(grow_smem, M, g->nslots, mt);
  const bool root_round = tasks[0].whole != 0;
  const uint32_t old_nnodes = g->nnodes;
  const uint32_t nnodes = root_round ? old_nnodes : old_nnodes + 2 * k;
  const uint32_t heap_size0 = g->heap_size;
  const int32_t nfree0 = g->nfree;

  // ---- A. results of the finished round -> node table (tail of expand_nodes), in parallel ----
  if (root_round) {
    if (tid == 0) nodes[0].res = res[0];
  } else {
    for (uint32_t j = tid; j < k; j += kGrowThreads) {
      const int32_t i = g->S[j];
      const NodeTask t = tasks[j];
      const uint32_t lc = (uint32_t) nodes[i].res.lcount;
      DevNode L{}, R{};
      L.lo = t.lo; L.n = lc; L.buf = (int32_t) t.dst;
      R.lo = t.lo + lc; R.n = t.n - lc; R.buf = (int32_t) t.dst;
      L.left = L.right = R.left = R.right = -1;
      L.hist = t.build_left ? t.slotB : t.slotD;
      R.hist = t.build_left ? t.slotD : t.slotB;
      L.res = res[2 * j];
      R.res = res[2 * j + 1];
      const int32_t li = (int32_t) (old_nnodes + 2 * j);
      nodes[li] = L;
      nodes[li + 1] = R;
      nodes[i].left = li; nodes[i].right = li + 1; nodes[i].expanded = 1;
    }
  }
  __syncthreads();   // global writes above are read back below by other threads of this block
  // ---- B. compact state -> shared memory ----
  for (uint32_t i = tid; i < nnodes; i += kGrowThreads) {
    const DevNode &nd = nodes[i];
    const double dv = nd.res.deviance;
    s.dev[i] = dv;
    s.rn[i] = (uint32_t) nd.res.n;
    s.nn[i] = nd.n;
    s.left[i] = nd.left;
    s.hist[i] = nd.hist;
    s.flags[i] = (uint8_t) (((dv > 0.0 && nd.res.valid) ? 1u : 0u) | (nd.expanded ? 2u : 0u) | (nd.pushed ? 4u : 0u));
  }
  for (uint32_t p = tid; p <= heap_size0; p += kGrowThreads) { s.hkey[p] = g->heap_key[p]; s.hval[p] = g->heap_val[p]; }
  for (int32_t q = (int32_t) tid; q < nfree0; q += (int32_t) kGrowThreads) s.free_slots[q] = g->free_slots[q];
  __syncthreads();

  // ---- C. sequential: heap replay, next expansion set, slot and block bookkeeping ----
  if (tid == 0) {
    uint32_t heap_size = heap_size0, taken = g->taken, root_done = g->root_done, nsplits = g->nsplits;
    int32_t nfree = nfree0;
    double rho = g->rho, sigma = g->sigma, beta = g->beta;
    const double ng = g->n_global;
    const uint32_t nleaves = g->nleaves;
    uint32_t error = 0;
    auto push_children = [&](int32_t i) {
      s.flags[i] |= 4u;
      const int32_t l = s.left[i];
      sh_heap_push(s, heap_size, s.dev[l], l);          // rt.cc:59-60, 72-73
      sh_heap_push(s, heap_size, s.dev[l + 1], l + 1);
      rho += (double) s.rn[l] / ng;
      sigma += (double) s.rn[i] / ng;
      nsplits++;
    };
    int32_t need = -1;
    if (!root_done) {
      if (s.flags[0] & 1u) {
        if (!(s.flags[0] & 2u)) need = 0;
        else { push_children(0); root_done = 1; }
      } else {
        root_done = 1;
      }
    }
    if (need < 0 && root_done) {
      while (heap_size != 0 && taken + heap_size < nleaves) {   // rt.cc:64-65
        const int32_t i = s.hval[1];
        if (s.flags[i] & 1u) {
          if (!(s.flags[i] & 2u)) { need = i; break; }
          sh_heap_pop(s, heap_size);
          push_children(i);
        } else {
          sh_heap_pop(s, heap_size);
          ++taken;                                              // rt.cc:78-79
        }
        if (s.hist[i] >= 0) { s.free_slots[nfree++] = s.hist[i]; s.hist[i] = -1; }   // rt.cc:83-84
      }
    }
    uint32_t ns = 0;
    if (need >= 0) {
      // expansion set: the blocking node plus the frontier nodes still reachable with the budget,
      // by decreasing deviance (ties: smaller node index)
      s.S[ns++] = need;
      if (need != 0) {
        const uint32_t budget = nleaves - taken - heap_size;   // successes left
        uint32_t nc = 0;
        for (uint32_t p = 1; p <= heap_size; ++p) {
          const int32_t i = s.hval[p];
          if (i != need && (s.flags[i] & 3u) == 1u) {
            const double key = s.hkey[p];
            uint32_t q = nc++;
            while (q > 0 && (key > s.ckey[q - 1] || (key == s.ckey[q - 1] && i < s.cval[q - 1]))) {
              s.ckey[q] = s.ckey[q - 1]; s.cval[q] = s.cval[q - 1]; --q;
            }
            s.ckey[q] = key; s.cval[q] = i;
          }
        }
        for (uint32_t q = 0; q < nc && ns < budget && ns < mt; ++q) s.S[ns++] = s.cval[q];
      }
      if (nnodes + 2 * ns > M) error = 2;
      // sizes of the built children are not in shared memory: left count = res.lcount of the node,
      // fetched in phase D; here only what needs a running total.  The left count of node i equals
      // the size of its (future) left child, which the partition derives from res.lcount — read it now.
      unsigned long long built_total = 0;
      for (uint32_t j = 0; j < ns; ++j) {
        const DevNode &nd = nodes[s.S[j]];
        const unsigned long long lc = nd.res.lcount, rc = nd.res.n - nd.res.lcount;
        const unsigned long long built = (g->exact || lc <= rc) ? lc : rc;
        s.nblk[j] = (uint32_t) built;                            // temporarily: built documents
        built_total += built;
      }
      unsigned long long dpb = (built_total + g->want_slices - 1) / g->want_slices;
      if (dpb < g->min_dpb) dpb = g->min_dpb;
      dpb = (dpb + 255ull) & ~255ull;
      if (dpb > (1ull << 20)) dpb = 1ull << 20;
      uint32_t part_blk = 0, hist_blk = 0;
      for (uint32_t j = 0; j < ns && !error; ++j) {
        if (nfree < 2) { error = 1; break; }
        s.slotB[j] = s.free_slots[--nfree];
        s.slotD[j] = s.free_slots[--nfree];
        s.part0[j] = part_blk;
        part_blk += max(1u, (s.nn[s.S[j]] + kPartItems - 1) / kPartItems);
        s.hist0[j] = hist_blk;
        const uint32_t built = s.nblk[j];
        beta += (double) built / ng;
        s.nblk[j] = max(1u, (uint32_t) ((built + dpb - 1) / dpb));
        hist_blk += s.nblk[j];
      }
      if (!error && hist_blk > g->max_slices) error = 3;
      s_dpb = (uint32_t) dpb;
      RoundHdr h{};
      h.steps = hdr->steps + 1;
      h.ntasks = ns; h.part_blocks = part_blk; h.hist_slices = hist_blk;
      if (!error) *next_hdr = h;
    }
    const bool done = need < 0 || error != 0;
    if (done) {
      // tree complete: keep only the splits the replay performed, list the leaves left to right
      uint32_t nl = 0;
      int32_t sp = 0;
      int32_t *stack = s.cval;                           // free now
      stack[sp++] = 0;
      while (sp > 0) {                                   // rtnode.cc:34-46
        const int32_t i = stack[--sp];
        if (!(s.flags[i] & 4u) || s.left[i] < 0) { s.leaf[nl++] = i; }
        else { stack[sp++] = s.left[i] + 1; stack[sp++] = s.left[i]; }
      }
      s_nl = nl;
    }
    g->nnodes = nnodes; g->heap_size = heap_size; g->taken = taken; g->root_done = root_done;
    g->nfree = nfree; g->rho = rho; g->sigma = sigma; g->beta = beta; g->nsplits = nsplits;
    if (!root_round) g->nrounds++;
    s_ns = ns; s_done = done ? 1u : 0u; s_error = error;
    // the heap can have grown past heap_size0: remember how much to write back
    s.hval[0] = (int32_t) heap_size;
    s.cval[M + 1] = nfree;
  }
  __syncthreads();

  // ---- D. write back, emit the next round's task records (head of expand_nodes), in parallel ----
  const uint32_t ns = s_ns, heap_size = (uint32_t) s.hval[0];
  const int32_t nfree = s.cval[M + 1];
  for (uint32_t i = tid; i < nnodes; i += kGrowThreads) {
    nodes[i].pushed = (s.flags[i] & 4u) ? 1u : 0u;
    nodes[i].hist = s.hist[i];
  }
  for (uint32_t p = tid + 1; p <= heap_size; p += kGrowThreads) { g->heap_key[p] = s.hkey[p]; g->heap_val[p] = s.hval[p]; }
  for (int32_t q = (int32_t) tid; q < nfree; q += (int32_t) kGrowThreads) g->free_slots[q] = s.free_slots[q];
  if (!s_done) {
    for (uint32_t j = tid; j < ns; j += kGrowThreads) {
      const int32_t i = s.S[j];
      const DevNode &nd = nodes[i];
      const unsigned long long lc = nd.res.lcount, rc = nd.res.n - nd.res.lcount;
      NodeTask t{};
      t.lo = nd.lo; t.n = nd.n; t.src = (uint32_t) nd.buf; t.dst = nd.buf == 2 ? 0u : (uint32_t) (1 - nd.buf);
      t.f = nd.res.feature; t.t = nd.res.threshold_idx;
      t.build_left = g->exact ? 1u : (lc <= rc ? 1u : 0u);
      t.whole = 0;
      t.slotP = s.hist[i];
      t.slotB = s.slotB[j];
      t.slotD = s.slotD[j];
      t.part_blk0 = s.part0[j];
      t.hist_blk0 = s.hist0[j];
      t.hist_dpb = s_dpb;
      t.hist_nblk = s.nblk[j];
      t.lcount = (uint32_t) lc;
      t.lc_known = 1u;
      t.sq0 = j;
      t.fused_sq = 1;
      t.parent_squares = nd.res.squares;
      next_tasks[j] = t;
      g->S[j] = i;
    }
  } else {
    const uint32_t nl = s_nl;
    // leaf segments in left-to-right order; nodes expanded speculatively but never split stay leaves
    for (uint32_t i = tid; i < nnodes; i += kGrowThreads)
      if (!(s.flags[i] & 4u)) { nodes[i].left = -1; nodes[i].right = -1; }
    if (tid == 0) {
      uint32_t blk = 0;
      for (uint32_t q = 0; q < nl; ++q) { s.rn[q] = blk; blk += (s.nn[s.leaf[q]] + kLeafItems - 1) / kLeafItems; }
      RoundHdr h{};
      h.steps = hdr->steps + 1;
      h.done = 1; h.error = s_error; h.nleaves = nl; h.leaf_blocks = blk;
      *next_hdr = h;
      // launches already queued beyond this round alternate between the two header copies: make
      // both say "nothing to do"
      hdr->ntasks = 0; hdr->part_blocks = 0; hdr->hist_slices = 0;
      out->nnodes = nnodes; out->nleaves = nl; out->nsplits = g->nsplits; out->nrounds = g->nrounds;
      out->rho = g->rho; out->sigma = g->sigma; out->beta = g->beta;
      out->error = s_error;
    }
    __syncthreads();
    for (uint32_t q = tid; q < nl; q += kGrowThreads) {
      const DevNode &nd = nodes[s.leaf[q]];
      segs[q] = LeafSeg{nd.lo, nd.n, (uint32_t) nd.buf, s.rn[q]};
    }
  }
  __syncthreads();
  if (tid == 0) {
    *ticket = 0;
    __threadfence_system();
    out->steps = hdr->steps + 1;
    __threadfence_system();
    if (s_done) out->done = 1;
    __threadfence_system();
  }
}

}  // namespace qr
