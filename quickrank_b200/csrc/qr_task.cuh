// quickrank_b200 — the task record shared by the tree kernels and the multi-GPU plumbing.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace qr {

// One node being expanded (RegressionTree::split, rt.cc:209-362): its document list is cut by
// bin(f, doc) <= t, the histogram of one child is built from that child's documents and the
// sibling's is derived as parent - built (rt.cc:337-347).
struct NodeTask {
  uint32_t lo, n;        // REFERENCE mode: the node's segment of the id buffer; FAST mode, whole node: n = local documents
  uint32_t src, dst;     // REFERENCE mode: id buffers 0 / 1, src == 2 means the identity list (root)
  uint32_t f, t;         // split feature and threshold index
  uint32_t build_left;   // 1: the left child's histogram is built from samples, 0: the right one's
  uint32_t whole;        // 1: histogram of the whole node (root refresh, mart.cc:335); no split
  int32_t slotP, slotB, slotD;  // histogram slots: parent, built child, derived child
  uint32_t node;         // FAST mode: id of the node being split (value of node_of_doc for its documents)
  uint32_t child0;       // FAST mode: id of the left child; the right child is child0 + 1
  uint32_t region0;      // FAST mode: where the built child's compacted (doc id, pseudo-response) list starts
  uint32_t part_blk0;    // first flat partition block of this task (REFERENCE mode)
  uint32_t hist_blk0;    // first flat histogram slice of this task
  uint32_t hist_dpb;     // documents per histogram slice
  uint32_t hist_nblk;    // histogram slices of this task
  uint32_t sq0;          // first squares partial of this task
  uint32_t fused_sq;     // REFERENCE: 1 = fma chain (child ctor), 0 = mul+add (root update)
  uint32_t lcount;       // local left count when the host knows it (single GPU), see lc_known
  uint32_t lc_known;     // 0: kernels read the count computed by part_prefix_kernel instead
  uint32_t walk;         // REFERENCE mode: 1 = the built child is large: hist_exact_walk_kernel (qr_exact_kernels.cuh)
  uint32_t sq_chunk0;    // REFERENCE mode: first chunk record of this task's ordered squares sum
  uint32_t stage1;       // sharded training, fused exchange: 1 + the staging slot the built child's LOCAL histogram
                         // is accumulated in (the split scan adds all ranks' staging slots into slotB); 0: the
                         // histogram is built in place in slotB
  double parent_squares;
};

// slot the histogram kernels accumulate the built child into
__host__ __device__ __forceinline__ int32_t build_slot(const NodeTask &t) { return t.stage1 ? (int32_t) t.stage1 - 1 : t.slotB; }

// ---- peer memory (sharded training on one NVSwitch node) ------------------------------------
constexpr int kMaxPeers = 8;
// What the split scan needs to read the peers' freshly built histograms itself (fused all-reduce + scan):
// every rank maps every other rank's histogram pool, squares partials and flag array over CUDA IPC.
struct PeerView {
  const unsigned long long *sum[kMaxPeers];   // histogram pools (sums) of all ranks, own pool at [rank]
  const uint32_t *cnt[kMaxPeers];             // histogram pools (counts)
  const ulonglong2 *sq[kMaxPeers];            // squares partials
  uint32_t *peer_flags[kMaxPeers];            // flag arrays: peer_flags[p][r] = last epoch rank r announced to p
  uint32_t *flags;                            // own flag array
  int rank, world;                            // world <= 1: nothing is exchanged inside the kernel
  uint32_t epoch;
  int with_counts;                            // 0: the counts are global already (root refresh)
  // feature-sliced exchange (leaf-wise growth): this rank adds up and scans only the features [f_lo, f_hi) and the
  // ranks exchange their per-child winners through the mailboxes (f_hi == 0: every rank scans every feature)
  uint32_t f_lo, f_hi;
  unsigned long long *mail[kMaxPeers];        // mailbox of every rank, own at [rank]
  uint32_t mail_tasks;                        // tasks per (parity, source rank) block of a mailbox
};
// A mailbox record is four 8-byte words, each carrying the round's epoch in its low half: a word is written and
// read whole, so a record whose four tags match is complete — no fence, no separate flag, one NVLink write of latency.
constexpr uint32_t kMailWords = 4;
__host__ __device__ __forceinline__ size_t mail_index(uint32_t epoch, int src, uint32_t mail_tasks, uint32_t task, int child) {
  return ((((size_t) (epoch & 1u) * kMaxPeers + (size_t) src) * mail_tasks + task) * 2 + (size_t) child) * kMailWords;
}

// flags: release store (everything this thread wrote or observed before it, at system scope) / acquire load
__device__ __forceinline__ uint32_t ld_flag(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// bounded spin on a peer's flag: a peer that never arrives (crashed process) must neither hang the GPU nor kill every
// rank's context with a trap; the timeout is reported through `err` (mapped host memory) and surfaces as QR_ECOMM
__device__ __forceinline__ void wait_flag_or_report(const uint32_t *p, uint32_t epoch, uint32_t *err) {
  const long long t0 = clock64();
  while ((int32_t) (ld_flag(p) - epoch) < 0) {
    if (clock64() - t0 > 40000000000ll) {   // ~20 s
      if (err) { *reinterpret_cast<volatile uint32_t *>(err) = 1u; __threadfence_system(); }
      return;
    }
  }
}

// Small rounds carry their task records inside the kernel parameters (constant bank): no
// host-to-device copy sits between the host's decision and the round's first kernel.
constexpr uint32_t kPackTasks = 12;
struct TaskPack { uint32_t n; uint32_t pad; NodeTask t[kPackTasks]; };

constexpr uint32_t kPartItems = 2048;   // documents per partition block (REFERENCE mode)

struct LeafSeg { uint32_t lo, n; uint32_t buf; uint32_t blk0; };  // REFERENCE mode leaf pass; buf 2 = identity (unsplit root)

// 128-bit unsigned accumulator for the exact sum of squared fixed-point pseudo-responses
struct U128 { unsigned long long lo, hi; };
__device__ __forceinline__ void u128_add(U128 &a, unsigned long long lo, unsigned long long hi) {
  a.lo += lo;
  a.hi += hi + (a.lo < lo);
}

// FAST mode: leaf output from the exact fixed-point sums of the pseudo-responses (scale qexp[0]) and of their weights
// (scale qexp[1]): rt.cc:178 mean of the pseudo-responses, rt.cc:200 Newton step
__device__ __forceinline__ double leaf_value_of(longlong2 s, unsigned long long n, bool newton, const int *qexp) {
  const double a = (double) s.x * ldexp(1.0, -qexp[0]);
  if (!newton) return a / (double) n;
  const double b = (double) s.y * ldexp(1.0, -qexp[1]);
  return b >= DBL_EPSILON ? a / b : 0.0;
}

}  // namespace qr
