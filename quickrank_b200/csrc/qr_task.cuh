// quickrank_b200 — the task record shared by the tree kernels and the multi-GPU plumbing.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

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
  uint32_t hist_nblk;    // histogram slices of this task
  uint32_t sq0;          // first squares partial of this task
  uint32_t fused_sq;     // REFERENCE: 1 = fma chain (child ctor), 0 = mul+add (root update)
  uint32_t lcount;       // local left count when the host knows it (single GPU), see lc_known
  uint32_t lc_known;     // 0: kernels read the count computed by part_prefix_kernel instead
  uint32_t pad0;
  double parent_squares;
};

// Small rounds carry their task records inside the kernel parameters (constant bank): no
// host-to-device copy sits between the host's decision and the round's first kernel.
constexpr uint32_t kPackTasks = 12;
struct TaskPack { uint32_t n; uint32_t pad; NodeTask t[kPackTasks]; };

constexpr uint32_t kPartItems = 2048;   // documents per partition block
constexpr uint32_t kSqParts = 16;       // squares partials per task (FAST)

// 128-bit unsigned accumulator for the exact sum of squared fixed-point pseudo-responses
struct U128 { unsigned long long lo, hi; };
__device__ __forceinline__ void u128_add(U128 &a, unsigned long long lo, unsigned long long hi) {
  a.lo += lo;
  a.hi += hi + (a.lo < lo);
}

}  // namespace qr
