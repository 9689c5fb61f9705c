// quickrank_b200 — device kernels of the training hot path (sm_100a).
//
// Every kernel names the reference loop it replaces (paths relative to the reference root,
// hpclab/quickrank @ c569a59).  Compiled with -fmad=false: every fused multiply-add below is an
// explicit fma() placed where the reference's Release build (g++ 13.3, FMA target) fuses one, so
// results can be compared bit for bit with the reference build ("Floating-point contract").
#pragma once

#include <cfloat>
#include <cuda_runtime.h>
#include <stdint.h>

#include "qr_internal.cuh"

namespace qr {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// ------------------------------------------------------------------------------------------
// Bin access.  Panels: d_panels[p * N + doc] is one uint4 holding the bins of document `doc` for
// the FPP = 16 / sizeof(BinT) consecutive features of panel p ("feature columns ... laid out for
// coalesced 128-bit loads": consecutive lanes read consecutive documents of one panel).
// ------------------------------------------------------------------------------------------
template <typename BinT>
__device__ __forceinline__ uint32_t load_bin(const uint4 *panels, size_t N, uint32_t f, uint32_t doc) {
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  const BinT *row = reinterpret_cast<const BinT *>(panels + (size_t) (f / FPP) * N + doc);
  return row[f % FPP];
}

// rotate the 16 bytes of v right by rb bytes: result byte k = input byte (k + rb) & 15
__device__ __forceinline__ uint4 rotate_bytes(uint4 v, uint32_t rb) {
  if (rb & 4u) { uint32_t t = v.x; v.x = v.y; v.y = v.z; v.z = v.w; v.w = t; }
  if (rb & 8u) { uint32_t t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
  const uint32_t sel = 0x3210u + 0x1111u * (rb & 3u);
  uint4 r;
  r.x = __byte_perm(v.x, v.y, sel);
  r.y = __byte_perm(v.y, v.z, sel);
  r.z = __byte_perm(v.z, v.w, sel);
  r.w = __byte_perm(v.w, v.x, sel);
  return r;
}

template <typename BinT>
__device__ __forceinline__ uint32_t extract_bin(const uint4 &v, int j) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};  // j is a compile-time constant after unrolling
  if (sizeof(BinT) == 1) return (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
  return (w[j >> 1] >> ((j & 1) * 16)) & 0xffffu;
}

// ------------------------------------------------------------------------------------------
// Init: threshold lists and bin map (Mart::init mart.cc:117-176, RTRootHistogram
// rtnode_histogram.cc:227-253).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t flip_float_bits(uint32_t x) {  // radix.cc:28-30
  return x ^ ((uint32_t) (-(int32_t) (x >> 31)) | 0x80000000u);
}
__device__ __forceinline__ uint32_t unflip_float_bits(uint32_t x) {  // radix.cc:31-33
  return x ^ (((x >> 31) - 1u) | 0x80000000u);
}

__global__ void flip_keys_kernel(const float *x, uint32_t *keys, size_t n, int *bad) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (!(fabsf(v) <= FLT_MAX)) *bad = 1;  // NaN or +-inf: the reference's stmap is undefined for them
  keys[i] = flip_float_bits(__float_as_uint(v));
}

// flag[i] = 1 where a new distinct value starts: the reference keeps u[k] and appends v when
// u[k] < v (mart.cc:148-151); on an ascending list that is v[i-1] < v[i].
__global__ void distinct_flags_kernel(const uint32_t *sorted_keys, float *vals, uint8_t *flags, size_t n) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = __uint_as_float(unflip_float_bits(sorted_keys[i]));
  vals[i] = v;
  flags[i] = (i == 0) ? 1 : (__uint_as_float(unflip_float_bits(sorted_keys[i - 1])) < v);
}

__global__ void transpose_kernel(const float *rowmajor, float *colmajor, size_t N, size_t F) {
  __shared__ float tile[32][33];
  size_t f0 = (size_t) blockIdx.y * 32, d0 = (size_t) blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    size_t d = d0 + r, f = f0 + threadIdx.x;
    if (d < N && f < F) tile[r][threadIdx.x] = rowmajor[d * F + f];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    size_t f = f0 + r, d = d0 + threadIdx.x;
    if (d < N && f < F) colmajor[f * N + d] = tile[threadIdx.x][r];
  }
}

// bin(f, doc) = smallest t with x <= thr[f][t] (rtnode_histogram.cc:241-251).  One thread builds
// one document's 16-byte panel row.
template <typename BinT>
__global__ void binning_kernel(const float *colmajor, size_t N, uint32_t F, const float *thr,
                               const uint32_t *thr_off, uint4 *panels, uint32_t npanels, uint4 *rows) {
  constexpr uint32_t FPP = kPanelBytes / sizeof(BinT);
  size_t d = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t p = blockIdx.y;
  if (d >= N || p >= npanels) return;
  union { uint4 v; BinT b[FPP]; } row;
  row.v = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (uint32_t j = 0; j < FPP; ++j) {
    uint32_t f = p * FPP + j;
    if (f < F) {
      float x = colmajor[(size_t) f * N + d];
      const float *t = thr + thr_off[f];
      uint32_t lo = 0, hi = thr_off[f + 1] - thr_off[f] - 1;  // last threshold is FLT_MAX >= x
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (x <= t[mid]) hi = mid; else lo = mid + 1;
      }
      row.b[j] = (BinT) lo;
    }
  }
  panels[(size_t) p * N + d] = row.v;
  // second copy, document-major: a document's npanels rows are contiguous, so the histogram blocks of the
  // npanels panels that gather the same (sparse) document list share its DRAM bursts (FAST mode only)
  if (rows != nullptr) rows[d * npanels + p] = row.v;
}

// ------------------------------------------------------------------------------------------
// Per-query ranking: std::sort(idx, comp = score[i] > score[j]) with libstdc++'s introsort,
// reproduced move for move so that tied scores land where the reference puts them
// (QueryResults::indexing_of_sorted_labels, queryresults.cc:37-53; SURVEY.md section 7.1 "Sort").
// One warp per query.  Queries whose scores are pairwise distinct have a unique sorted order and
// are ranked in parallel by counting; only queries with ties take the sequential replica.
// Also evaluates DCG/NDCG of the query (dcg.cc:33-57, ndcg.cc:49-58).
// ------------------------------------------------------------------------------------------
// (score, position) pairs are sorted in place: one 16-byte shared-memory access per element moved
// or compared instead of an index load followed by a dependent score load.
struct __align__(16) SortElem { double key; uint32_t id; uint32_t pad; };

__device__ __forceinline__ bool se_comp(const SortElem &a, const SortElem &b) { return a.key > b.key; }

__device__ void se_adjust_heap(SortElem *v, int first, int hole, int len, SortElem value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (se_comp(v[first + child], v[first + child - 1])) child--;
    v[first + hole] = v[first + child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[first + hole] = v[first + child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && se_comp(v[first + parent], value)) {
    v[first + hole] = v[first + parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[first + hole] = value;
}

__device__ void se_heap_sort(SortElem *v, int first, int last) {
  int len = last - first;
  if (len >= 2) {
    int parent = (len - 2) / 2;
    for (;;) {
      const SortElem val = v[first + parent];
      se_adjust_heap(v, first, parent, len, val);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    const SortElem val = v[last];
    v[last] = v[first];
    se_adjust_heap(v, first, 0, last - first, val);
  }
}

__device__ __forceinline__ void se_unguarded_linear_insert(SortElem *v, int last) {
  const SortElem val = v[last];
  int next = last - 1;
  SortElem nv = v[next];
  while (se_comp(val, nv)) {
    v[last] = nv;
    last = next;
    --next;
    nv = v[next];
  }
  v[last] = val;
}

__device__ void se_insertion_sort(SortElem *v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (se_comp(v[i], v[first])) {
      const SortElem val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1];
      v[first] = val;
    } else {
      se_unguarded_linear_insert(v, i);
    }
  }
}

// sequential; executed by one lane
__device__ void se_std_sort(SortElem *v, int n) {
  if (n <= 0) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) ++lg;
  // __introsort_loop with an explicit stack (sub-ranges are disjoint, so their order is free)
  int st_first[64], st_last[64], st_depth[64];
  int sp = 0;
  st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
    while (last - first > 16) {
      if (depth == 0) { se_heap_sort(v, first, last); break; }
      --depth;
      const int mid = first + (last - first) / 2;
      {  // __move_median_to_first(first, first+1, mid, last-1)
        const int a = first + 1, b = mid, c = last - 1;
        int pick;
        const SortElem va = v[a], vb = v[b], vc = v[c];
        if (se_comp(va, vb)) {
          if (se_comp(vb, vc)) pick = b;
          else if (se_comp(va, vc)) pick = c;
          else pick = a;
        } else if (se_comp(va, vc)) pick = a;
        else if (se_comp(vb, vc)) pick = c;
        else pick = b;
        const SortElem t = v[first]; v[first] = v[pick]; v[pick] = t;
      }
      int lo = first + 1, hi = last;
      const SortElem pivot = v[first];
      for (;;) {  // __unguarded_partition(first+1, last, first)
        while (se_comp(v[lo], pivot)) ++lo;
        --hi;
        while (se_comp(pivot, v[hi])) --hi;
        if (!(lo < hi)) break;
        const SortElem t = v[lo]; v[lo] = v[hi]; v[hi] = t;
        ++lo;
      }
      if (sp < 64) { st_first[sp] = lo; st_last[sp] = last; st_depth[sp] = depth; ++sp; }
      last = lo;
    }
  }
  if (n > 16) {
    se_insertion_sort(v, 0, 16);
    for (int i = 16; i != n; ++i) se_unguarded_linear_insert(v, i);
  } else {
    se_insertion_sort(v, 0, n);
  }
}

// per-warp shared memory of rank_kernel: SortElem[maxlen] | uint32 idx[maxlen]
__host__ __device__ inline size_t rank_smem_per_warp(uint32_t maxlen) { return (size_t) maxlen * 20; }

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
rank_kernel(const double *__restrict__ scores, const float *__restrict__ labels,
            const double *__restrict__ gain, const uint32_t *__restrict__ qoff,
            const double *__restrict__ idcg, const double *__restrict__ lg, uint32_t Q,
            uint32_t maxlen, size_t cutoff, uint32_t *__restrict__ rankpos,
            double *__restrict__ qndcg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t q = blockIdx.x * WARPS + warp;
  if (q >= Q) return;
  // blockIdx.y: one of several score vectors over the same documents (the line search's candidates, qr_linesearch.cu):
  // vector y starts at scores + y * qoff[Q], its per-query values at qndcg + y * Q; rankpos is then not written
  if (gridDim.y > 1) {
    scores += (size_t) blockIdx.y * qoff[Q];
    qndcg += (size_t) blockIdx.y * Q;
  }
  SortElem *el = reinterpret_cast<SortElem *>(smem_raw) + (size_t) warp * maxlen;
  uint32_t *idx = reinterpret_cast<uint32_t *>(reinterpret_cast<SortElem *>(smem_raw) + (size_t) WARPS * maxlen) +
                  (size_t) warp * maxlen;
  const uint32_t off = qoff[q], n = qoff[q + 1] - off;
  for (uint32_t i = lane; i < n; i += 32) {   // queryresults.cc:50-51: identity, then sort
    SortElem e;
    e.key = scores[off + i]; e.id = i; e.pad = 0;
    el[i] = e;
  }
  __syncwarp();
  // parallel rank-by-counting, valid when no two scores are equal
  bool tie = false;
  for (uint32_t i = lane; i < n; i += 32) {
    const double si = el[i].key;
    uint32_t gt = 0, eq = 0;
    for (uint32_t j = 0; j < n; ++j) {
      const double sj = el[j].key;
      gt += sj > si;
      eq += sj == si;
    }
    if (eq > 1 || si != si) tie = true;
    else idx[gt] = i;
  }
  tie = __any_sync(0xffffffffu, tie);
  if (tie) {
    __syncwarp();
    if (lane == 0) se_std_sort(el, (int) n);
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32) idx[i] = el[i].id;
  }
  __syncwarp();
  if (gridDim.y == 1)
    for (uint32_t i = lane; i < n; i += 32) rankpos[off + i] = idx[i];
  if (lane == 0) {
    double r = 0.0;
    if (n > 0) {
      const double id = idcg[q];
      if (id > 0) {                                   // ndcg.cc:54-57
        const uint32_t size = cutoff < n ? (uint32_t) cutoff : n;
        double dcg = 0.0;
        for (uint32_t i = 0; i < size; ++i)           // dcg.cc:36-37
          dcg += (gain[off + idx[i]] - 1.0) / lg[i];
        r = dcg / id;
      }
    }
    qndcg[q] = r;
  }
}

// mean over queries (metric.h:96-105).  REFERENCE: sequential sum in query order.
__global__ void ndcg_mean_kernel(const double *qndcg, uint32_t Q, uint32_t Qdiv, bool exact, int qshift, double *out) {
  __shared__ double part[1024];
  if (exact) {
    // one warp, values fetched 32 at a time, summed in order by every lane; block b: the b-th vector of Q values
    qndcg += (size_t) blockIdx.x * Q;
    out += blockIdx.x;
    if (threadIdx.x >= 32) return;
    double acc = 0.0;
    for (uint32_t base = 0; base < Q; base += 32) {
      uint32_t i = base + lane_id();
      double v = i < Q ? qndcg[i] : 0.0;
      uint32_t cnt = min(32u, Q - base);
      for (uint32_t k = 0; k < cnt; ++k) acc += __shfl_sync(0xffffffffu, v, k);
    }
    if (threadIdx.x == 0) out[0] = Qdiv ? acc / (double) Qdiv : 0.0;
    return;
  }
  // FAST: an exact integer sum of the per-query values in fixed point (scale 2^qshift: Q_global values below 2 stay
  // under 2^63), so the mean is the same for any reduction shape and any sharding of the queries over GPUs
  long long acc = 0;
  for (uint32_t i = threadIdx.x; i < Q; i += blockDim.x) acc += __double2ll_rn(ldexp(qndcg[i], qshift));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  long long *parti = reinterpret_cast<long long *>(part);
  if (lane_id() == 0) parti[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) tot += parti[w];
    *reinterpret_cast<long long *>(out) = tot;   // the caller (all-reduces and) scales it
  }
}

// ------------------------------------------------------------------------------------------
// LambdaMART pseudo-responses (LambdaMart::compute_pseudoresponses, lambdamart.cc:62-152, with
// Ndcg::jacobian, ndcg.cc:60-92, evaluated on the fly instead of materialised).
//
// One warp per query.  All visited pairs have one member among the top c = min(cutoff, n) ranks
// ("heavy" ranks).  Pair terms are computed once per sweep by the lane that owns the other member
// and handed to the heavy rank's lane through shared memory; every document then adds its terms
// in exactly the order the reference's j/k loops produce them:
//   heavy X:  [j<X, label_j>label_X: -]  [k=0..n-1, label_X>label_k: +]  [j>X, label_j>label_X: -]
//   light b:  [a<c, label_a>label_b: -]  [a<c, label_b>label_a: +]
// ------------------------------------------------------------------------------------------
constexpr int kHG = 16;           // heavy ranks handled per sweep (upper bound; the host passes min(kHG, cutoff))

struct PairTerm { double rho, d; };

// exp() for the logistic term of lambdamart.cc:132-134: glibc's algorithm (sysdeps/ieee754/dbl-64/e_exp.c, the
// ARM optimized-routines exp since glibc 2.28: N = 128 table of (tail, scale bits), degree-5 polynomial), restated
// operation by operation with the multiply-adds fused the way the FMA build of libm (the ifunc variant an x86-64-v3
// host selects) fuses them.  Checked bit-for-bit against this image's libm on 2e7 random arguments
// (scripts/exp_glibc_check.c): the pseudo-responses are therefore bit-identical to the reference's at EVERY
// iteration, not just at scores = 0 — and it is 13 FP64 operations where the CUDA library exp() is ~45 (FP64 is
// the scarce pipe of this part).  |x| >= 512 and |x| < 2^-54 take glibc's special cases.
__device__ const unsigned long long kExpTable[256] = {
    0x0000000000000000ull, 0x3ff0000000000000ull, 0x3c9b3b4f1a88bf6eull, 0x3feff63da9fb3335ull,
    0xbc7160139cd8dc5dull, 0x3fefec9a3e778061ull, 0xbc905e7a108766d1ull, 0x3fefe315e86e7f85ull,
    0x3c8cd2523567f613ull, 0x3fefd9b0d3158574ull, 0xbc8bce8023f98efaull, 0x3fefd06b29ddf6deull,
    0x3c60f74e61e6c861ull, 0x3fefc74518759bc8ull, 0x3c90a3e45b33d399ull, 0x3fefbe3ecac6f383ull,
    0x3c979aa65d837b6dull, 0x3fefb5586cf9890full, 0x3c8eb51a92fdeffcull, 0x3fefac922b7247f7ull,
    0x3c3ebe3d702f9cd1ull, 0x3fefa3ec32d3d1a2ull, 0xbc6a033489906e0bull, 0x3fef9b66affed31bull,
    0xbc9556522a2fbd0eull, 0x3fef9301d0125b51ull, 0xbc5080ef8c4eea55ull, 0x3fef8abdc06c31ccull,
    0xbc91c923b9d5f416ull, 0x3fef829aaea92de0ull, 0x3c80d3e3e95c55afull, 0x3fef7a98c8a58e51ull,
    0xbc801b15eaa59348ull, 0x3fef72b83c7d517bull, 0xbc8f1ff055de323dull, 0x3fef6af9388c8deaull,
    0x3c8b898c3f1353bfull, 0x3fef635beb6fcb75ull, 0xbc96d99c7611eb26ull, 0x3fef5be084045cd4ull,
    0x3c9aecf73e3a2f60ull, 0x3fef54873168b9aaull, 0xbc8fe782cb86389dull, 0x3fef4d5022fcd91dull,
    0x3c8a6f4144a6c38dull, 0x3fef463b88628cd6ull, 0x3c807a05b0e4047dull, 0x3fef3f49917ddc96ull,
    0x3c968efde3a8a894ull, 0x3fef387a6e756238ull, 0x3c875e18f274487dull, 0x3fef31ce4fb2a63full,
    0x3c80472b981fe7f2ull, 0x3fef2b4565e27cddull, 0xbc96b87b3f71085eull, 0x3fef24dfe1f56381ull,
    0x3c82f7e16d09ab31ull, 0x3fef1e9df51fdee1ull, 0xbc3d219b1a6fbffaull, 0x3fef187fd0dad990ull,
    0x3c8b3782720c0ab4ull, 0x3fef1285a6e4030bull, 0x3c6e149289cecb8full, 0x3fef0cafa93e2f56ull,
    0x3c834d754db0abb6ull, 0x3fef06fe0a31b715ull, 0x3c864201e2ac744cull, 0x3fef0170fc4cd831ull,
    0x3c8fdd395dd3f84aull, 0x3feefc08b26416ffull, 0xbc86a3803b8e5b04ull, 0x3feef6c55f929ff1ull,
    0xbc924aedcc4b5068ull, 0x3feef1a7373aa9cbull, 0xbc9907f81b512d8eull, 0x3feeecae6d05d866ull,
    0xbc71d1e83e9436d2ull, 0x3feee7db34e59ff7ull, 0xbc991919b3ce1b15ull, 0x3feee32dc313a8e5ull,
    0x3c859f48a72a4c6dull, 0x3feedea64c123422ull, 0xbc9312607a28698aull, 0x3feeda4504ac801cull,
    0xbc58a78f4817895bull, 0x3feed60a21f72e2aull, 0xbc7c2c9b67499a1bull, 0x3feed1f5d950a897ull,
    0x3c4363ed60c2ac11ull, 0x3feece086061892dull, 0x3c9666093b0664efull, 0x3feeca41ed1d0057ull,
    0x3c6ecce1daa10379ull, 0x3feec6a2b5c13cd0ull, 0x3c93ff8e3f0f1230ull, 0x3feec32af0d7d3deull,
    0x3c7690cebb7aafb0ull, 0x3feebfdad5362a27ull, 0x3c931dbdeb54e077ull, 0x3feebcb299fddd0dull,
    0xbc8f94340071a38eull, 0x3feeb9b2769d2ca7ull, 0xbc87deccdc93a349ull, 0x3feeb6daa2cf6642ull,
    0xbc78dec6bd0f385full, 0x3feeb42b569d4f82ull, 0xbc861246ec7b5cf6ull, 0x3feeb1a4ca5d920full,
    0x3c93350518fdd78eull, 0x3feeaf4736b527daull, 0x3c7b98b72f8a9b05ull, 0x3feead12d497c7fdull,
    0x3c9063e1e21c5409ull, 0x3feeab07dd485429ull, 0x3c34c7855019c6eaull, 0x3feea9268a5946b7ull,
    0x3c9432e62b64c035ull, 0x3feea76f15ad2148ull, 0xbc8ce44a6199769full, 0x3feea5e1b976dc09ull,
    0xbc8c33c53bef4da8ull, 0x3feea47eb03a5585ull, 0xbc845378892be9aeull, 0x3feea34634ccc320ull,
    0xbc93cedd78565858ull, 0x3feea23882552225ull, 0x3c5710aa807e1964ull, 0x3feea155d44ca973ull,
    0xbc93b3efbf5e2228ull, 0x3feea09e667f3bcdull, 0xbc6a12ad8734b982ull, 0x3feea012750bdabfull,
    0xbc6367efb86da9eeull, 0x3fee9fb23c651a2full, 0xbc80dc3d54e08851ull, 0x3fee9f7df9519484ull,
    0xbc781f647e5a3ecfull, 0x3fee9f75e8ec5f74ull, 0xbc86ee4ac08b7db0ull, 0x3fee9f9a48a58174ull,
    0xbc8619321e55e68aull, 0x3fee9feb564267c9ull, 0x3c909ccb5e09d4d3ull, 0x3feea0694fde5d3full,
    0xbc7b32dcb94da51dull, 0x3feea11473eb0187ull, 0x3c94ecfd5467c06bull, 0x3feea1ed0130c132ull,
    0x3c65ebe1abd66c55ull, 0x3feea2f336cf4e62ull, 0xbc88a1c52fb3cf42ull, 0x3feea427543e1a12ull,
    0xbc9369b6f13b3734ull, 0x3feea589994cce13ull, 0xbc805e843a19ff1eull, 0x3feea71a4623c7adull,
    0xbc94d450d872576eull, 0x3feea8d99b4492edull, 0x3c90ad675b0e8a00ull, 0x3feeaac7d98a6699ull,
    0x3c8db72fc1f0eab4ull, 0x3feeace5422aa0dbull, 0xbc65b6609cc5e7ffull, 0x3feeaf3216b5448cull,
    0x3c7bf68359f35f44ull, 0x3feeb1ae99157736ull, 0xbc93091fa71e3d83ull, 0x3feeb45b0b91ffc6ull,
    0xbc5da9b88b6c1e29ull, 0x3feeb737b0cdc5e5ull, 0xbc6c23f97c90b959ull, 0x3feeba44cbc8520full,
    0xbc92434322f4f9aaull, 0x3feebd829fde4e50ull, 0xbc85ca6cd7668e4bull, 0x3feec0f170ca07baull,
    0x3c71affc2b91ce27ull, 0x3feec49182a3f090ull, 0x3c6dd235e10a73bbull, 0x3feec86319e32323ull,
    0xbc87c50422622263ull, 0x3feecc667b5de565ull, 0x3c8b1c86e3e231d5ull, 0x3feed09bec4a2d33ull,
    0xbc91bbd1d3bcbb15ull, 0x3feed503b23e255dull, 0x3c90cc319cee31d2ull, 0x3feed99e1330b358ull,
    0x3c8469846e735ab3ull, 0x3feede6b5579fdbfull, 0xbc82dfcd978e9db4ull, 0x3feee36bbfd3f37aull,
    0x3c8c1a7792cb3387ull, 0x3feee89f995ad3adull, 0xbc907b8f4ad1d9faull, 0x3feeee07298db666ull,
    0xbc55c3d956dcaebaull, 0x3feef3a2b84f15fbull, 0xbc90a40e3da6f640ull, 0x3feef9728de5593aull,
    0xbc68d6f438ad9334ull, 0x3feeff76f2fb5e47ull, 0xbc91eee26b588a35ull, 0x3fef05b030a1064aull,
    0x3c74ffd70a5fddcdull, 0x3fef0c1e904bc1d2ull, 0xbc91bdfbfa9298acull, 0x3fef12c25bd71e09ull,
    0x3c736eae30af0cb3ull, 0x3fef199bdd85529cull, 0x3c8ee3325c9ffd94ull, 0x3fef20ab5fffd07aull,
    0x3c84e08fd10959acull, 0x3fef27f12e57d14bull, 0x3c63cdaf384e1a67ull, 0x3fef2f6d9406e7b5ull,
    0x3c676b2c6c921968ull, 0x3fef3720dcef9069ull, 0xbc808a1883ccb5d2ull, 0x3fef3f0b555dc3faull,
    0xbc8fad5d3ffffa6full, 0x3fef472d4a07897cull, 0xbc900dae3875a949ull, 0x3fef4f87080d89f2ull,
    0x3c74a385a63d07a7ull, 0x3fef5818dcfba487ull, 0xbc82919e2040220full, 0x3fef60e316c98398ull,
    0x3c8e5a50d5c192acull, 0x3fef69e603db3285ull, 0x3c843a59ac016b4bull, 0x3fef7321f301b460ull,
    0xbc82d52107b43e1full, 0x3fef7c97337b9b5full, 0xbc892ab93b470dc9ull, 0x3fef864614f5a129ull,
    0x3c74b604603a88d3ull, 0x3fef902ee78b3ff6ull, 0x3c83c5ec519d7271ull, 0x3fef9a51fbc74c83ull,
    0xbc8ff7128fd391f0ull, 0x3fefa4afa2a490daull, 0xbc8dae98e223747dull, 0x3fefaf482d8e67f1ull,
    0x3c8ec3bc41aa2008ull, 0x3fefba1bee615a27ull, 0x3c842b94c3a9eb32ull, 0x3fefc52b376bba97ull,
    0x3c8a64a931d185eeull, 0x3fefd0765b6e4540ull, 0xbc8e37bae43be3edull, 0x3fefdbfdad9cbe14ull,
    0x3c77893b4d91cd9dull, 0x3fefe7c1819e90d8ull, 0x3c5305c14160cc89ull, 0x3feff3c22b8f71f1ull};

__device__ __forceinline__ double exp_lambda(double x, const unsigned long long *tab) {
  const uint32_t abstop = (uint32_t) ((unsigned long long) __double_as_longlong(x) >> 52) & 0x7ffu;
  if (abstop < 0x3c9u) return 1.0 + x;          // |x| < 2^-54
  if (abstop >= 0x408u) return exp(x);          // |x| >= 512: over/underflow handling (never reached by score differences)
  const double z = 0x1.71547652b82fep0 * 128.0 * x;                 // InvLn2N * x
  double kd = z + 0x1.8p52;                                          // round to nearest through the shift
  const unsigned long long ki = (unsigned long long) __double_as_longlong(kd);
  kd -= 0x1.8p52;
  const double r = fma(kd, -0x1.cf79abc9e3b3ap-47, fma(kd, -0x1.62e42fefa0000p-8, x));
  const unsigned long long idx = 2ull * (ki & 127ull), top = ki << 45;
  const double tail = __longlong_as_double((long long) tab[idx]);
  const unsigned long long sbits = tab[idx + 1] + top;
  const double r2 = r * r;
  // tail + r + r2 * (C2 + r * C3) + r2 * r2 * (C4 + r * C5), contracted left to right
  const double t1 = tail + r;
  const double t2 = fma(r, 0x1.555555555543cp-3, 0x1.ffffffffffdbdp-2);
  const double t3 = fma(r2, t2, t1);
  const double t4 = fma(r, 0x1.1111167a4d017p-7, 0x1.55555cf172b91p-5);
  const double tmp = fma(r2 * r2, t4, t3);
  const double scale = __longlong_as_double((long long) sbits);
  return fma(scale, tmp, scale);
}

// kept out of line: exp() and two FP64 divisions inline at every call site blow the loop body past
// the instruction cache
__device__ __noinline__ PairTerm pair_term(const double *s, const double *g, const double *invlg, const unsigned long long *exptab,
                                           double idcg, uint32_t c, uint32_t hi, uint32_t lo) {
  const uint32_t i = hi < lo ? hi : lo, j = hi < lo ? lo : hi;
  const double disc = (j < c) ? (invlg[j] - invlg[i]) : (-invlg[i]);       // ndcg.cc:76-86
  const double jac = disc * (g[i] - g[j]) / idcg;
  PairTerm t;
  t.d = fabs(jac);                                                          // lambdamart.cc:130
  t.rho = __drcp_rn(1.0 + exp_lambda(s[hi] - s[lo], exptab));               // lambdamart.cc:132-134 (IEEE reciprocal = 1.0 / x)
  return t;
}

// per-warp shared memory of lambda_kernel: s[maxlen] g[maxlen] (double) | stage[hg][32] (double2)
// | lab[maxlen] (float) pos[maxlen] (uint32)
__host__ __device__ inline size_t lambda_smem_per_warp(uint32_t maxlen, uint32_t hg) {
  return (size_t) maxlen * 24 + (size_t) hg * 32 * 16;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lambda_kernel(const double *__restrict__ scores, const float *__restrict__ labels,
              const double *__restrict__ gain, const uint32_t *__restrict__ qoff,
              const double *__restrict__ idcg_q, const double *__restrict__ invlg,
              const uint32_t *__restrict__ rankpos, uint32_t Q, uint32_t maxlen, size_t cutoff,
              uint32_t hg, double *lam, double *wgt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned long long s_exptab[256];
  for (uint32_t i = threadIdx.x; i < 256; i += WARPS * 32) s_exptab[i] = kExpTable[i];
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t q = blockIdx.x * WARPS + warp;
  if (q >= Q) return;
  unsigned char *base = smem_raw + (size_t) warp * lambda_smem_per_warp(maxlen, hg);
  double *s = reinterpret_cast<double *>(base);
  double *g = s + maxlen;
  double2 *stage = reinterpret_cast<double2 *>(g + maxlen);   // stage[x * 32 + lane]
  float *lab = reinterpret_cast<float *>(stage + (size_t) hg * 32);
  uint32_t *pos = reinterpret_cast<uint32_t *>(lab + maxlen);

  const uint32_t off = qoff[q], n = qoff[q + 1] - off;
  for (uint32_t r = lane; r < n; r += 32) {
    const uint32_t p = rankpos[off + r], d = off + p;
    s[r] = scores[d];
    g[r] = gain[d];
    lab[r] = labels[d];
    pos[r] = d;
    lam[d] = 0.0;                                   // lambdamart.cc:77-78
    wgt[d] = 0.0;
  }
  __syncwarp();
  if (n == 0) return;
  const double idcg = idcg_q[q];
  if (!(idcg > 0.0)) return;                        // ndcg.cc:69-70: jacobian stays zero
  const uint32_t c = cutoff < n ? (uint32_t) cutoff : n;

  for (int sweep = 0; sweep < 2; ++sweep) {
    for (uint32_t h0 = 0; h0 < c; h0 += hg) {
      const uint32_t hc = min(hg, c - h0);
      const bool heavy_lane = lane < hc;
      const uint32_t X = h0 + lane;                 // this lane's heavy rank (if heavy_lane)
      double aL = 0.0, aW = 0.0;
      if (sweep == 1 && heavy_lane) { aL = lam[pos[X]]; aW = wgt[pos[X]]; }

      if (sweep == 0) {
        // S1: b < X with label_b > label_X  ->  p[X] -= lambda(b, X)
        for (uint32_t cb = 0; cb < h0 + hc; cb += 32) {
          const uint32_t b = cb + lane;
          const float labb = b < n ? lab[b] : 0.f;
          uint32_t mymask = 0;
#pragma unroll 1
          for (uint32_t x = 0; x < hc; ++x) {
            const uint32_t Xx = h0 + x;
            const bool app = b < Xx && labb > lab[Xx];
            if (app) {
              const PairTerm t = pair_term(s, g, invlg, s_exptab, idcg, c, b, Xx);
              stage[x * 32 + lane] = make_double2(t.rho, t.d);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, app);
            if (lane == x) mymask = m;
          }
          __syncwarp();
          while (mymask) {
            const int bi = __ffs(mymask) - 1;
            mymask &= mymask - 1;
            const double2 t = stage[lane * 32 + bi];
            aL = fma(-t.x, t.y, aL);                             // lambdamart.cc:138
            aW = fma((1.0 - t.x) * t.x, t.y, aW);                // lambdamart.cc:140
          }
          __syncwarp();
        }
      }
      // S2 (sweep 0): all b with label_X > label_b  ->  p[X] += lambda(X, b); light b: p[b] -= ...
      // S3 (sweep 1): b > X with label_b > label_X  ->  p[X] -= lambda(b, X); light b: p[b] += ...
      const uint32_t cb0 = sweep == 0 ? 0u : (h0 & ~31u);
      for (uint32_t cb = cb0; cb < n; cb += 32) {
        const uint32_t b = cb + lane;
        const bool bvalid = b < n;
        const bool light = bvalid && b >= c;
        const float labb = bvalid ? lab[b] : 0.f;
        double bL = 0.0, bW = 0.0;
        if (light) { bL = lam[pos[b]]; bW = wgt[pos[b]]; }
        uint32_t mymask = 0;
#pragma unroll 1
        for (uint32_t x = 0; x < hc; ++x) {
          const uint32_t Xx = h0 + x;
          bool app = false;
          if (bvalid && b != Xx) {
            const float lx = lab[Xx];
            app = sweep == 0 ? (lx > labb) : (b > Xx && labb > lx);
          }
          if (app) {
            const PairTerm t = sweep == 0 ? pair_term(s, g, invlg, s_exptab, idcg, c, Xx, b)
                                          : pair_term(s, g, invlg, s_exptab, idcg, c, b, Xx);
            stage[x * 32 + lane] = make_double2(t.rho, t.d);
            if (light) {
              // lambdamart.cc:137-140 seen from the non-heavy member of the pair
              bL = fma(sweep == 0 ? -t.rho : t.rho, t.d, bL);
              bW = fma((1.0 - t.rho) * t.rho, t.d, bW);
            }
          }
          const uint32_t m = __ballot_sync(0xffffffffu, app);
          if (lane == x) mymask = m;
        }
        if (light) { lam[pos[b]] = bL; wgt[pos[b]] = bW; }
        __syncwarp();
        while (mymask) {
          const int bi = __ffs(mymask) - 1;
          mymask &= mymask - 1;
          const double2 t = stage[lane * 32 + bi];
          aL = fma(sweep == 0 ? t.x : -t.x, t.y, aL);
          aW = fma((1.0 - t.x) * t.x, t.y, aW);
        }
        __syncwarp();
      }
      if (heavy_lane) { lam[pos[X]] = aL; wgt[pos[X]] = aW; }
      __syncwarp();
    }
  }
}

// MART residuals (Mart::compute_pseudoresponses, mart.cc:418-431)
__global__ void mart_pseudo_kernel(const double *scores, const float *labels, size_t N, double *lam) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) lam[i] = (double) labels[i] - scores[i];
}

// A document sample's view of the sampled context's scores (qr_sample_pull_scores): each document's own score, and
// the score it is RANKED by — lambdamart.cc:94 copies scores_on_training_[d], d the document's position within its
// query, not offset + d, so a sampled query is sorted by the scores of the first documents of the dataset.
__global__ void sample_gather_kernel(const double *__restrict__ full_scores, const uint32_t *__restrict__ src,
                                     const uint32_t *__restrict__ key, size_t N, double *scores, double *rankkey) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  scores[i] = full_scores[src[i]];
  if (key) rankkey[i] = full_scores[key[i]];
}

// The bin matrix of a document sample, gathered from the sampled context's panels (same thresholds, same layout):
// panels[p][i] = from[p][src[i]], plus the document-major copy when the sample keeps one.
__global__ void sample_gather_panels_kernel(const uint4 *__restrict__ from, size_t from_N, const uint32_t *__restrict__ src,
                                            size_t N, uint32_t npanels, uint4 *panels, uint4 *rows) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = blockIdx.y;
  if (i >= N) return;
  const uint4 v = from[(size_t) p * from_N + src[i]];
  panels[(size_t) p * N + i] = v;
  if (rows != nullptr) rows[i * npanels + p] = v;
}

// ---- fixed-point view of the pseudo-responses (FAST histogram mode) -----------------------
__global__ void maxabs_kernel(const double *lam, size_t N, unsigned long long *maxbits) {
  double m = 0.0;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t) gridDim.x * blockDim.x)
    m = fmax(m, fabs(lam[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane_id() == 0 && m > 0.0) atomicMax(maxbits, (unsigned long long) __double_as_longlong(m));
}

// qexp = 2^k scale such that N_total * max|q| < 2^62; thread 0: pseudo-responses, thread 1: their weights (Newton
// denominators of the leaf outputs, rt.cc:186-200)
__global__ void choose_scale_kernel(const unsigned long long *maxbits, int log2n_ceil, int *qexp) {
  double m = __longlong_as_double((long long) maxbits[threadIdx.x]);
  int e = 0;
  if (m > 0.0) frexp(m, &e);        // m = f * 2^e, f in [0.5, 1)
  qexp[threadIdx.x] = (62 - log2n_ceil) - e;    // |lam * 2^qexp| < 2^(62 - log2n)
}

// also puts every document back into the root (node 0) for the tree about to be grown
__global__ void quantize_kernel(const double *lam, size_t N, const int *qexp, long long *lamq, uint16_t *node) {
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    lamq[i] = __double2ll_rn(ldexp(lam[i], *qexp));
    node[i] = 0;
  }
}


}  // namespace qr
