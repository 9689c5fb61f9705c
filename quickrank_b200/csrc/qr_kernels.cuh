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
                               const uint32_t *thr_off, uint4 *panels, uint32_t npanels) {
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
__global__ void ndcg_mean_kernel(const double *qndcg, uint32_t Q, uint32_t Qdiv, bool exact, double *out) {
  __shared__ double part[1024];
  if (exact) {
    // one warp, values fetched 32 at a time, summed in order by every lane
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
  // deterministic: contiguous chunk per thread, then a fixed-shape tree
  uint32_t per = (Q + blockDim.x - 1) / blockDim.x;
  uint32_t b = threadIdx.x * per, e = min(Q, b + per);
  double acc = 0.0;
  for (uint32_t i = b; i < e; ++i) acc += qndcg[i];
  part[threadIdx.x] = acc;
  __syncthreads();
  for (uint32_t st = blockDim.x >> 1; st > 0; st >>= 1) {
    if (threadIdx.x < st) part[threadIdx.x] += part[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = Qdiv ? part[0] / (double) Qdiv : 0.0;
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

// kept out of line: exp() and two FP64 divisions inline at every call site blow the loop body past
// the instruction cache
__device__ __noinline__ PairTerm pair_term(const double *s, const double *g, const double *invlg,
                                           double idcg, uint32_t c, uint32_t hi, uint32_t lo) {
  const uint32_t i = hi < lo ? hi : lo, j = hi < lo ? lo : hi;
  const double disc = (j < c) ? (invlg[j] - invlg[i]) : (-invlg[i]);       // ndcg.cc:76-86
  const double jac = disc * (g[i] - g[j]) / idcg;
  PairTerm t;
  t.d = fabs(jac);                                                          // lambdamart.cc:130
  t.rho = 1.0 / (1.0 + exp(s[hi] - s[lo]));                                 // lambdamart.cc:132-134
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
              const PairTerm t = pair_term(s, g, invlg, idcg, c, b, Xx);
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
            const PairTerm t = sweep == 0 ? pair_term(s, g, invlg, idcg, c, Xx, b)
                                          : pair_term(s, g, invlg, idcg, c, b, Xx);
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

// ---- fixed-point view of the pseudo-responses (FAST histogram mode) -----------------------
__global__ void maxabs_kernel(const double *lam, size_t N, unsigned long long *maxbits) {
  double m = 0.0;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (size_t) gridDim.x * blockDim.x)
    m = fmax(m, fabs(lam[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane_id() == 0 && m > 0.0) atomicMax(maxbits, (unsigned long long) __double_as_longlong(m));
}

// qexp = 2^k scale such that N_total * max|q| < 2^62
__global__ void choose_scale_kernel(const unsigned long long *maxbits, int log2n_ceil, int *qexp) {
  double m = __longlong_as_double((long long) *maxbits);
  int e = 0;
  if (m > 0.0) frexp(m, &e);        // m = f * 2^e, f in [0.5, 1)
  *qexp = (62 - log2n_ceil) - e;    // |lam * 2^qexp| < 2^(62 - log2n)
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
