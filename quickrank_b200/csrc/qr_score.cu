// quickrank_b200 — ensemble scoring (the quickscore path).
//
// Replaces LTR_Algorithm::score_dataset (ltr_algorithm.cc:44-52) -> Ensemble::score_instance
// (ensemble.cc:111-118) -> RTNode::score_instance (rtnode.h:134-152) of the reference: for every
// row-major document, sum over trees (in tree order, FP64, product and sum rounded separately as in
// the reference build) of weight_t * leaf_t(doc), the node test being `x[feature] <= threshold` on
// floats.
//
// B200-first formulation.  The float test only ever compares a feature value with one of the
// ensemble's own thresholds, so each document is first re-coded, per feature the model uses, as
//     code(x) = number of distinct thresholds of that feature that are < x
// (one byte when a feature has <= 255 distinct thresholds, two otherwise): then
//     x <= threshold_k   <=>   code(x) <= k        (k = rank of the threshold in the sorted list)
// exactly, NaN included (code = list size, never <= k: the reference's `<=` is false for NaN).
// A 700-feature document shrinks from 2.8 KB of floats to 700 bytes, so a tile of documents lives in
// shared memory next to two buffers of ensemble chunks (bulk asynchronous copies, one being
// walked while the next arrives); 1, 2 or 4 threads share a document, each walking its own trees of
// the chunk, and add the weighted leaves in tree order.  The walk is bound by shared-memory
// wavefronts (one 8-byte node + one code per level, both at data-dependent addresses), not by HBM.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "qr_internal.cuh"

namespace qr {

// 8-byte internal node.  Leaves are not nodes: a child reference with bit 0 set IS the leaf.
//   lo = (threshold rank << 16) | slot          so that  code <= rank  <=>  (code << 16) <= lo
//   hi = left | (right << 16), each a child reference:
//        bit 0 clear: byte offset of the child node from the chunk's first node (a multiple of 8)
//        bit 0 set  : (index of the leaf's weight*output product in the chunk's table) << 1 | 1
struct __align__(8) CodeNode {
  uint32_t lo, hi;
};

constexpr int kChunkTrees = 16;         // tree slots per chunk (unused slots hold a leaf-only tree worth +0.0)
constexpr uint32_t kMaxChunkNodes = 8191, kMaxChunkLeaves = 32767;

// One chunk of the ensemble as a single contiguous blob, fetched with ONE bulk asynchronous copy
// (cp.async.bulk -> UBLKCP) into one of two shared-memory buffers while the other is being walked:
//   [ChunkHeader | CodeNode nodes[] | double product[]]      product[0] = +0.0 (unused tree slots)
// product[] holds weight_t * leaf output, rounded once exactly as `leaf * weight` is in
// ensemble.cc:116 (the reference build does not fuse that multiply into the sum).
struct ChunkHeader {
  uint32_t prod_off;                    // byte offset of product[] inside the blob
  uint32_t t0, nt, pad;                 // first tree of the chunk and how many of the slots hold a tree
  uint16_t root[kChunkTrees];           // child reference of each tree's root
};
constexpr uint32_t kNodesOff = sizeof(ChunkHeader);
static_assert(sizeof(ChunkHeader) % 16 == 0, "chunk header must keep 16-byte alignment");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// code(x) for every (document, used feature): binary search in the feature's sorted thresholds.
// A block re-codes a tile of kEncDocs documents.  Rows are fetched coalesced into shared memory;
// then every warp takes one feature at a time with lane = document, so that the 32 searches of a
// warp walk the SAME threshold list (a handful of L1 lines per step instead of 32 scattered ones);
// the codes go back through shared memory so that the global rows are written coalesced.
constexpr int kEncDocs = 32, kEncSlots = 512, kEncThreads = 512;
// shared-memory rows hold `cols` = min(nslots, kEncSlots) entries, padded to an odd number of words
__host__ __device__ inline uint32_t enc_x_words(uint32_t cols) { return cols | 1u; }
__host__ __device__ inline uint32_t enc_code_words(uint32_t cols, uint32_t code_bytes) { return ((cols * code_bytes + 3u) / 4u) | 1u; }

template <typename CodeT>
__global__ void __launch_bounds__(kEncThreads)
encode_kernel(const float *__restrict__ docs, size_t n0, size_t n, size_t F, const uint32_t *__restrict__ slot_feature,
              const uint32_t *__restrict__ thr_off, const float *__restrict__ thr, uint32_t nslots, uint32_t stride,
              CodeT *__restrict__ codes) {
  extern __shared__ __align__(16) unsigned char enc_raw[];
  const uint32_t cols = min(nslots, (uint32_t) kEncSlots);
  const uint32_t xw = enc_x_words(cols), cw = enc_code_words(cols, (uint32_t) sizeof(CodeT));
  float *xs = reinterpret_cast<float *>(enc_raw);                     // [kEncDocs][xw]
  uint32_t *cs = reinterpret_cast<uint32_t *>(enc_raw) + kEncDocs * xw;   // [kEncDocs][cw] words of packed codes
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = kEncThreads / 32;
  const size_t d0 = (size_t) blockIdx.x * kEncDocs;
  const uint32_t nd = (uint32_t) min((size_t) kEncDocs, n - d0);
  for (uint32_t c0 = 0; c0 < nslots; c0 += kEncSlots) {
    const uint32_t cn = min((uint32_t) kEncSlots, nslots - c0);
    for (uint32_t r = warp; r < nd; r += nwarps) {
      const float *row = docs + (n0 + d0 + r) * F;
      for (uint32_t j = lane; j < cn; j += 32u) xs[r * xw + j] = row[slot_feature[c0 + j]];
    }
    __syncthreads();
    for (uint32_t j = warp; j < cn; j += nwarps) {
      const uint32_t t0 = thr_off[c0 + j], len = thr_off[c0 + j + 1] - t0;
      const float *t = thr + t0;
      if (lane < nd) {
        const float x = xs[lane * xw + j];
        // number of thresholds < x (all of them for NaN: the reference's `x <= t` is then false everywhere)
        uint32_t pos = 0;
        for (uint32_t step = len ? 1u << (31 - __clz(len)) : 0u; step; step >>= 1) {
          const uint32_t np = pos + step;
          if (np <= len && !(t[np - 1] >= x)) pos = np;
        }
        reinterpret_cast<CodeT *>(cs + lane * cw)[j] = (CodeT) pos;
      }
    }
    __syncthreads();
    const uint32_t words = (cn * (uint32_t) sizeof(CodeT) + 3u) / 4u;
    for (uint32_t r = warp; r < nd; r += nwarps) {
      const uint32_t *src = cs + r * cw;
      uint32_t *dst = reinterpret_cast<uint32_t *>(codes + (d0 + r) * stride + c0);
      for (uint32_t wd = lane; wd < words; wd += 32u) dst[wd] = src[wd];
    }
    __syncthreads();
  }
}

__device__ __forceinline__ uint2 lds_u64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
template <typename CodeT> __device__ __forceinline__ uint32_t lds_code(uint32_t row, uint32_t slot);
template <> __device__ __forceinline__ uint32_t lds_code<uint8_t>(uint32_t row, uint32_t slot) { return lds_u8(row + slot); }
template <> __device__ __forceinline__ uint32_t lds_code<uint16_t>(uint32_t row, uint32_t slot) { return lds_u16(row + 2u * slot); }

// TPD threads share a document and walk every TPD-th tree of a chunk.  TILE: the block's documents
// are staged in shared memory (rows padded to an odd number of 32-bit words so that the lanes of a
// warp, which read different documents, hit different banks); otherwise the codes are read from
// global memory.
// PARTIAL: also stores every tree's weighted output as a float, partial[doc][tree] — the per-tree score matrix of
// Ensemble::partial_scores_instance (ensemble.cc:121-131) cast to Feature as Driver::extract_partial_scores does
// (driver.cc:411-446), the input of CLEAVER and of the line search.
template <typename CodeT, int TPD, bool TILE, bool PARTIAL>
__global__ void __launch_bounds__(1024)
score_codes_kernel(const CodeT *__restrict__ codes, size_t n, uint32_t stride, uint32_t tile_words,
                   const unsigned char *__restrict__ chunks, uint32_t chunk_bytes, uint32_t nchunks,
                   uint32_t docs_per_block, double *__restrict__ scores, float *__restrict__ partial, uint32_t ntrees) {
  static_assert(kChunkTrees % TPD == 0, "a chunk is walked in whole groups of TPD trees");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];
  const uint32_t s_base = smem_u32(smem_raw);
  const uint32_t s_tile = s_base + 2u * chunk_bytes;

  const uint32_t sub = threadIdx.x % TPD;                  // which trees of a group this thread walks
  const uint32_t ldoc = threadIdx.x / TPD;                 // document within the block
  const size_t d0 = (size_t) blockIdx.x * docs_per_block;
  const size_t d = d0 + ldoc;
  const uint32_t ndocs = (uint32_t) min((size_t) docs_per_block, n - d0);

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && nchunks > 0) {
    mbar_expect_tx(&s_bar[0], chunk_bytes);
    bulk_copy_g2s(smem_raw, chunks, chunk_bytes, &s_bar[0]);
  }
  // rows past the end of the dataset walk document d0 and are not stored
  const uint32_t row = ldoc < ndocs ? ldoc : 0u;
  const uint32_t my = s_tile + row * tile_words * 4u;
  const CodeT *myg = codes + (d0 + row) * stride;
  if (TILE) {
    // word-wise copy into rows of tile_words (odd) 32-bit words
    const uint32_t row_words = stride * (uint32_t) sizeof(CodeT) / 4;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(codes + d0 * stride);
    uint32_t *dst = reinterpret_cast<uint32_t *>(smem_raw + 2 * (size_t) chunk_bytes);
    for (uint32_t r = threadIdx.x / 32u; r < ndocs; r += blockDim.x / 32u)
      for (uint32_t w = threadIdx.x % 32u; w < row_words; w += 32u) dst[r * tile_words + w] = src[r * row_words + w];
    __syncthreads();
  }
  double sum = 0.0;
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t b = c & 1u;
    // prefetch the next chunk into the other buffer (its previous contents were consumed before the
    // __syncthreads at the end of the previous iteration)
    if (threadIdx.x == 0 && c + 1 < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s_bar[b ^ 1u], chunk_bytes);
      bulk_copy_g2s(smem_raw + (size_t) (b ^ 1u) * chunk_bytes, chunks + (size_t) (c + 1) * chunk_bytes, chunk_bytes,
                    &s_bar[b ^ 1u]);
    }
    mbar_wait(&s_bar[b], (c >> 1) & 1u);
    const uint32_t buf = s_base + b * chunk_bytes;
    const uint32_t nodes = buf + kNodesOff;
    const uint32_t prod = buf + lds_u32(buf);
    const uint32_t roots = buf + (uint32_t) offsetof(ChunkHeader, root) + 2u * sub;
    const uint32_t ct0 = PARTIAL ? lds_u32(buf + (uint32_t) offsetof(ChunkHeader, t0)) : 0u;
    const uint32_t cnt = PARTIAL ? lds_u32(buf + (uint32_t) offsetof(ChunkHeader, nt)) : 0u;
#pragma unroll 1
    for (uint32_t g = 0; g < (uint32_t) kChunkTrees; g += TPD) {
      uint32_t cur = lds_u16(roots + 2u * g);
      while (!(cur & 1u)) {
        const uint2 nd = lds_u64(nodes + cur);
        const uint32_t slot = nd.x & 0xFFFFu;
        const uint32_t code = TILE ? lds_code<CodeT>(my, slot) : (uint32_t) myg[slot];
        cur = (code << 16) <= nd.x ? (nd.y & 0xFFFFu) : (nd.y >> 16);   // rtnode.h:141-144
      }
      const double val = lds_f64(prod + ((cur & 0xFFFEu) << 2));
      if (PARTIAL) { if (ldoc < ndocs && g + sub < cnt) partial[d * ntrees + ct0 + g + sub] = (float) val; }
      // ordered accumulation (ensemble.cc:116): the TPD lanes of a document add the group's products in
      // tree order; each lane performs every addition, so all of them hold the same running sum
      if (TPD == 1) {
        sum = __dadd_rn(sum, val);
      } else {
        const uint32_t gbase = (threadIdx.x & 31u) & ~(uint32_t) (TPD - 1);
#pragma unroll
        for (int j = 0; j < TPD; ++j) sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, val, gbase + j));
      }
    }
    __syncthreads();   // every thread is done with buffer b before it is refilled
  }
  if (ldoc < ndocs && sub == 0 && (!PARTIAL || scores != nullptr)) scores[d] = sum;
}

}  // namespace qr

struct qr_scorer {
  int device = 0;
  size_t ntrees = 0, F = 0;
  uint32_t nslots = 0, stride = 0;     // used features; codes per document row (padded to 16 bytes)
  int code_bytes = 1;
  uint32_t nchunks = 0, chunk_bytes = 0;
  unsigned char *d_chunks = nullptr;   // nchunks blobs of chunk_bytes
  uint32_t *d_slot_feature = nullptr, *d_thr_off = nullptr;
  float *d_thr = nullptr;
  void *d_codes = nullptr;             // scratch for one batch of documents
  size_t batch_docs = 0;
  cudaStream_t stream = nullptr;
  uint64_t launches = 0;
  // qr_score_dataset (host buffers): two staging slices, uploads on their own stream
  float *d_in[2] = {nullptr, nullptr};
  double *d_out[2] = {nullptr, nullptr};
  size_t io_docs = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
};

using namespace qr;

// shared-memory row length in 32-bit words: the global row, padded to an odd word count
static uint32_t score_tile_words(const qr_scorer *s) {
  const uint32_t w = s->stride * (uint32_t) s->code_bytes / 4;
  return w | 1u;
}

// Launch shape: TPD threads per document and documents per block.  Narrow rows leave room for many
// documents per SM, so one or two threads per document already fill the SM with warps; wide rows
// (hundreds of bytes) need TPD = 4 to reach the same warp count from the few hundred documents
// that fit.  docs == 0: not even 32 rows fit next to the two chunk buffers, the kernel then reads
// the codes from global memory.
struct ScoreShape { int tpd; uint32_t docs; };
static ScoreShape score_shape(const qr_scorer *s) {
  const size_t budget = 220 * 1024, fixed = 2 * (size_t) s->chunk_bytes + 1024, row = (size_t) score_tile_words(s) * 4;
  ScoreShape sh{4, 0};
  if (budget <= fixed) return sh;
  const size_t fit = (budget - fixed) / row / 32 * 32;
  if (fit < 32) return sh;
  double best = -1;
  const int tpds[] = {1, 2, 4};
  for (int tpd : tpds) {
    for (uint32_t docs = 1024 / tpd; docs >= 32; docs /= 2) {
      if (docs > fit) continue;
      // resident warps per SM: blocks limited by shared memory and by 2048 threads
      const size_t smem = fixed + docs * row;
      const size_t blocks = std::min<size_t>(std::min<size_t>(227 * 1024 / smem, 2048 / (docs * tpd)), 32);
      const double warps = (double) blocks * docs * tpd / 32;
      // more resident warps first; then fewer threads per document (less exchange), bigger blocks
      const double score = std::min(warps, 64.0) * 1000 - tpd * 10 + docs / 1024.0;
      if (score > best) { best = score; sh.tpd = tpd; sh.docs = docs; }
    }
  }
  return sh;
}

template <typename CodeT, int TPD, bool TILE>
static cudaError_t score_launch(qr_scorer *s, size_t n, uint32_t dpb, uint32_t tile_words, size_t smem, double *scores,
                                float *partial) {
  const unsigned sg = (unsigned) ((n + dpb - 1) / dpb);
  if (partial) {
    cudaFuncSetAttribute(score_codes_kernel<CodeT, TPD, TILE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    score_codes_kernel<CodeT, TPD, TILE, true><<<sg, dpb * TPD, smem, s->stream>>>(
        (const CodeT *) s->d_codes, n, s->stride, tile_words, s->d_chunks, s->chunk_bytes, s->nchunks, dpb, scores, partial,
        (uint32_t) s->ntrees);
  } else {
    cudaFuncSetAttribute(score_codes_kernel<CodeT, TPD, TILE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    score_codes_kernel<CodeT, TPD, TILE, false><<<sg, dpb * TPD, smem, s->stream>>>(
        (const CodeT *) s->d_codes, n, s->stride, tile_words, s->d_chunks, s->chunk_bytes, s->nchunks, dpb, scores, nullptr,
        (uint32_t) s->ntrees);
  }
  return cudaGetLastError();
}

template <typename CodeT>
static cudaError_t score_dispatch(qr_scorer *s, const ScoreShape &sh, size_t n, uint32_t dpb, uint32_t tile_words, size_t smem,
                                  double *scores, float *partial) {
  if (sh.docs == 0) return score_launch<CodeT, 4, false>(s, n, dpb, tile_words, smem, scores, partial);
  if (sh.tpd == 1) return score_launch<CodeT, 1, true>(s, n, dpb, tile_words, smem, scores, partial);
  if (sh.tpd == 2) return score_launch<CodeT, 2, true>(s, n, dpb, tile_words, smem, scores, partial);
  return score_launch<CodeT, 4, true>(s, n, dpb, tile_words, smem, scores, partial);
}

extern "C" {

// The weight a ranker() generated by the reference's conditional-operator generator multiplies a tree's leaf with:
// the XML weight read as a float and printed with three decimals and an `f` suffix
// (generate_conditional_operators.cc:95-105), i.e. the float nearest to that decimal, promoted to double.
static double condop_weight(double w) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.3f", (double) (float) w);
  return (double) strtof(buf, nullptr);
}

static int scorer_create_impl(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F, int device,
                              qr_scorer **out);

int qr_scorer_create(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F, int device,
                     qr_scorer **out) {
  return scorer_create_impl(trees, weights, ntrees, F, device, out);
}

int qr_scorer_create_ex(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F, int device,
                        unsigned flags, qr_scorer **out) {
  if (flags & ~(unsigned) QR_SCORER_CONDOP_WEIGHTS) { set_error("qr_scorer_create_ex: unknown flags 0x%x", flags); return QR_EINVAL; }
  if (!(flags & QR_SCORER_CONDOP_WEIGHTS) || !weights) return scorer_create_impl(trees, weights, ntrees, F, device, out);
  std::vector<double> w(ntrees);
  for (size_t t = 0; t < ntrees; ++t) w[t] = condop_weight(weights[t]);
  return scorer_create_impl(trees, w.data(), ntrees, F, device, out);
}

}  // extern "C"

static int scorer_create_impl(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F, int device,
                              qr_scorer **out) {
  if (!out) { set_error("qr_scorer_create: null out"); return QR_EINVAL; }
  *out = nullptr;
  if ((!trees || !weights) && ntrees) { set_error("qr_scorer_create: null argument"); return QR_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return QR_ENODEVICE;
  }
  // distinct thresholds per used feature
  std::map<uint32_t, std::vector<float>> thr_of;
  size_t max_tree_nodes = 1;
  for (size_t t = 0; t < ntrees; ++t) {
    const qr_flat_tree &ft = trees[t];
    if (ft.nnodes == 0) { set_error("tree %zu is empty", t); return QR_EINVAL; }
    if (ft.nnodes > 8191) { set_error("tree %zu has %u nodes; at most 8191 are supported", t, ft.nnodes); return QR_ELIMIT; }
    max_tree_nodes = std::max<size_t>(max_tree_nodes, ft.nnodes);
    for (uint32_t i = 0; i < ft.nnodes; ++i) {
      if (ft.feature[i] < 0) continue;
      if ((size_t) ft.feature[i] >= F || ft.left[i] < 0 || ft.right[i] < 0 || (uint32_t) ft.left[i] >= ft.nnodes ||
          (uint32_t) ft.right[i] >= ft.nnodes) {
        set_error("tree %zu node %u is malformed", t, i);
        return QR_EINVAL;
      }
      if (ft.threshold[i] != ft.threshold[i]) { set_error("tree %zu node %u has a NaN threshold", t, i); return QR_EINVAL; }
      thr_of[(uint32_t) ft.feature[i]].push_back(ft.threshold[i]);
    }
  }
  if (thr_of.size() > 65534) { set_error("the model uses %zu features; at most 65534 are supported", thr_of.size()); return QR_ELIMIT; }
  std::vector<uint32_t> slot_feature, thr_off(1, 0);
  std::vector<float> thr_flat;
  std::map<uint32_t, uint32_t> slot_of;
  size_t max_thr = 0;
  for (auto &kv : thr_of) {
    std::vector<float> &v = kv.second;
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    slot_of[kv.first] = (uint32_t) slot_feature.size();
    slot_feature.push_back(kv.first);
    thr_flat.insert(thr_flat.end(), v.begin(), v.end());
    thr_off.push_back((uint32_t) thr_flat.size());
    max_thr = std::max(max_thr, v.size());
  }
  if (max_thr > 65535) { set_error("a feature has %zu distinct thresholds; at most 65535 are supported", max_thr); return QR_ELIMIT; }

  // chunk blobs: up to kChunkTrees whole trees each, at most ~16 KB of nodes + products (a single
  // larger tree gets a chunk of its own); every blob is padded to the size of the largest one
  const size_t soft_payload = std::max<size_t>(16 * 1024, (max_tree_nodes + 1) * 8);
  struct Chunk { size_t t0, t1, inner, leaves; };
  std::vector<Chunk> plan;
  {
    Chunk cur{0, 0, 0, 0};
    for (size_t t = 0; t < ntrees; ++t) {
      size_t nl = 0;
      for (uint32_t i = 0; i < trees[t].nnodes; ++i) nl += trees[t].feature[i] < 0;
      const size_t ni = trees[t].nnodes - nl;
      if (cur.t1 > cur.t0 && (cur.t1 - cur.t0 >= (size_t) kChunkTrees || (cur.inner + ni + cur.leaves + nl) * 8 > soft_payload ||
                              cur.inner + ni > kMaxChunkNodes || cur.leaves + nl + 1 > kMaxChunkLeaves)) {
        plan.push_back(cur);
        cur = Chunk{t, t, 0, 0};
      }
      cur.t1 = t + 1;
      cur.inner += ni;
      cur.leaves += nl;
    }
    if (cur.t1 > cur.t0) plan.push_back(cur);
  }
  size_t chunk_bytes = sizeof(ChunkHeader) + 8;
  for (auto &c : plan) chunk_bytes = std::max(chunk_bytes, sizeof(ChunkHeader) + (c.inner + c.leaves + 1) * 8);
  chunk_bytes = (chunk_bytes + 127) & ~(size_t) 127;
  std::vector<unsigned char> blob(plan.size() * chunk_bytes, 0);
  for (size_t ci = 0; ci < plan.size(); ++ci) {
    const Chunk &c = plan[ci];
    unsigned char *base = blob.data() + ci * chunk_bytes;
    ChunkHeader *h = reinterpret_cast<ChunkHeader *>(base);
    h->prod_off = (uint32_t) (sizeof(ChunkHeader) + c.inner * 8);
    h->t0 = (uint32_t) c.t0;
    h->nt = (uint32_t) (c.t1 - c.t0);
    CodeNode *nodes = reinterpret_cast<CodeNode *>(base + kNodesOff);
    double *product = reinterpret_cast<double *>(base + h->prod_off);
    product[0] = 0.0;
    for (int k = 0; k < kChunkTrees; ++k) h->root[k] = 1;       // unused slot: leaf reference to product[0]
    size_t no = 0, lo = 1;
    for (size_t t = c.t0; t < c.t1; ++t) {
      const qr_flat_tree &ft = trees[t];
      // child reference of every node of this tree
      std::vector<uint16_t> ref(ft.nnodes);
      for (uint32_t i = 0; i < ft.nnodes; ++i) {
        if (ft.feature[i] >= 0) ref[i] = (uint16_t) (8 * no++);
        else {
          const volatile double prod = ft.value[i] * weights[t];   // rounded on its own (ensemble.cc:116)
          product[lo] = prod;
          ref[i] = (uint16_t) ((lo++ << 1) | 1u);
        }
      }
      h->root[t - c.t0] = ref[0];
      for (uint32_t i = 0; i < ft.nnodes; ++i) {
        if (ft.feature[i] < 0) continue;
        const uint32_t sidx = slot_of[(uint32_t) ft.feature[i]];
        const float *tb = thr_flat.data() + thr_off[sidx], *te = thr_flat.data() + thr_off[sidx + 1];
        const uint32_t rank = (uint32_t) (std::lower_bound(tb, te, ft.threshold[i]) - tb);   // rank of this threshold
        CodeNode &cn = nodes[ref[i] / 8];
        cn.lo = (rank << 16) | sidx;
        cn.hi = (uint32_t) ref[ft.left[i]] | ((uint32_t) ref[ft.right[i]] << 16);
      }
    }
  }

  qr_scorer *s = new qr_scorer();
  if (device >= 0) {
    if (device >= ndev) { delete s; set_error("device %d out of range", device); return QR_EINVAL; }
    cudaSetDevice(device);
  }
  cudaGetDevice(&s->device);
  s->ntrees = ntrees;
  s->F = F;
  s->nslots = (uint32_t) slot_feature.size();
  s->code_bytes = max_thr <= 255 ? 1 : 2;
  const uint32_t per16 = 16 / s->code_bytes;
  s->stride = std::max<uint32_t>(per16, (s->nslots + per16 - 1) / per16 * per16);
  s->nchunks = (uint32_t) plan.size();
  s->chunk_bytes = (uint32_t) chunk_bytes;
  *out = s;
  QR_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
#define QR_UPLOAD(dst, vec, T)                                                                       \
  QR_CUDA(cudaMalloc((void **) &dst, std::max<size_t>((vec).size(), 1) * sizeof(T)));                \
  QR_CUDA(cudaMemcpy(dst, (vec).data(), (vec).size() * sizeof(T), cudaMemcpyHostToDevice))
  QR_UPLOAD(s->d_chunks, blob, unsigned char);
  QR_UPLOAD(s->d_slot_feature, slot_feature, uint32_t);
  QR_UPLOAD(s->d_thr_off, thr_off, uint32_t);
  QR_UPLOAD(s->d_thr, thr_flat, float);
#undef QR_UPLOAD
  return QR_OK;
}

extern "C" {

int qr_scorer_destroy(qr_scorer *s) {
  if (!s) return QR_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
  void *ptrs[] = {s->d_chunks, s->d_slot_feature, s->d_thr_off, s->d_thr, s->d_codes, s->d_in[0], s->d_in[1], s->d_out[0], s->d_out[1]};
  for (void *p : ptrs) if (p) cudaFree(p);
  for (int b = 0; b < 2; ++b) {
    if (s->ev_in[b]) cudaEventDestroy(s->ev_in[b]);
    if (s->ev_free[b]) cudaEventDestroy(s->ev_free[b]);
  }
  if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return QR_OK;
}

// encode + walk of device-resident documents; `partial` (device, [N][ntrees] floats) may be null, and so may
// `scores` when `partial` is given
static int score_device_impl(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores, float *partial) {
  if (!s || !docs || (!scores && !partial)) { set_error("qr_score_dataset_device: null argument"); return QR_EINVAL; }
  if (F != s->F) { set_error("dataset has %zu features, the model was built for %zu", F, s->F); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  const size_t doc_bytes = (size_t) s->stride * s->code_bytes;
  const size_t batch = std::min<size_t>(N, std::max<size_t>(1024, ((size_t) 1 << 30) / doc_bytes / 1024 * 1024));
  if (s->batch_docs < batch) {
    if (s->d_codes) cudaFree(s->d_codes);
    s->d_codes = nullptr;
    QR_CUDA(cudaMalloc(&s->d_codes, batch * doc_bytes));
    QR_CUDA(cudaMemsetAsync(s->d_codes, 0, batch * doc_bytes, s->stream));
    s->batch_docs = batch;
  }
  const ScoreShape sh = score_shape(s);
  const bool tile = sh.docs != 0;
  const uint32_t dpb = tile ? sh.docs : 64;
  const uint32_t tile_words = score_tile_words(s);
  const size_t smem = 2 * (size_t) s->chunk_bytes + (tile ? (size_t) dpb * tile_words * 4 : 0);
  if (smem > 224 * 1024) { set_error("a chunk of the ensemble (%u bytes) does not fit in shared memory", s->chunk_bytes); return QR_ELIMIT; }
  for (size_t n0 = 0; n0 < N; n0 += batch) {
    const size_t n = std::min(batch, N - n0);
    if (s->nslots) {
      const unsigned eg = (unsigned) ((n + kEncDocs - 1) / kEncDocs);
      const uint32_t cols = std::min<uint32_t>(s->nslots, kEncSlots);
      const size_t esm = (size_t) kEncDocs * 4 * (enc_x_words(cols) + enc_code_words(cols, (uint32_t) s->code_bytes));
      if (s->code_bytes == 1) {
        cudaFuncSetAttribute(encode_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) esm);
        encode_kernel<uint8_t><<<eg, kEncThreads, esm, s->stream>>>(docs, n0, n, F, s->d_slot_feature, s->d_thr_off, s->d_thr,
                                                                    s->nslots, s->stride, (uint8_t *) s->d_codes);
      } else {
        cudaFuncSetAttribute(encode_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) esm);
        encode_kernel<uint16_t><<<eg, kEncThreads, esm, s->stream>>>(docs, n0, n, F, s->d_slot_feature, s->d_thr_off, s->d_thr,
                                                                     s->nslots, s->stride, (uint16_t *) s->d_codes);
      }
      s->launches += 1;
    }
    double *sc = scores ? scores + n0 : nullptr;
    float *pa = partial ? partial + n0 * s->ntrees : nullptr;
    QR_CUDA(s->code_bytes == 1 ? score_dispatch<uint8_t>(s, sh, n, dpb, tile_words, smem, sc, pa)
                               : score_dispatch<uint16_t>(s, sh, n, dpb, tile_words, smem, sc, pa));
    s->launches += 1;
  }
  return QR_OK;
}

int qr_score_dataset_device(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!scores) { set_error("qr_score_dataset_device: null argument"); return QR_EINVAL; }
  return score_device_impl(s, docs, N, F, scores, nullptr);
}

int qr_score_partial_device(qr_scorer *s, const float *docs, size_t N, size_t F, float *partial, double *scores) {
  if (!partial) { set_error("qr_score_partial_device: null argument"); return QR_EINVAL; }
  return score_device_impl(s, docs, N, F, scores, partial);
}

// Host buffers: the per-tree score matrix [N][ntrees] (floats) of Driver::extract_partial_scores
// (driver.cc:411-446); `scores` (may be NULL) also receives the ensemble scores.  Slices of at most 256 MB of
// output pass through the device one after the other.
int qr_score_partial(qr_scorer *s, const float *docs, size_t N, size_t F, float *partial, double *scores) {
  if (!s || !docs || !partial) { set_error("qr_score_partial: null argument"); return QR_EINVAL; }
  if (F != s->F) { set_error("dataset has %zu features, the model was built for %zu", F, s->F); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0 || s->ntrees == 0) return QR_OK;
  const size_t row = std::max(F * sizeof(float), s->ntrees * sizeof(float));
  const size_t slice = std::min(N, std::max<size_t>(1024, ((size_t) 256 << 20) / row));
  float *d_in = nullptr, *d_part = nullptr;
  double *d_sc = nullptr;
  int rc = QR_OK;
  if (cudaMalloc((void **) &d_in, slice * F * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void **) &d_part, slice * s->ntrees * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void **) &d_sc, slice * sizeof(double)) != cudaSuccess) {
    set_error("qr_score_partial: out of device memory");
    rc = QR_ECUDA;
  }
  for (size_t n0 = 0; n0 < N && rc == QR_OK; n0 += slice) {
    const size_t n = std::min(slice, N - n0);
    cudaError_t e = cudaMemcpyAsync(d_in, docs + n0 * F, n * F * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
      rc = score_device_impl(s, d_in, n, F, d_sc, d_part);
      if (rc != QR_OK) break;
      e = cudaMemcpyAsync(partial + n0 * s->ntrees, d_part, n * s->ntrees * sizeof(float), cudaMemcpyDeviceToHost, s->stream);
    }
    if (e == cudaSuccess && scores) e = cudaMemcpyAsync(scores + n0, d_sc, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { set_error("qr_score_partial failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  }
  cudaStreamSynchronize(s->stream);
  if (d_in) cudaFree(d_in);
  if (d_part) cudaFree(d_part);
  if (d_sc) cudaFree(d_sc);
  return rc;
}

int qr_scorer_sync(qr_scorer *s) {
  if (!s) return QR_OK;
  cudaSetDevice(s->device);
  QR_CUDA(cudaStreamSynchronize(s->stream));
  return QR_OK;
}

uint64_t qr_scorer_launch_count(qr_scorer *s) { return s ? s->launches : 0; }

int qr_scorer_timer(qr_scorer *s, int stop, double *ms) {
  static thread_local cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!s) { set_error("null scorer"); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (!e0) { QR_CUDA(cudaEventCreate(&e0)); QR_CUDA(cudaEventCreate(&e1)); }
  if (!stop) { QR_CUDA(cudaEventRecord(e0, s->stream)); return QR_OK; }
  QR_CUDA(cudaEventRecord(e1, s->stream));
  QR_CUDA(cudaEventSynchronize(e1));
  float f = 0;
  QR_CUDA(cudaEventElapsedTime(&f, e0, e1));
  if (ms) *ms = f;
  return QR_OK;
}

// Host buffers (the reference's Dataset::data_): the documents stream through the device in slices,
// the upload of slice k+1 (copy stream) overlapping the encode + walk of slice k (compute stream).
int qr_score_dataset(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!s || !docs || !scores) { set_error("qr_score_dataset: null argument"); return QR_EINVAL; }
  if (F != s->F) { set_error("dataset has %zu features, the model was built for %zu", F, s->F); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  // slice: ~1/8 of the dataset, between 16 Ki documents and 256 MB of floats
  const size_t max_docs = std::max<size_t>(1, ((size_t) 256 << 20) / (F * sizeof(float)));
  const size_t slice = std::min(N, std::min(max_docs, std::max<size_t>(16384, (N + 7) / 8)));
  if (s->io_docs < slice) {
    for (int b = 0; b < 2; ++b) {
      if (s->d_in[b]) cudaFree(s->d_in[b]);
      if (s->d_out[b]) cudaFree(s->d_out[b]);
      s->d_in[b] = nullptr;
      s->d_out[b] = nullptr;
    }
    s->io_docs = 0;
    for (int b = 0; b < 2; ++b) {
      QR_CUDA(cudaMalloc((void **) &s->d_in[b], slice * F * sizeof(float)));
      QR_CUDA(cudaMalloc((void **) &s->d_out[b], slice * sizeof(double)));
    }
    s->io_docs = slice;
  }
  if (!s->copy_stream) {
    QR_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
      QR_CUDA(cudaEventCreateWithFlags(&s->ev_in[b], cudaEventDisableTiming));
      QR_CUDA(cudaEventCreateWithFlags(&s->ev_free[b], cudaEventDisableTiming));
    }
  }
  int rc = QR_OK;
  size_t k = 0;
  for (size_t n0 = 0; n0 < N && rc == QR_OK; n0 += slice, ++k) {
    const size_t n = std::min(slice, N - n0);
    const int b = (int) (k & 1);
    // the copy stream may overwrite d_in[b] only after the walk of slice k-2 has consumed it
    if (k >= 2) QR_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_free[b], 0));
    cudaError_t e = cudaMemcpyAsync(s->d_in[b], docs + n0 * F, n * F * sizeof(float), cudaMemcpyHostToDevice, s->copy_stream);
    if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; break; }
    QR_CUDA(cudaEventRecord(s->ev_in[b], s->copy_stream));
    QR_CUDA(cudaStreamWaitEvent(s->stream, s->ev_in[b], 0));
    rc = qr_score_dataset_device(s, s->d_in[b], n, F, s->d_out[b]);
    if (rc != QR_OK) break;
    QR_CUDA(cudaEventRecord(s->ev_free[b], s->stream));
    e = cudaMemcpyAsync(scores + n0, s->d_out[b], n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  }
  cudaError_t e = cudaStreamSynchronize(s->copy_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  if (e != cudaSuccess && rc == QR_OK) { set_error("scoring failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  return rc;
}

int qr_score_document(qr_scorer *s, const float *doc, size_t F, double *score) {
  return qr_score_dataset(s, doc, 1, F, score);
}

}  // extern "C"
