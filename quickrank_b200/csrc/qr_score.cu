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
// A 700-feature document shrinks from 2.8 KB of floats to 700 bytes, so a tile of 256 documents
// lives in shared memory next to a chunk of the ensemble; every thread owns one document, walks
// kIlp trees at a time (independent dependent-load chains) and adds their leaves in tree order.
#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "qr_internal.cuh"

namespace qr {

// 8-byte node: internal {slot, code, left, right} (children relative to the tree's first node);
// leaf {0xFFFF, 0, leaf index (relative to the tree's first leaf), 0}
struct __align__(8) CodeNode {
  uint16_t slot, code, left, right;
};

constexpr int kTpd = 4;                 // threads per document: each walks every kTpd-th tree of a chunk
constexpr int kChunkTrees = 16;         // trees per chunk (<= kTpd * kWalks)
constexpr int kWalks = kChunkTrees / kTpd;   // concurrent walks per thread (independent load chains)
constexpr int kMaxDocs = 256;           // documents per block (kMaxDocs * kTpd = 1024 threads)

// One chunk of the ensemble as a single contiguous blob, fetched with ONE bulk asynchronous copy
// (cp.async.bulk -> UBLKCP) into one of two shared-memory buffers while the other is being walked.
struct ChunkHeader {
  uint32_t ntrees, nodes_off, leaves_off, pad;      // byte offsets inside the blob
  uint16_t node_first[kChunkTrees + 1];             // first node of each tree (index into the blob's nodes)
  uint16_t leaf_first[kChunkTrees + 1];             // first leaf of each tree
  uint32_t pad2[3];
  double weight[kChunkTrees];
};
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

// code(x) for every (document, used feature): binary search in the feature's sorted thresholds
template <typename CodeT>
__global__ void encode_kernel(const float *__restrict__ docs, size_t n0, size_t n, size_t F,
                              const uint32_t *__restrict__ slot_feature, const uint32_t *__restrict__ thr_off,
                              const float *__restrict__ thr, uint32_t nslots, uint32_t stride, CodeT *codes) {
  const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nslots) return;
  const size_t d = idx / nslots;
  const uint32_t s = (uint32_t) (idx % nslots);
  const float x = docs[(n0 + d) * F + slot_feature[s]];
  const float *t = thr + thr_off[s];
  uint32_t lo = 0, hi = thr_off[s + 1] - thr_off[s];   // first k with t[k] >= x  ==  #thresholds < x
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (t[mid] >= x) hi = mid; else lo = mid + 1;
  }
  codes[d * stride + s] = (CodeT) lo;
}

// TILE: the block's documents are staged in shared memory (rows padded to an odd number of 32-bit
// words so that the lanes of a warp, which read different documents, hit different banks).
template <typename CodeT, bool TILE>
__global__ void __launch_bounds__(kMaxDocs * kTpd)
score_codes_kernel(const CodeT *__restrict__ codes, size_t n, uint32_t stride, uint32_t tile_words,
                   const unsigned char *__restrict__ chunks, uint32_t chunk_bytes, uint32_t nchunks,
                   uint32_t docs_per_block, double *__restrict__ scores) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];
  unsigned char *s_buf[2] = {smem_raw, smem_raw + chunk_bytes};
  uint32_t *s_tile = reinterpret_cast<uint32_t *>(smem_raw + 2 * (size_t) chunk_bytes);

  const uint32_t sub = threadIdx.x % kTpd;                 // which trees of a chunk this thread walks
  const uint32_t ldoc = threadIdx.x / kTpd;                // document within the block
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
    bulk_copy_g2s(s_buf[0], chunks, chunk_bytes, &s_bar[0]);
  }
  const CodeT *my;
  if (TILE) {
    // word-wise copy into rows of tile_words (odd) 32-bit words
    const uint32_t row_words = stride * (uint32_t) sizeof(CodeT) / 4;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(codes + d0 * stride);
    for (uint32_t i = threadIdx.x; i < ndocs * row_words; i += blockDim.x)
      s_tile[(i / row_words) * tile_words + (i % row_words)] = src[i];
    my = reinterpret_cast<const CodeT *>(s_tile + (size_t) ldoc * tile_words);
    __syncthreads();
  } else {
    my = codes + (d < n ? d : 0) * stride;
  }
  const bool active = d < n;
  double sum = 0.0;
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t b = c & 1u;
    // prefetch the next chunk into the other buffer (its previous contents were consumed before the
    // __syncthreads at the end of the previous iteration)
    if (threadIdx.x == 0 && c + 1 < nchunks) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&s_bar[b ^ 1u], chunk_bytes);
      bulk_copy_g2s(s_buf[b ^ 1u], chunks + (size_t) (c + 1) * chunk_bytes, chunk_bytes, &s_bar[b ^ 1u]);
    }
    mbar_wait(&s_bar[b], (c >> 1) & 1u);
    const ChunkHeader *h = reinterpret_cast<const ChunkHeader *>(s_buf[b]);
    const CodeNode *nodes = reinterpret_cast<const CodeNode *>(s_buf[b] + h->nodes_off);
    const double *leaves = reinterpret_cast<const double *>(s_buf[b] + h->leaves_off);
    const uint32_t nt = h->ntrees;
    // this thread's walks: trees sub, sub + kTpd, ...
    double val[kWalks];
    {
      uint32_t base[kWalks];
      CodeNode nd[kWalks];
      bool live[kWalks];
#pragma unroll
      for (int k = 0; k < kWalks; ++k) {
        const uint32_t t = sub + k * kTpd;
        live[k] = active && t < nt;
        base[k] = live[k] ? h->node_first[t] : 0u;
        nd[k] = nodes[base[k]];
        if (!live[k]) nd[k].slot = 0xFFFFu;
      }
      bool any = true;
      while (any) {      // the walks advance together; finished ones idle on their leaf
        any = false;
#pragma unroll
        for (int k = 0; k < kWalks; ++k) {
          if (nd[k].slot != 0xFFFFu) {
            const uint32_t code = my[nd[k].slot];
            nd[k] = nodes[base[k] + (code <= nd[k].code ? nd[k].left : nd[k].right)];   // rtnode.h:141-144
            any = true;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kWalks; ++k) {
        const uint32_t t = sub + k * kTpd;
        val[k] = live[k] ? __dmul_rn(leaves[h->leaf_first[t] + nd[k].left], h->weight[t]) : 0.0;   // ensemble.cc:116
      }
    }
    // ordered accumulation: the kTpd lanes of a document add the chunk's products in tree order
    // (each lane performs every addition, so all of them hold the same running sum)
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gbase = lane & ~(uint32_t) (kTpd - 1);
#pragma unroll
    for (int k = 0; k < kWalks; ++k) {
#pragma unroll
      for (int j = 0; j < kTpd; ++j) {
        const double v = __shfl_sync(0xffffffffu, val[k], gbase + j);
        if ((uint32_t) (k * kTpd + j) < nt) sum = __dadd_rn(sum, v);
      }
    }
    __syncthreads();   // every thread is done with buffer b before it is refilled
  }
  if (active && sub == 0) scores[d] = sum;
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
};

using namespace qr;

extern "C" {

int qr_scorer_create(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F, int device,
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

  // chunk blobs: up to kChunkTrees whole trees each, at most ~16 KB of nodes + leaves (a single
  // larger tree gets a chunk of its own); every blob is padded to the size of the largest one
  const size_t soft_payload = std::max<size_t>(16 * 1024, max_tree_nodes * 16);
  struct Chunk { size_t t0, t1, nodes, leaves; };
  std::vector<Chunk> plan;
  {
    Chunk cur{0, 0, 0, 0};
    for (size_t t = 0; t < ntrees; ++t) {
      size_t nl = 0;
      for (uint32_t i = 0; i < trees[t].nnodes; ++i) nl += trees[t].feature[i] < 0;
      const size_t nn = trees[t].nnodes;
      if (cur.t1 > cur.t0 && (cur.t1 - cur.t0 >= (size_t) kChunkTrees || (cur.nodes + nn) * 8 + (cur.leaves + nl) * 8 > soft_payload ||
                              cur.nodes + nn > 65535)) {
        plan.push_back(cur);
        cur = Chunk{t, t, 0, 0};
      }
      cur.t1 = t + 1;
      cur.nodes += nn;
      cur.leaves += nl;
    }
    if (cur.t1 > cur.t0) plan.push_back(cur);
  }
  size_t chunk_bytes = sizeof(ChunkHeader);
  for (auto &c : plan) chunk_bytes = std::max(chunk_bytes, sizeof(ChunkHeader) + ((c.nodes * 8 + 15) & ~(size_t) 15) + c.leaves * 8);
  chunk_bytes = (chunk_bytes + 127) & ~(size_t) 127;
  std::vector<unsigned char> blob(plan.size() * chunk_bytes, 0);
  for (size_t ci = 0; ci < plan.size(); ++ci) {
    const Chunk &c = plan[ci];
    unsigned char *base = blob.data() + ci * chunk_bytes;
    ChunkHeader *h = reinterpret_cast<ChunkHeader *>(base);
    h->ntrees = (uint32_t) (c.t1 - c.t0);
    h->nodes_off = (uint32_t) sizeof(ChunkHeader);
    h->leaves_off = (uint32_t) (sizeof(ChunkHeader) + ((c.nodes * 8 + 15) & ~(size_t) 15));
    CodeNode *nodes = reinterpret_cast<CodeNode *>(base + h->nodes_off);
    double *leaves = reinterpret_cast<double *>(base + h->leaves_off);
    size_t no = 0, lo = 0;
    for (size_t t = c.t0; t < c.t1; ++t) {
      const qr_flat_tree &ft = trees[t];
      const size_t k = t - c.t0;
      h->node_first[k] = (uint16_t) no;
      h->leaf_first[k] = (uint16_t) lo;
      h->weight[k] = weights[t];
      std::vector<uint32_t> leaf_index(ft.nnodes, 0);
      uint32_t nl = 0;
      for (uint32_t i = 0; i < ft.nnodes; ++i)
        if (ft.feature[i] < 0) leaf_index[i] = nl++;
      for (uint32_t i = 0; i < ft.nnodes; ++i) {
        CodeNode cn;
        if (ft.feature[i] >= 0) {
          const uint32_t sidx = slot_of[(uint32_t) ft.feature[i]];
          const float *tb = thr_flat.data() + thr_off[sidx], *te = thr_flat.data() + thr_off[sidx + 1];
          cn.slot = (uint16_t) sidx;
          cn.code = (uint16_t) (std::lower_bound(tb, te, ft.threshold[i]) - tb);   // rank of this threshold
          cn.left = (uint16_t) ft.left[i];
          cn.right = (uint16_t) ft.right[i];
        } else {
          cn.slot = 0xFFFFu; cn.code = 0; cn.left = (uint16_t) leaf_index[i]; cn.right = 0;
          leaves[lo + leaf_index[i]] = ft.value[i];
        }
        nodes[no + i] = cn;
      }
      no += ft.nnodes;
      lo += nl;
    }
    h->node_first[c.t1 - c.t0] = (uint16_t) no;
    h->leaf_first[c.t1 - c.t0] = (uint16_t) lo;
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

int qr_scorer_destroy(qr_scorer *s) {
  if (!s) return QR_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  void *ptrs[] = {s->d_chunks, s->d_slot_feature, s->d_thr_off, s->d_thr, s->d_codes};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return QR_OK;
}

// shared-memory row length in 32-bit words: the global row, padded to an odd word count
static uint32_t score_tile_words(const qr_scorer *s) {
  const uint32_t w = s->stride * (uint32_t) s->code_bytes / 4;
  return w | 1u;
}

// documents per block: as many (<= kMaxDocs, multiple of 8) as fit next to the two chunk buffers;
// 0 = not even 32 documents fit, the kernel then reads the codes from global memory
static uint32_t score_docs_per_block(const qr_scorer *s) {
  const size_t budget = 220 * 1024, fixed = 2 * (size_t) s->chunk_bytes, row = (size_t) score_tile_words(s) * 4;
  if (budget <= fixed) return 0;
  const size_t docs = (budget - fixed) / row / 8 * 8;
  return docs >= 32 ? (uint32_t) std::min<size_t>(docs, kMaxDocs) : 0;
}

int qr_score_dataset_device(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!s || !docs || !scores) { set_error("qr_score_dataset_device: null argument"); return QR_EINVAL; }
  if (F != s->F) { set_error("dataset has %zu features, the model was built for %zu", F, s->F); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  const size_t doc_bytes = (size_t) s->stride * s->code_bytes;
  const size_t batch = std::min<size_t>(N, std::max<size_t>(kMaxDocs, ((size_t) 1 << 30) / doc_bytes / kMaxDocs * kMaxDocs));
  if (s->batch_docs < batch) {
    if (s->d_codes) cudaFree(s->d_codes);
    s->d_codes = nullptr;
    QR_CUDA(cudaMalloc(&s->d_codes, batch * doc_bytes));
    QR_CUDA(cudaMemsetAsync(s->d_codes, 0, batch * doc_bytes, s->stream));
    s->batch_docs = batch;
  }
  const uint32_t tile_docs = score_docs_per_block(s);
  const bool tile = tile_docs != 0;
  const uint32_t dpb = tile ? tile_docs : 64;
  const uint32_t tile_words = score_tile_words(s);
  const size_t smem = 2 * (size_t) s->chunk_bytes + (tile ? (size_t) dpb * tile_words * 4 : 0);
  if (smem > 224 * 1024) { set_error("a chunk of the ensemble (%u bytes) does not fit in shared memory", s->chunk_bytes); return QR_ELIMIT; }
  for (size_t n0 = 0; n0 < N; n0 += batch) {
    const size_t n = std::min(batch, N - n0);
    const size_t work = n * std::max<uint32_t>(s->nslots, 1);
    const unsigned eg = (unsigned) ((work + 255) / 256);
    const unsigned sg = (unsigned) ((n + dpb - 1) / dpb);
#define QR_SCORE_LAUNCH(T, TILE)                                                                              \
  do {                                                                                                        \
    if (s->nslots)                                                                                            \
      encode_kernel<T><<<eg, 256, 0, s->stream>>>(docs, n0, n, F, s->d_slot_feature, s->d_thr_off, s->d_thr,  \
                                                   s->nslots, s->stride, (T *) s->d_codes);                  \
    cudaFuncSetAttribute(score_codes_kernel<T, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
    score_codes_kernel<T, TILE><<<sg, dpb * kTpd, smem, s->stream>>>(                                         \
        (const T *) s->d_codes, n, s->stride, tile_words, s->d_chunks, s->chunk_bytes, s->nchunks, dpb,       \
        scores + n0);                                                                                         \
    s->launches += 2;                                                                                         \
  } while (0)
    if (s->code_bytes == 1) { if (tile) QR_SCORE_LAUNCH(uint8_t, true); else QR_SCORE_LAUNCH(uint8_t, false); }
    else { if (tile) QR_SCORE_LAUNCH(uint16_t, true); else QR_SCORE_LAUNCH(uint16_t, false); }
#undef QR_SCORE_LAUNCH
    QR_CUDA(cudaGetLastError());
  }
  return QR_OK;
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

int qr_score_dataset(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!s || !docs || !scores) { set_error("qr_score_dataset: null argument"); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  // host buffers: stream the documents through the device in batches
  const size_t batch = std::min<size_t>(N, std::max<size_t>(1, ((size_t) 1 << 30) / (F * sizeof(float))));
  float *d_docs = nullptr;
  double *d_scores = nullptr;
  QR_CUDA(cudaMalloc((void **) &d_docs, batch * F * sizeof(float)));
  QR_CUDA(cudaMalloc((void **) &d_scores, batch * sizeof(double)));
  int rc = QR_OK;
  for (size_t n0 = 0; n0 < N && rc == QR_OK; n0 += batch) {
    const size_t n = std::min(batch, N - n0);
    cudaError_t e = cudaMemcpyAsync(d_docs, docs + n0 * F, n * F * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    if (e != cudaSuccess) { set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; break; }
    rc = qr_score_dataset_device(s, d_docs, n, F, d_scores);
    if (rc != QR_OK) break;
    e = cudaMemcpyAsync(scores + n0, d_scores, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  }
  cudaFree(d_docs);
  cudaFree(d_scores);
  return rc;
}

int qr_score_document(qr_scorer *s, const float *doc, size_t F, double *score) {
  return qr_score_dataset(s, doc, 1, F, score);
}

}  // extern "C"
