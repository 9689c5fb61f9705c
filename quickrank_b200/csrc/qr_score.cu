// quickrank_b200 — ensemble scoring (the quickscore path).
//
// Replaces LTR_Algorithm::score_dataset (ltr_algorithm.cc:44-52) -> Ensemble::score_instance
// (ensemble.cc:111-118) -> RTNode::score_instance (rtnode.h:134-152) of the reference: for every
// row-major document, sum over trees (in tree order, FP64, product and sum rounded separately as in
// the oracle build) of weight_t * leaf_t(doc), the node test being `x[feature] <= threshold` on floats.
#include <algorithm>
#include <cstring>
#include <vector>

#include "qr_internal.cuh"

namespace qr {

// One 16-byte node: internal {feature, threshold, left, right} (children as absolute node indices);
// leaf {-1, 0, value bits lo, value bits hi}.
struct __align__(16) PackedNode {
  int32_t feature;
  float threshold;
  int32_t a, b;
};

__global__ void __launch_bounds__(256)
score_kernel(const float *__restrict__ docs, size_t N, size_t F, const PackedNode *__restrict__ nodes,
             const uint32_t *__restrict__ roots, const double *__restrict__ weights, uint32_t ntrees,
             double *__restrict__ scores) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float *d = docs + i * F;
  double sum = 0.0;
  for (uint32_t t = 0; t < ntrees; ++t) {
    uint32_t nd = roots[t];
    PackedNode pn = nodes[nd];
    while (pn.feature >= 0) {
      nd = d[pn.feature] <= pn.threshold ? (uint32_t) pn.a : (uint32_t) pn.b;   // rtnode.h:141-144
      pn = nodes[nd];
    }
    const double leaf = __hiloint2double(pn.b, pn.a);
    sum = __dadd_rn(sum, __dmul_rn(leaf, weights[t]));                           // ensemble.cc:116
  }
  scores[i] = sum;
}

}  // namespace qr

struct qr_scorer {
  int device = 0;
  size_t ntrees = 0, F = 0;
  qr::PackedNode *d_nodes = nullptr;
  uint32_t *d_roots = nullptr;
  double *d_weights = nullptr;
  cudaStream_t stream = nullptr;
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
  std::vector<PackedNode> nodes;
  std::vector<uint32_t> roots(ntrees);
  for (size_t t = 0; t < ntrees; ++t) {
    const qr_flat_tree &ft = trees[t];
    if (ft.nnodes == 0) { set_error("tree %zu is empty", t); return QR_EINVAL; }
    const uint32_t base = (uint32_t) nodes.size();
    roots[t] = base;
    for (uint32_t i = 0; i < ft.nnodes; ++i) {
      PackedNode pn;
      if (ft.feature[i] >= 0) {
        if ((size_t) ft.feature[i] >= F || ft.left[i] < 0 || ft.right[i] < 0 ||
            (uint32_t) ft.left[i] >= ft.nnodes || (uint32_t) ft.right[i] >= ft.nnodes) {
          set_error("tree %zu node %u is malformed", t, i);
          return QR_EINVAL;
        }
        pn.feature = ft.feature[i];
        pn.threshold = ft.threshold[i];
        pn.a = (int32_t) (base + ft.left[i]);
        pn.b = (int32_t) (base + ft.right[i]);
      } else {
        pn.feature = -1;
        pn.threshold = 0.f;
        long long bits;
        memcpy(&bits, &ft.value[i], 8);
        pn.a = (int32_t) (bits & 0xffffffffll);
        pn.b = (int32_t) (bits >> 32);
      }
      nodes.push_back(pn);
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
  *out = s;
  QR_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  QR_CUDA(cudaMalloc((void **) &s->d_nodes, std::max<size_t>(nodes.size(), 1) * sizeof(PackedNode)));
  QR_CUDA(cudaMalloc((void **) &s->d_roots, std::max<size_t>(ntrees, 1) * sizeof(uint32_t)));
  QR_CUDA(cudaMalloc((void **) &s->d_weights, std::max<size_t>(ntrees, 1) * sizeof(double)));
  QR_CUDA(cudaMemcpy(s->d_nodes, nodes.data(), nodes.size() * sizeof(PackedNode), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(s->d_roots, roots.data(), ntrees * sizeof(uint32_t), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(s->d_weights, weights, ntrees * sizeof(double), cudaMemcpyHostToDevice));
  return QR_OK;
}

int qr_scorer_destroy(qr_scorer *s) {
  if (!s) return QR_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->d_nodes) cudaFree(s->d_nodes);
  if (s->d_roots) cudaFree(s->d_roots);
  if (s->d_weights) cudaFree(s->d_weights);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return QR_OK;
}

int qr_score_dataset_device(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!s || !docs || !scores) { set_error("qr_score_dataset_device: null argument"); return QR_EINVAL; }
  if (F != s->F) { set_error("dataset has %zu features, the model was built for %zu", F, s->F); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  score_kernel<<<(unsigned) ((N + 255) / 256), 256, 0, s->stream>>>(docs, N, F, s->d_nodes, s->d_roots, s->d_weights,
                                                                 (uint32_t) s->ntrees, scores);
  QR_CUDA(cudaGetLastError());
  return QR_OK;
}

int qr_scorer_sync(qr_scorer *s) {
  if (!s) return QR_OK;
  cudaSetDevice(s->device);
  QR_CUDA(cudaStreamSynchronize(s->stream));
  return QR_OK;
}

int qr_score_dataset(qr_scorer *s, const float *docs, size_t N, size_t F, double *scores) {
  if (!s || !docs || !scores) { set_error("qr_score_dataset: null argument"); return QR_EINVAL; }
  cudaSetDevice(s->device);
  if (N == 0) return QR_OK;
  float *d_docs = nullptr;
  double *d_scores = nullptr;
  QR_CUDA(cudaMalloc((void **) &d_docs, N * F * sizeof(float)));
  QR_CUDA(cudaMalloc((void **) &d_scores, N * sizeof(double)));
  QR_CUDA(cudaMemcpyAsync(d_docs, docs, N * F * sizeof(float), cudaMemcpyHostToDevice, s->stream));
  int rc = qr_score_dataset_device(s, d_docs, N, F, d_scores);
  if (rc == QR_OK) {
    QR_CUDA(cudaMemcpyAsync(scores, d_scores, N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    QR_CUDA(cudaStreamSynchronize(s->stream));
  }
  cudaFree(d_docs);
  cudaFree(d_scores);
  return rc;
}

int qr_score_document(qr_scorer *s, const float *doc, size_t F, double *score) {
  return qr_score_dataset(s, doc, 1, F, score);
}

}  // extern "C"
