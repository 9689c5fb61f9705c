// quickrank_b200 — line search over a row-major score matrix (the per-tree partial scores CLEAVER optimises,
// driver.cc:411-446): the passes over documents of LineSearch::learn (src/learning/linear/line_search.cc:153-416) —
// a weighted sum per document for every candidate weight vector, and NDCG@k of each resulting ranking — on the GPU.
// The host keeps the search itself (window, points, acceptance rules): quickrank_b200/linesearch.py.
//
// Arithmetic follows the reference's Release build statement by statement (checked in the disassembly of
// oracle/_ref/obj/learning/linear/line_search.o): LineSearch::score and preCompute add w[f]*x[s][f] with a separate
// multiply and add, in feature order, from 0; pre_sum = total - w[f]*x (multiply, subtract); the step-1 candidates are
// fma(point, x, pre_sum); the step-2 candidates accumulate fma(fma(step[f], p, w[f]), x, score).  The library is
// compiled with -fmad=false, so only the fma() written below are fused.  Rankings and NDCG come from the training
// context's own kernels (rank_kernel: libstdc++ introsort replica for tied scores; sequential mean over queries,
// metric.h:77-92), through a one-feature REFERENCE-mode context that holds the labels and query boundaries.
#include <algorithm>
#include <cstring>
#include <vector>

#include "qr_internal.cuh"

namespace qr {

constexpr uint32_t kLsWarps = 8;
constexpr uint32_t kLsMaxPoints = 32;

// A warp takes 32 consecutive documents and walks the features 32 at a time: the tile is loaded row by row (each
// row 128 contiguous bytes of one document) and read back column-wise, lane d following document d.
struct LsTile { float v[32][33]; };

__device__ __forceinline__ void ls_load_tile(LsTile &t, const float *__restrict__ x, size_t N, uint32_t T, size_t doc0,
                                             uint32_t f0, uint32_t lane) {
#pragma unroll 8
  for (uint32_t r = 0; r < 32; ++r) {
    const size_t d = doc0 + r;
    t.v[r][lane] = (d < N && f0 + lane < T) ? x[d * T + f0 + lane] : 0.f;
  }
}

// out[s] = sum_f w[f] * x[s][f]: separate multiply and add, feature order, from 0 (line_search.cc:447-482)
__global__ void __launch_bounds__(kLsWarps * 32)
ls_total_kernel(const float *__restrict__ x, size_t N, uint32_t T, const double *__restrict__ w, double *out) {
  __shared__ LsTile tiles[kLsWarps];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const size_t doc0 = ((size_t) blockIdx.x * kLsWarps + warp) * 32u;
  if (doc0 >= N) return;
  LsTile &t = tiles[warp];
  double acc = 0.0;
  for (uint32_t f0 = 0; f0 < T; f0 += 32) {
    __syncwarp();
    ls_load_tile(t, x, N, T, doc0, f0, lane);
    __syncwarp();
    const uint32_t nf = min(32u, T - f0);
    for (uint32_t j = 0; j < nf; ++j) acc = __dadd_rn(acc, __dmul_rn(w[f0 + j], (double) t.v[lane][j]));
  }
  if (doc0 + lane < N) out[doc0 + lane] = acc;
}

// step 1 (line_search.cc:252-272): scores[p][s] = fma(points[p], x[s][f], total[s] - w_f * x[s][f])
__global__ void ls_feature_points_kernel(const float *__restrict__ x, size_t N, uint32_t T, uint32_t f, double w_f,
                                         const double *__restrict__ total, const double *__restrict__ points,
                                         uint32_t np, double *scores) {
  const size_t s = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const double xv = (double) x[s * T + f];
  const double pre = __dsub_rn(total[s], __dmul_rn(w_f, xv));
  for (uint32_t p = 0; p < np; ++p) scores[(size_t) p * N + s] = fma(points[p], xv, pre);
}

// step 2 (line_search.cc:303-316): scores[p][s] = sum_f fma(fma(step[f], p, w[f]), x[s][f], .), p = p0 .. p0 + np - 1
__global__ void __launch_bounds__(kLsWarps * 32)
ls_line_points_kernel(const float *__restrict__ x, size_t N, uint32_t T, const double *__restrict__ w,
                      const double *__restrict__ step, uint32_t p0, uint32_t np, double *scores) {
  __shared__ LsTile tiles[kLsWarps];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const size_t doc0 = ((size_t) blockIdx.x * kLsWarps + warp) * 32u;
  if (doc0 >= N) return;
  LsTile &t = tiles[warp];
  double acc[kLsMaxPoints];
#pragma unroll
  for (uint32_t p = 0; p < kLsMaxPoints; ++p) acc[p] = 0.0;
  for (uint32_t f0 = 0; f0 < T; f0 += 32) {
    __syncwarp();
    ls_load_tile(t, x, N, T, doc0, f0, lane);
    __syncwarp();
    const uint32_t nf = min(32u, T - f0);
    for (uint32_t j = 0; j < nf; ++j) {
      const double xv = (double) t.v[lane][j];
      const double wf = w[f0 + j], sf = step[f0 + j];
#pragma unroll
      for (uint32_t p = 0; p < kLsMaxPoints; ++p)
        if (p < np) acc[p] = fma(fma(sf, (double) (p0 + p), wf), xv, acc[p]);
    }
  }
  if (doc0 + lane < N) {
#pragma unroll
    for (uint32_t p = 0; p < kLsMaxPoints; ++p)
      if (p < np) scores[(size_t) p * N + doc0 + lane] = acc[p];
  }
}

// CLEAVER's quality-loss passes (quality_loss_pruning.cc:59-70, quality_loss_adv_pruning.cc:60-81): candidate c is
// the ensemble without column cols[c]: scores[c][s] = total[s] - w[col] * x[s][col] (multiply, then subtract)
__global__ void ls_drop_points_kernel(const float *__restrict__ x, size_t N, uint32_t T, const uint32_t *__restrict__ cols,
                                      const double *__restrict__ w, const double *__restrict__ total, uint32_t nc,
                                      double *scores) {
  const size_t s = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= N) return;
  const double tot = total[s];
  const float *row = x + s * T;
  for (uint32_t c = 0; c < nc; ++c) {
    const uint32_t f = cols[c];
    scores[(size_t) c * N + s] = __dsub_rn(tot, __dmul_rn(w[f], (double) row[f]));
  }
}

// the reference's running scores after a column is pruned (quality_loss_adv_pruning.cc:88-92): total -= w_f * x[.][f]
__global__ void ls_drop_column_kernel(const float *__restrict__ x, size_t N, uint32_t T, uint32_t f, double w_f, double *total) {
  const size_t s = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (s < N) total[s] = __dsub_rn(total[s], __dmul_rn(w_f, (double) x[s * T + f]));
}

// ScoreLossPruning (score_loss_pruning.cc:58-63): out[f] = sum over documents, in document order, of
// w[f] * x[s][f] / total[s] — one sequentially rounded FP64 chain per column.  Thread f owns column f (a warp reads 32
// consecutive columns of a row: coalesced); the quotients of a batch of documents are computed ahead of the chain of
// additions, which is the only dependent part.
constexpr uint32_t kSlBatch = 16;
__global__ void ls_score_loss_kernel(const float *__restrict__ x, size_t N, uint32_t T, const double *__restrict__ w,
                                     const double *__restrict__ total, double *out) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= T) return;
  const double wf = w[f];
  double acc = 0.0;
  size_t s = 0;
  for (; s + kSlBatch <= N; s += kSlBatch) {
    double q[kSlBatch];
#pragma unroll
    for (uint32_t j = 0; j < kSlBatch; ++j) q[j] = __ddiv_rn(__dmul_rn(wf, (double) x[(s + j) * T + f]), total[s + j]);
#pragma unroll
    for (uint32_t j = 0; j < kSlBatch; ++j) acc = __dadd_rn(acc, q[j]);
  }
  for (; s < N; ++s) acc = __dadd_rn(acc, __ddiv_rn(__dmul_rn(wf, (double) x[s * T + f]), total[s]));
  out[f] = acc;
}

}  // namespace qr

using namespace qr;

struct qr_linesearch {
  int device = 0;
  size_t N = 0, T = 0;
  qr_ctx *ctx = nullptr;         // labels, query boundaries, ranking and NDCG kernels (one dummy feature)
  float *d_x = nullptr;          // [N][T]
  double *d_w = nullptr, *d_step = nullptr, *d_points = nullptr;   // [T], [T], [kLsMaxPoints]
  double *d_total = nullptr;     // [N] weighted sum under `total_w`
  double *d_scores = nullptr;    // [kLsMaxPoints][N] candidate score vectors
  std::vector<double> total_w;   // weights d_total was computed for (empty: none)
};

static int ls_metric_of(qr_linesearch *ls, double *scores_dev, double *metric) {
  return evaluate_vectors(ls->ctx, scores_dev, 1, metric);
}

static int ls_upload(qr_linesearch *ls, double *dst, const double *src, size_t n) {
  QR_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, ls->ctx->stream));
  QR_CUDA(cudaStreamSynchronize(ls->ctx->stream));   // (pageable source)
  return QR_OK;
}

static int ls_total(qr_linesearch *ls, const double *weights) {
  if (ls->total_w.size() == ls->T && memcmp(ls->total_w.data(), weights, ls->T * sizeof(double)) == 0) return QR_OK;
  QR_TRY(ls_upload(ls, ls->d_w, weights, ls->T));
  const unsigned grid = (unsigned) ((ls->N + kLsWarps * 32 - 1) / (kLsWarps * 32));
  ls_total_kernel<<<grid, kLsWarps * 32, 0, ls->ctx->stream>>>(ls->d_x, ls->N, (uint32_t) ls->T, ls->d_w, ls->d_total);
  QR_CUDA(cudaGetLastError());
  ls->ctx->launches++;
  ls->total_w.assign(weights, weights + ls->T);
  return QR_OK;
}

extern "C" {

int qr_ls_create(const float *x_rowmajor, size_t N, size_t T, const float *labels, const uint64_t *qoffsets, size_t Q,
                 uint32_t cutoff, int device, qr_linesearch **out) {
  if (!x_rowmajor || !labels || !qoffsets || !out || N == 0 || T == 0 || Q == 0) {
    set_error("qr_ls_create: bad arguments");
    return QR_EINVAL;
  }
  *out = nullptr;
  qr_linesearch *ls = new qr_linesearch();
  ls->N = N; ls->T = T;
  qr_params p;
  memset(&p, 0, sizeof(p));
  p.algo = QR_ALGO_LAMBDAMART;
  p.hist_mode = QR_HIST_REFERENCE;   // sequential mean over queries, as Metric::evaluate_dataset
  p.nleaves = 2; p.minleafsupport = 1; p.shrinkage = 1.0; p.ndcg_cutoff = cutoff; p.device = device;
  std::vector<float> dummy(N, 0.f);
  int rc = qr_ctx_create(dummy.data(), N, 1, labels, qoffsets, Q, &p, &ls->ctx);
  if (rc != QR_OK) { delete ls; return rc; }
  ls->device = ls->ctx->device;
  auto fail = [&](const char *what) {
    set_error("qr_ls_create: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
    qr_ls_destroy(ls);
    return QR_ECUDA;
  };
  if (cudaMalloc((void **) &ls->d_x, N * T * sizeof(float)) != cudaSuccess) return fail("score matrix");
  if (cudaMalloc((void **) &ls->d_w, T * sizeof(double)) != cudaSuccess) return fail("weights");
  if (cudaMalloc((void **) &ls->d_step, T * sizeof(double)) != cudaSuccess) return fail("steps");
  if (cudaMalloc((void **) &ls->d_points, kLsMaxPoints * sizeof(double)) != cudaSuccess) return fail("points");
  if (cudaMalloc((void **) &ls->d_total, N * sizeof(double)) != cudaSuccess) return fail("totals");
  if (cudaMalloc((void **) &ls->d_scores, (size_t) kLsMaxPoints * N * sizeof(double)) != cudaSuccess) return fail("candidate scores");
  if (cudaMemcpy(ls->d_x, x_rowmajor, N * T * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return fail("copy of the score matrix");
  *out = ls;
  return QR_OK;
}

int qr_ls_destroy(qr_linesearch *ls) {
  if (!ls) return QR_OK;
  if (ls->ctx) cudaSetDevice(ls->ctx->device);
  void *ptrs[] = {ls->d_x, ls->d_w, ls->d_step, ls->d_points, ls->d_total, ls->d_scores};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (ls->ctx) qr_ctx_destroy(ls->ctx);
  delete ls;
  return QR_OK;
}

int qr_ls_evaluate(qr_linesearch *ls, const double *weights, double *metric) {
  if (!ls || !weights || !metric) { set_error("qr_ls_evaluate: bad arguments"); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_total(ls, weights));
  return ls_metric_of(ls, ls->d_total, metric);
}

int qr_ls_feature_points(qr_linesearch *ls, const double *weights, uint32_t f, const double *points, uint32_t npoints,
                         double *metrics) {
  if (!ls || !weights || !points || !metrics || f >= ls->T) { set_error("qr_ls_feature_points: bad arguments"); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_total(ls, weights));
  for (uint32_t p0 = 0; p0 < npoints; p0 += kLsMaxPoints) {
    const uint32_t np = std::min(kLsMaxPoints, npoints - p0);
    QR_TRY(ls_upload(ls, ls->d_points, points + p0, np));
    ls_feature_points_kernel<<<(unsigned) ((ls->N + 255) / 256), 256, 0, ls->ctx->stream>>>(
        ls->d_x, ls->N, (uint32_t) ls->T, f, weights[f], ls->d_total, ls->d_points, np, ls->d_scores);
    QR_CUDA(cudaGetLastError());
    ls->ctx->launches++;
    QR_TRY(evaluate_vectors(ls->ctx, ls->d_scores, np, metrics + p0));
  }
  return QR_OK;
}

int qr_ls_line_points(qr_linesearch *ls, const double *weights, const double *step, uint32_t npoints, double *metrics) {
  if (!ls || !weights || !step || !metrics) { set_error("qr_ls_line_points: bad arguments"); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_upload(ls, ls->d_w, weights, ls->T));
  QR_TRY(ls_upload(ls, ls->d_step, step, ls->T));
  const unsigned grid = (unsigned) ((ls->N + kLsWarps * 32 - 1) / (kLsWarps * 32));
  for (uint32_t p0 = 0; p0 < npoints; p0 += kLsMaxPoints) {
    const uint32_t np = std::min(kLsMaxPoints, npoints - p0);
    ls_line_points_kernel<<<grid, kLsWarps * 32, 0, ls->ctx->stream>>>(ls->d_x, ls->N, (uint32_t) ls->T, ls->d_w, ls->d_step,
                                                                      p0, np, ls->d_scores);
    QR_CUDA(cudaGetLastError());
    ls->ctx->launches++;
    QR_TRY(evaluate_vectors(ls->ctx, ls->d_scores, np, metrics + p0));
  }
  return QR_OK;
}

int qr_ls_drop_points(qr_linesearch *ls, const double *weights, const uint32_t *cols, uint32_t ncols, double *metrics) {
  if (!ls || !weights || !cols || !metrics) { set_error("qr_ls_drop_points: bad arguments"); return QR_EINVAL; }
  for (uint32_t c = 0; c < ncols; ++c)
    if (cols[c] >= ls->T) { set_error("qr_ls_drop_points: column %u out of range", cols[c]); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_total(ls, weights));
  // (after qr_ls_drop_column the cached sums belong to weights the device copy does not hold any more)
  QR_TRY(ls_upload(ls, ls->d_w, weights, ls->T));
  uint32_t *d_cols = reinterpret_cast<uint32_t *>(ls->d_points);   // kLsMaxPoints doubles hold kLsMaxPoints indices
  for (uint32_t c0 = 0; c0 < ncols; c0 += kLsMaxPoints) {
    const uint32_t nc = std::min(kLsMaxPoints, ncols - c0);
    QR_CUDA(cudaMemcpyAsync(d_cols, cols + c0, nc * sizeof(uint32_t), cudaMemcpyHostToDevice, ls->ctx->stream));
    QR_CUDA(cudaStreamSynchronize(ls->ctx->stream));
    ls_drop_points_kernel<<<(unsigned) ((ls->N + 255) / 256), 256, 0, ls->ctx->stream>>>(
        ls->d_x, ls->N, (uint32_t) ls->T, d_cols, ls->d_w, ls->d_total, nc, ls->d_scores);
    QR_CUDA(cudaGetLastError());
    ls->ctx->launches++;
    QR_TRY(evaluate_vectors(ls->ctx, ls->d_scores, nc, metrics + c0));
  }
  return QR_OK;
}

int qr_ls_drop_column(qr_linesearch *ls, const double *weights, uint32_t f) {
  if (!ls || !weights || f >= ls->T) { set_error("qr_ls_drop_column: bad arguments"); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_total(ls, weights));
  ls_drop_column_kernel<<<(unsigned) ((ls->N + 255) / 256), 256, 0, ls->ctx->stream>>>(ls->d_x, ls->N, (uint32_t) ls->T, f,
                                                                                     weights[f], ls->d_total);
  QR_CUDA(cudaGetLastError());
  ls->ctx->launches++;
  ls->total_w[f] = 0.0;   // the sums now stand for these weights: the next call with them keeps the running sums
  return QR_OK;
}

int qr_ls_score_loss(qr_linesearch *ls, const double *weights, double *loss) {
  if (!ls || !weights || !loss) { set_error("qr_ls_score_loss: bad arguments"); return QR_EINVAL; }
  QR_CUDA(cudaSetDevice(ls->device));
  QR_TRY(ls_total(ls, weights));
  QR_TRY(ls_upload(ls, ls->d_w, weights, ls->T));
  ls_score_loss_kernel<<<(unsigned) ((ls->T + 127) / 128), 128, 0, ls->ctx->stream>>>(ls->d_x, ls->N, (uint32_t) ls->T, ls->d_w,
                                                                                    ls->d_total, ls->d_step);
  QR_CUDA(cudaGetLastError());
  ls->ctx->launches++;
  QR_CUDA(cudaMemcpyAsync(loss, ls->d_step, ls->T * sizeof(double), cudaMemcpyDeviceToHost, ls->ctx->stream));
  QR_CUDA(cudaStreamSynchronize(ls->ctx->stream));
  return QR_OK;
}

uint64_t qr_ls_launch_count(qr_linesearch *ls) { return ls && ls->ctx ? ls->ctx->launches : 0; }

}  // extern "C"
