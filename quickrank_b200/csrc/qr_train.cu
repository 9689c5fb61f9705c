// quickrank_b200 — host side of the training C ABI (include/quickrank_b200.h).
//
// The host keeps what the reference keeps on the host inside RegressionTree::fit (rt.cc:49-163):
// the max-heap of frontier nodes keyed by deviance (maxheap.h:31-106) and the node bookkeeping.
// All per-document and per-bin work runs in the kernels of qr_kernels.cuh.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <functional>
#include <limits>

#include <cub/cub.cuh>

#include "qr_comm.cuh"
#include "qr_kernels.cuh"

namespace qr {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

#define QR_LAUNCH(ctx, phase, kernel, grid, block, smem, ...)                          \
  do {                                                                                 \
    kernel<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);                         \
    (ctx)->launches++;                                                                 \
    (ctx)->phase_launches[phase]++;                                                    \
    cudaError_t _le = cudaGetLastError();                                              \
    if (_le != cudaSuccess) {                                                          \
      qr::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,     \
                    cudaGetErrorString(_le));                                          \
      return QR_ECUDA;                                                                 \
    }                                                                                  \
  } while (0)

struct PhaseTimer {
  qr_ctx *c;
  int phase;
  PhaseTimer(qr_ctx *ctx, int ph) : c(ctx), phase(ph) {
    if (c->profiling) cudaEventRecord(c->ev0, c->stream);
  }
  ~PhaseTimer() {
    if (c->profiling) {
      cudaEventRecord(c->ev1, c->stream);
      cudaEventSynchronize(c->ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, c->ev0, c->ev1);
      c->phase_ms[phase] += ms;
    }
  }
};

template <typename T>
static int dev_alloc(T **p, size_t count) {
  QR_CUDA(cudaMalloc((void **) p, std::max<size_t>(count, 1) * sizeof(T)));
  return QR_OK;
}

static int ceil_log2(size_t n) {
  int k = 0;
  while (((size_t) 1 << k) < n) ++k;
  return k;
}

// ------------------------------------------------------------------------------------------
// Init: thresholds + bin map
// ------------------------------------------------------------------------------------------
static int build_binning(qr_ctx *c, const float *d_col) {
  const size_t N = c->N, F = c->F;
  cudaStream_t st = c->stream;
  uint32_t *d_keys = nullptr, *d_keys_out = nullptr;
  float *d_vals = nullptr, *d_uniq = nullptr;
  uint8_t *d_flags = nullptr;
  int *d_num = nullptr, *d_bad = nullptr;
  void *d_tmp = nullptr;
  QR_TRY(dev_alloc(&d_keys, N));
  QR_TRY(dev_alloc(&d_keys_out, N));
  QR_TRY(dev_alloc(&d_vals, N));
  QR_TRY(dev_alloc(&d_uniq, N));
  QR_TRY(dev_alloc(&d_flags, N));
  QR_TRY(dev_alloc(&d_num, 1));
  QR_TRY(dev_alloc(&d_bad, 1));
  QR_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  size_t tmp_sort = 0, tmp_sel = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_sort, d_keys, d_keys_out, (int) N, 0, 32, st);
  cub::DeviceSelect::Flagged(nullptr, tmp_sel, d_vals, d_flags, d_uniq, d_num, (int) N, st);
  const size_t tmp_bytes = std::max(tmp_sort, tmp_sel);
  QR_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));

  c->thr.assign(F, std::vector<float>());
  const size_t nth = (size_t) c->p.nthresholds;
  const unsigned blocks = (unsigned) ((N + 255) / 256);
  uint32_t max_bin = 0;
  for (size_t f = 0; f < F; ++f) {
    const float *x = d_col + f * N;
    flip_keys_kernel<<<blocks, 256, 0, st>>>(x, d_keys, N, d_bad);
    size_t tb = tmp_bytes;
    QR_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tb, d_keys, d_keys_out, (int) N, 0, 32, st));
    distinct_flags_kernel<<<blocks, 256, 0, st>>>(d_keys_out, d_vals, d_flags, N);
    tb = tmp_bytes;
    QR_CUDA(cub::DeviceSelect::Flagged(d_tmp, tb, d_vals, d_flags, d_uniq, d_num, (int) N, st));
    int nu = 0;
    QR_CUDA(cudaMemcpyAsync(&nu, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
    QR_CUDA(cudaStreamSynchronize(st));
    std::vector<float> &t = c->thr[f];
    if (nth == 0 || (size_t) nu <= nth) {        // mart.cc:155-158: distinct values + FLT_MAX
      t.resize((size_t) nu + 1);
      QR_CUDA(cudaMemcpy(t.data(), d_uniq, (size_t) nu * sizeof(float), cudaMemcpyDeviceToHost));
      t[nu] = FLT_MAX;
      max_bin = std::max<uint32_t>(max_bin, (uint32_t) (nu - 1));
    } else {                                     // mart.cc:159-169: equal width, float accumulation
      float fmin = 0, fmax = 0;
      QR_CUDA(cudaMemcpy(&fmin, d_vals, sizeof(float), cudaMemcpyDeviceToHost));
      QR_CUDA(cudaMemcpy(&fmax, d_vals + (N - 1), sizeof(float), cudaMemcpyDeviceToHost));
      t.resize(nth + 1);
      float cur = fmin;
      const float step = (float) std::fabs((double) (fmax - cur)) / (float) nth;
      for (size_t j = 0; j != nth; cur += step) t[j++] = cur;
      t[nth] = FLT_MAX;
      max_bin = std::max<uint32_t>(max_bin, (uint32_t) nth);
    }
  }
  int bad = 0;
  QR_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_keys); cudaFree(d_keys_out); cudaFree(d_vals); cudaFree(d_uniq);
  cudaFree(d_flags); cudaFree(d_num); cudaFree(d_bad); cudaFree(d_tmp);
  if (bad) {
    set_error("feature matrix contains NaN or infinite values (the reference's bin map is undefined for them)");
    return QR_EINVAL;
  }
  if (max_bin > 65535u) {
    set_error("a feature has %u occupied bins; this build stores bins in at most 16 bits "
              "(use --num-thresholds)", max_bin + 1);
    return QR_ELIMIT;
  }
  c->bin_bytes = max_bin <= 255u ? 1 : 2;
  c->fpp = qr::kPanelBytes / c->bin_bytes;
  c->npanels = (uint32_t) ((F + c->fpp - 1) / c->fpp);
  c->thr_off.assign(F + 1, 0);
  c->max_thr = 0;
  for (size_t f = 0; f < F; ++f) {
    c->thr_off[f + 1] = c->thr_off[f] + (uint32_t) c->thr[f].size();
    c->max_thr = std::max<uint32_t>(c->max_thr, (uint32_t) c->thr[f].size());
  }
  c->ncells = c->thr_off[F];
  c->max_panel_cells = 0;
  for (uint32_t p = 0; p < c->npanels; ++p) {
    size_t f0 = (size_t) p * c->fpp, f1 = std::min(F, f0 + c->fpp);
    c->max_panel_cells = std::max(c->max_panel_cells, c->thr_off[f1] - c->thr_off[f0]);
  }
  // upload thresholds, build panels
  std::vector<float> flat(c->ncells);
  for (size_t f = 0; f < F; ++f) std::copy(c->thr[f].begin(), c->thr[f].end(), flat.begin() + c->thr_off[f]);
  float *d_thr = nullptr;
  QR_TRY(dev_alloc(&d_thr, c->ncells));
  QR_TRY(dev_alloc(&c->d_thr_off, F + 1));
  QR_CUDA(cudaMemcpy(d_thr, flat.data(), c->ncells * sizeof(float), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_thr_off, c->thr_off.data(), (F + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  QR_TRY(dev_alloc(&c->d_panels, (size_t) c->npanels * N));
  dim3 grid((unsigned) ((N + 127) / 128), c->npanels);
  if (c->bin_bytes == 1)
    binning_kernel<uint8_t><<<grid, 128, 0, st>>>(d_col, N, (uint32_t) F, d_thr, c->d_thr_off, c->d_panels, c->npanels);
  else
    binning_kernel<uint16_t><<<grid, 128, 0, st>>>(d_col, N, (uint32_t) F, d_thr, c->d_thr_off, c->d_panels, c->npanels);
  QR_CUDA(cudaGetLastError());
  QR_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_thr);
  return QR_OK;
}

static int ctx_create_common(const float *feat, bool rowmajor, size_t N, size_t F, const float *labels,
                             const uint64_t *qoffsets, size_t Q, const qr_params *params, qr_ctx **out) {
  if (!feat || !labels || !qoffsets || !params || !out || N == 0 || F == 0 || Q == 0) {
    set_error("qr_ctx_create: null or empty argument");
    return QR_EINVAL;
  }
  if (N >= ((size_t) 1 << 31)) { set_error("N >= 2^31 documents per GPU is not supported"); return QR_ELIMIT; }
  if (qoffsets[0] != 0 || qoffsets[Q] != N) { set_error("query offsets must start at 0 and end at N"); return QR_EINVAL; }
  if (params->algo > QR_ALGO_OBVLAMBDAMART) { set_error("unknown algo %u", params->algo); return QR_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return QR_ENODEVICE;
  }
  qr_ctx *c = new qr_ctx();
  c->p = *params;
  if (params->device >= 0) {
    if (params->device >= ndev) { delete c; set_error("device %d out of range", params->device); return QR_EINVAL; }
    cudaSetDevice(params->device);
  }
  cudaGetDevice(&c->device);
  c->N = N; c->F = F; c->Q = Q; c->N_global = N; c->Q_global = Q;
  c->cutoff = params->ndcg_cutoff == 0 ? std::numeric_limits<size_t>::max() : (size_t) params->ndcg_cutoff;  // metric.h:65-67
  c->lambda = params->algo == QR_ALGO_LAMBDAMART || params->algo == QR_ALGO_OBVLAMBDAMART;
  c->oblivious = params->algo == QR_ALGO_OBVMART || params->algo == QR_ALGO_OBVLAMBDAMART;
  c->exact = params->hist_mode == QR_HIST_REFERENCE;
  if (c->oblivious && (params->treedepth == 0 || params->treedepth > 16)) {
    delete c; set_error("treedepth must be in 1..16"); return QR_EINVAL;
  }
  if (!c->oblivious && params->nleaves < 1) { delete c; set_error("nleaves must be >= 1"); return QR_EINVAL; }
  *out = c;  // destroyed by the caller on failure paths below via qr_ctx_destroy
  QR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  QR_CUDA(cudaEventCreate(&c->ev0));
  QR_CUDA(cudaEventCreate(&c->ev1));
  QR_CUDA(cudaEventCreate(&c->ev_t0));
  QR_CUDA(cudaEventCreate(&c->ev_t1));

  // features -> device column-major (VerticalDataset layout), then bins; floats are released
  float *d_col = nullptr;
  QR_TRY(dev_alloc(&d_col, N * F));
  if (rowmajor) {
    float *d_row = nullptr;
    QR_TRY(dev_alloc(&d_row, N * F));
    QR_CUDA(cudaMemcpyAsync(d_row, feat, N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned) ((F + 31) / 32), (unsigned) ((N + 31) / 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(d_row, d_col, N, F);
    QR_CUDA(cudaGetLastError());
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_row);
  } else {
    QR_CUDA(cudaMemcpyAsync(d_col, feat, N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  int rc = build_binning(c, d_col);
  cudaFree(d_col);
  if (rc != QR_OK) return rc;

  // labels, gains, query offsets, ideal DCG per query, discount tables (host glibc, once)
  std::vector<uint32_t> qoff(Q + 1);
  uint32_t maxlen = 0;
  for (size_t q = 0; q <= Q; ++q) {
    if (q && qoffsets[q] < qoffsets[q - 1]) { set_error("query offsets must be non-decreasing"); return QR_EINVAL; }
    qoff[q] = (uint32_t) qoffsets[q];
    if (q) maxlen = std::max<uint32_t>(maxlen, qoff[q] - qoff[q - 1]);
  }
  c->maxlen = (maxlen + 3u) & ~3u;
  std::vector<double> gain(N), idcg(Q), lg(c->maxlen + 1), invlg(c->maxlen + 1);
  for (size_t i = 0; i < N; ++i) gain[i] = std::pow(2.0, (double) labels[i]);              // dcg.cc:37
  for (uint32_t i = 0; i <= c->maxlen; ++i) {
    lg[i] = std::log2((double) ((float) i + 2.0f));                                          // dcg.cc:37
    invlg[i] = 1.0 / std::log2((double) (i + 2));                                            // ndcg.cc:79
  }
  {
    std::vector<float> tmp;
    for (size_t q = 0; q < Q; ++q) {                                                         // ndcg.cc:35-47
      tmp.assign(labels + qoff[q], labels + qoff[q + 1]);
      std::sort(tmp.begin(), tmp.end(), std::greater<int>());
      const size_t size = std::min(c->cutoff, tmp.size());
      double dcg = 0.0;
      for (size_t i = 0; i < size; ++i) dcg += (std::pow(2.0, (double) tmp[i]) - 1.0) / lg[i];
      idcg[q] = dcg;
    }
  }
  QR_TRY(dev_alloc(&c->d_labels, N));
  QR_TRY(dev_alloc(&c->d_gain, N));
  QR_TRY(dev_alloc(&c->d_qoff, Q + 1));
  QR_TRY(dev_alloc(&c->d_idcg, Q));
  QR_TRY(dev_alloc(&c->d_lg, c->maxlen + 1));
  QR_TRY(dev_alloc(&c->d_invlg, c->maxlen + 1));
  QR_CUDA(cudaMemcpy(c->d_labels, labels, N * sizeof(float), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_gain, gain.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_qoff, qoff.data(), (Q + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_idcg, idcg.data(), Q * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_lg, lg.data(), (c->maxlen + 1) * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_invlg, invlg.data(), (c->maxlen + 1) * sizeof(double), cudaMemcpyHostToDevice));

  // state arrays (mart.cc:121-122, lambdamart.cc:38: zero-initialised)
  QR_TRY(dev_alloc(&c->d_scores, N));
  QR_TRY(dev_alloc(&c->d_lambda, N));
  QR_TRY(dev_alloc(&c->d_weight, N));
  QR_TRY(dev_alloc(&c->d_lamq, N));
  QR_TRY(dev_alloc(&c->d_maxabs, 1));
  QR_TRY(dev_alloc(&c->d_qexp, 1));
  QR_TRY(dev_alloc(&c->d_rankpos, N));
  QR_TRY(dev_alloc(&c->d_qndcg, Q));
  QR_TRY(dev_alloc(&c->d_metric, 1));
  QR_TRY(dev_alloc(&c->d_ids[0], N));
  QR_TRY(dev_alloc(&c->d_ids[1], N));
  QR_TRY(dev_alloc(&c->d_leaf_of_doc, N));
  QR_TRY(dev_alloc(&c->d_blockcnt, (N + kPartItems - 1) / kPartItems + 1));
  QR_TRY(dev_alloc(&c->d_partials, 1024));
  QR_CUDA(cudaMemset(c->d_scores, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_lambda, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_weight, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_leaf_of_doc, 0, N * sizeof(uint32_t)));
  QR_CUDA(cudaMemset(c->d_qexp, 0, sizeof(int)));

  const size_t maxleaves = c->oblivious ? ((size_t) 1 << params->treedepth) : params->nleaves;
  c->nslots = (int) (c->oblivious ? 3 * maxleaves / 2 + 4 : 2 * maxleaves + 4);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t hist_bytes = (size_t) c->nslots * c->ncells * 12;
  if (hist_bytes + (64u << 20) > free_b) {
    set_error("histogram pool needs %zu MB (%d nodes x %u cells); not enough device memory — bound the "
              "bin count with --num-thresholds", hist_bytes >> 20, c->nslots, c->ncells);
    return QR_ENOMEM;
  }
  QR_TRY(dev_alloc(&c->d_hist_sum, (size_t) c->nslots * c->ncells));
  QR_TRY(dev_alloc(&c->d_hist_cnt, (size_t) c->nslots * c->ncells));
  for (int i = c->nslots - 1; i >= 0; --i) c->free_slots.push_back(i);
  QR_TRY(dev_alloc(&c->d_fbest_score, 2 * F));
  QR_TRY(dev_alloc(&c->d_fbest_t, 2 * F));
  QR_TRY(dev_alloc(&c->d_res, 2));
  QR_CUDA(cudaMallocHost((void **) &c->h_res, 2 * sizeof(SplitResult)));
  QR_TRY(dev_alloc(&c->d_leafval, maxleaves + 1));
  QR_CUDA(cudaMallocHost((void **) &c->h_leafval, (maxleaves + 1) * sizeof(double)));
  QR_TRY(dev_alloc(&c->d_obv_scores, c->ncells));

  // opt in to large dynamic shared memory where needed
  const size_t hist_smem = (size_t) c->max_panel_cells * 12;
  if (hist_smem <= 200 * 1024) {
    const int bytes = (int) hist_smem;
    cudaFuncSetAttribute(hist_fast_kernel<uint8_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_fast_kernel<uint8_t, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_fast_kernel<uint16_t, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_fast_kernel<uint16_t, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  }
  QR_CUDA(cudaGetLastError());
  return QR_OK;
}

// ------------------------------------------------------------------------------------------
// ranking / metric / pseudo-responses
// ------------------------------------------------------------------------------------------
constexpr int kRankWarps = 4;
constexpr int kLambdaWarps = 4;

static int ensure_ranking(qr_ctx *c) {
  if (c->ranking_valid) return QR_OK;
  PhaseTimer pt(c, PH_RANK);
  const size_t smem = (size_t) kRankWarps * c->maxlen * 12;
  if (smem > 200 * 1024) { set_error("longest query (%u documents) exceeds the ranking kernel's shared-memory budget", c->maxlen); return QR_ELIMIT; }
  static bool attr_set = false;
  if (!attr_set || smem > 48 * 1024) {
    cudaFuncSetAttribute(rank_kernel<kRankWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
    attr_set = true;
  }
  const unsigned grid = (unsigned) ((c->Q + kRankWarps - 1) / kRankWarps);
  QR_LAUNCH(c, PH_RANK, rank_kernel<kRankWarps>, grid, kRankWarps * 32, smem, c->d_scores, c->d_labels,
            c->d_gain, c->d_qoff, c->d_idcg, c->d_lg, (uint32_t) c->Q, c->maxlen, c->cutoff,
            c->d_rankpos, c->d_qndcg);
  c->ranking_valid = true;
  return QR_OK;
}

static int compute_pseudo(qr_ctx *c) {
  if (!c->lambda) {
    PhaseTimer pt(c, PH_PSEUDO);
    QR_LAUNCH(c, PH_PSEUDO, mart_pseudo_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_scores,
              c->d_labels, c->N, c->d_lambda);
    return QR_OK;
  }
  QR_TRY(ensure_ranking(c));
  PhaseTimer pt(c, PH_PSEUDO);
  const size_t smem = (size_t) kLambdaWarps * ((size_t) c->maxlen * 24 + (size_t) 32 * kStageStride * 16);
  if (smem > 200 * 1024) { set_error("longest query (%u documents) exceeds the lambda kernel's shared-memory budget", c->maxlen); return QR_ELIMIT; }
  cudaFuncSetAttribute(lambda_kernel<kLambdaWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
  const unsigned grid = (unsigned) ((c->Q + kLambdaWarps - 1) / kLambdaWarps);
  QR_LAUNCH(c, PH_PSEUDO, lambda_kernel<kLambdaWarps>, grid, kLambdaWarps * 32, smem, c->d_scores,
            c->d_labels, c->d_gain, c->d_qoff, c->d_idcg, c->d_invlg, c->d_rankpos, (uint32_t) c->Q,
            c->maxlen, c->cutoff, c->d_lambda, c->d_weight);
  return QR_OK;
}

static int evaluate(qr_ctx *c, double *metric) {
  QR_TRY(ensure_ranking(c));
  {
    PhaseTimer pt(c, PH_RANK);
    if (c->comm) {
      // sum of per-query NDCG over all ranks, divided by the global query count
      QR_LAUNCH(c, PH_RANK, ndcg_mean_kernel, 1, c->exact ? 32 : 1024, 0, c->d_qndcg, (uint32_t) c->Q, 1u,
                c->exact, c->d_metric);
      QR_TRY(comm_allreduce_sum_f64(c->comm, c->d_metric, 1, c->stream));
    } else {
      QR_LAUNCH(c, PH_RANK, ndcg_mean_kernel, 1, c->exact ? 32 : 1024, 0, c->d_qndcg, (uint32_t) c->Q,
                (uint32_t) c->Q, c->exact, c->d_metric);
    }
  }
  double m = 0;
  QR_CUDA(cudaMemcpyAsync(&m, c->d_metric, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  if (c->comm) m /= (double) c->Q_global;
  if (metric) *metric = m;
  return QR_OK;
}

// ------------------------------------------------------------------------------------------
// tree growth
// ------------------------------------------------------------------------------------------
template <typename F>
static int dispatch_bins(const qr_ctx *c, F &&fn) {
  return c->bin_bytes == 1 ? fn(uint8_t()) : fn(uint16_t());
}

static int alloc_slot(qr_ctx *c) {
  if (c->free_slots.empty()) return -1;
  int s = c->free_slots.back();
  c->free_slots.pop_back();
  return s;
}
static void release_slot(qr_ctx *c, int &s) {
  if (s >= 0) c->free_slots.push_back(s);
  s = -1;
}

static int prepare_fixed_point(qr_ctx *c) {
  if (c->exact) return QR_OK;
  PhaseTimer pt(c, PH_HIST);
  QR_CUDA(cudaMemsetAsync(c->d_maxabs, 0, sizeof(unsigned long long), c->stream));
  QR_LAUNCH(c, PH_HIST, maxabs_kernel, 296, 256, 0, c->d_lambda, c->N, c->d_maxabs);
  if (c->comm) QR_TRY(comm_allreduce_max_u64(c->comm, c->d_maxabs, 1, c->stream));
  QR_LAUNCH(c, PH_HIST, choose_scale_kernel, 1, 1, 0, c->d_maxabs, ceil_log2(c->N_global) + 1, c->d_qexp);
  QR_LAUNCH(c, PH_HIST, quantize_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_lambda, c->N,
            c->d_qexp, c->d_lamq);
  return QR_OK;
}

// histogram of the documents ids[buf][lo, lo+n) (dense: documents lo..lo+n-1) into `slot`
// (+ squares partials into d_partials; returns their count)
static int build_hist(qr_ctx *c, int slot, bool dense, int buf, uint32_t lo, uint32_t n, bool root,
                      uint32_t *n_partials) {
  unsigned long long *hs = c->d_hist_sum + (size_t) slot * c->ncells;
  uint32_t *hc = c->d_hist_cnt + (size_t) slot * c->ncells;
  const uint32_t *ids = c->d_ids[buf];
  PhaseTimer pt(c, PH_HIST);
  QR_CUDA(cudaMemsetAsync(hs, 0, (size_t) c->ncells * 8, c->stream));
  QR_CUDA(cudaMemsetAsync(hc, 0, (size_t) c->ncells * 4, c->stream));
  const uint32_t F = (uint32_t) c->F;
  if (c->exact) {
    const unsigned grid = (F + 3) / 4;
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      if (dense) QR_LAUNCH(c, PH_HIST, (hist_exact_kernel<B, true>), grid, 128, 0, c->d_panels, c->N, ids, lo, n, c->d_lambda, c->d_thr_off, F, (double *) hs, hc);
      else QR_LAUNCH(c, PH_HIST, (hist_exact_kernel<B, false>), grid, 128, 0, c->d_panels, c->N, ids, lo, n, c->d_lambda, c->d_thr_off, F, (double *) hs, hc);
      return QR_OK;
    }));
    // root: unfused (rtnode_histogram.cc:199-203 in the oracle build); children: fused (:65-69)
    if (dense) QR_LAUNCH(c, PH_HIST, squares_exact_kernel<true>, 1, 32, 0, c->d_lambda, ids, lo, n, !root, c->d_partials);
    else QR_LAUNCH(c, PH_HIST, squares_exact_kernel<false>, 1, 32, 0, c->d_lambda, ids, lo, n, !root, c->d_partials);
    *n_partials = 1;
    return QR_OK;
  }
  const uint32_t want_slices = std::max<uint32_t>(1, (148u * 4u + c->npanels - 1) / c->npanels);
  uint32_t dpb = std::max<uint32_t>(2048u, (n + want_slices - 1) / want_slices);
  dpb = (dpb + 255u) & ~255u;
  const uint32_t slices = std::max<uint32_t>(1, (n + dpb - 1) / dpb);
  const size_t smem = (size_t) c->max_panel_cells * 12;
  const bool use_smem = smem <= 200 * 1024;
  dim3 grid(slices, c->npanels);
  QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    if (use_smem) {
      if (dense) QR_LAUNCH(c, PH_HIST, (hist_fast_kernel<B, true, true>), grid, 256, smem, c->d_panels, c->N, ids, lo, n, c->d_lamq, c->d_thr_off, F, hs, hc, dpb);
      else QR_LAUNCH(c, PH_HIST, (hist_fast_kernel<B, false, true>), grid, 256, smem, c->d_panels, c->N, ids, lo, n, c->d_lamq, c->d_thr_off, F, hs, hc, dpb);
    } else {
      if (dense) QR_LAUNCH(c, PH_HIST, (hist_fast_kernel<B, true, false>), grid, 256, 0, c->d_panels, c->N, ids, lo, n, c->d_lamq, c->d_thr_off, F, hs, hc, dpb);
      else QR_LAUNCH(c, PH_HIST, (hist_fast_kernel<B, false, false>), grid, 256, 0, c->d_panels, c->N, ids, lo, n, c->d_lamq, c->d_thr_off, F, hs, hc, dpb);
    }
    return QR_OK;
  }));
  const uint32_t sq_blocks = std::min<uint32_t>(64u, std::max<uint32_t>(1u, n / 4096u));
  if (dense) QR_LAUNCH(c, PH_HIST, squares_fast_kernel<true>, sq_blocks, 256, 0, c->d_lambda, ids, lo, n, c->d_partials);
  else QR_LAUNCH(c, PH_HIST, squares_fast_kernel<false>, sq_blocks, 256, 0, c->d_lambda, ids, lo, n, c->d_partials);
  *n_partials = sq_blocks;
  return QR_OK;
}

// multi-GPU: sum the freshly built (per-bin, not yet cumulative) histogram and the squares over ranks
static int reduce_hist(qr_ctx *c, int slot, uint32_t *n_partials) {
  if (!c->comm) return QR_OK;
  unsigned long long *hs = c->d_hist_sum + (size_t) slot * c->ncells;
  uint32_t *hc = c->d_hist_cnt + (size_t) slot * c->ncells;
  return comm_reduce_hist(c, hs, hc, n_partials);
}

// cumulative + right = parent - left + split scan; results land in c->h_res[0..1]
static int finalize_nodes(qr_ctx *c, int mode, int slotP, int slotL, int slotR, uint32_t n_partials,
                          double parent_squares) {
  PhaseTimer pt(c, PH_SCAN);
  FinalizeArgs a;
  a.hsum = c->d_hist_sum; a.hcnt = c->d_hist_cnt; a.ncells = c->ncells;
  a.slotP = slotP; a.slotL = slotL; a.slotR = slotR; a.mode = mode;
  a.minls = c->p.minleafsupport; a.qexp = c->d_qexp;
  a.fbest_score = c->d_fbest_score; a.fbest_t = c->d_fbest_t; a.F = (uint32_t) c->F;
  Finalize2Args b;
  b.hsum = c->d_hist_sum; b.hcnt = c->d_hist_cnt; b.ncells = c->ncells;
  b.slotL = slotL; b.slotR = slotR; b.mode = mode; b.qexp = c->d_qexp;
  b.fbest_score = c->d_fbest_score; b.fbest_t = c->d_fbest_t; b.F = (uint32_t) c->F;
  b.sq_partials = c->d_partials; b.n_partials = n_partials; b.parent_squares = parent_squares;
  b.res = c->d_res;
  if (c->exact) {
    QR_LAUNCH(c, PH_SCAN, finalize_kernel<true>, (unsigned) c->F, 256, 0, a, c->d_thr_off);
    QR_LAUNCH(c, PH_SCAN, finalize2_kernel<true>, 1, 32, 0, b, c->d_thr_off);
  } else {
    QR_LAUNCH(c, PH_SCAN, finalize_kernel<false>, (unsigned) c->F, 256, 0, a, c->d_thr_off);
    QR_LAUNCH(c, PH_SCAN, finalize2_kernel<false>, 1, 32, 0, b, c->d_thr_off);
  }
  QR_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, 2 * sizeof(SplitResult), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}

static int partition_node(qr_ctx *c, bool dense, int buf, uint32_t lo, uint32_t n, uint32_t f, uint32_t t,
                          uint32_t lcount, int dst_buf) {
  PhaseTimer pt(c, PH_PARTITION);
  const unsigned blocks = (n + kPartItems - 1) / kPartItems;
  const uint32_t *src = c->d_ids[buf];
  uint32_t *dst = c->d_ids[dst_buf];
  return dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    if (dense) {
      QR_LAUNCH(c, PH_PARTITION, (partition_count_kernel<B, true>), blocks, 256, 0, c->d_panels, c->N, src, lo, n, f, t, c->d_blockcnt);
      QR_LAUNCH(c, PH_PARTITION, (partition_scatter_kernel<B, true>), blocks, 256, 0, c->d_panels, c->N, src, dst, lo, n, f, t, c->d_blockcnt, lcount);
    } else {
      QR_LAUNCH(c, PH_PARTITION, (partition_count_kernel<B, false>), blocks, 256, 0, c->d_panels, c->N, src, lo, n, f, t, c->d_blockcnt);
      QR_LAUNCH(c, PH_PARTITION, (partition_scatter_kernel<B, false>), blocks, 256, 0, c->d_panels, c->N, src, dst, lo, n, f, t, c->d_blockcnt, lcount);
    }
    return QR_OK;
  });
}

// host replica of MaxHeap<RTNode*> (maxheap.h:31-106): same sift rules, so equal keys pop in
// the same order as in the reference
struct NodeHeap {
  struct Item { double key; int val; };
  std::vector<Item> arr;
  size_t size = 0;
  NodeHeap() { arr.push_back({DBL_MAX, -1}); }
  void push(double key, int val) {
    ++size;
    if (arr.size() <= size) arr.resize(size + 1);
    size_t p = size;
    while (key > arr[p >> 1].key) { arr[p] = arr[p >> 1]; p >>= 1; }
    arr[p] = {key, val};
  }
  int top() const { return arr[1].val; }
  void pop() {
    const Item last = arr[size--];
    size_t child, p = 1;
    while ((p << 1) <= size) {
      child = p << 1;
      if (child < size && arr[child + 1].key > arr[child].key) ++child;
      if (last.key < arr[child].key) arr[p] = arr[child];
      else break;
      p = child;
    }
    arr[p] = last;
  }
};

// the documents of node i live in ids[buf][lo, lo+n); the root is the identity list
static bool node_dense(const qr_ctx *c, int i) { return i == 0; }

// RegressionTree::split (rt.cc:209-362) for node i whose best split is already known
static int split_node(qr_ctx *c, int i, bool build_child_hists) {
  HostNode nd = c->nodes[i];
  const uint32_t f = nd.res.feature, t = nd.res.threshold_idx;
  const uint32_t lc = (uint32_t) nd.res.lcount;
  const bool dense = node_dense(c, i);
  const int dst_buf = dense ? 0 : 1 - nd.buf;
  QR_TRY(partition_node(c, dense, nd.buf, nd.lo, nd.n, f, t, lc, dst_buf));
  HostNode L, R;
  L.lo = nd.lo; L.n = lc; L.buf = dst_buf;
  R.lo = nd.lo + lc; R.n = nd.n - lc; R.buf = dst_buf;
  if (build_child_hists) {
    L.hist = alloc_slot(c);
    R.hist = alloc_slot(c);
    if (L.hist < 0 || R.hist < 0) { set_error("internal: histogram pool exhausted"); return QR_ECUDA; }
    uint32_t n_part = 0;
    QR_TRY(build_hist(c, L.hist, false, L.buf, L.lo, L.n, false, &n_part));
    QR_TRY(reduce_hist(c, L.hist, &n_part));
    QR_TRY(finalize_nodes(c, 1, nd.hist, L.hist, R.hist, n_part, nd.res.squares));
    L.res = c->h_res[0];
    R.res = c->h_res[1];
    if (c->comm) { L.n = (uint32_t) 0 + L.n; }  // local sizes stay local; res.n is global
  }
  const int li = (int) c->nodes.size();
  c->nodes.push_back(L);
  c->nodes.push_back(R);
  c->nodes[i].left = li;
  c->nodes[i].right = li + 1;
  c->rho += (double) lc / (double) c->N;
  c->sigma += (double) nd.n / (double) c->N;
  c->nsplits++;
  return QR_OK;
}

static int fit_leafwise(qr_ctx *c) {
  const size_t nleaves = c->p.nleaves;
  NodeHeap heap;
  size_t taken = 0;
  auto can_split = [&](int i) {
    const SplitResult &r = c->nodes[i].res;
    return r.deviance > 0.0 && r.valid;        // rt.cc:212, 312
  };
  if (can_split(0)) {
    QR_TRY(split_node(c, 0, true));
    heap.push(c->nodes[c->nodes[0].left].res.deviance, c->nodes[0].left);     // rt.cc:59-60
    heap.push(c->nodes[c->nodes[0].right].res.deviance, c->nodes[0].right);
  }
  while (heap.size != 0 && (nleaves == 0 || taken + heap.size < nleaves)) {   // rt.cc:64-65
    const int i = heap.top();
    heap.pop();
    if (can_split(i)) {
      QR_TRY(split_node(c, i, true));
      heap.push(c->nodes[c->nodes[i].left].res.deviance, c->nodes[i].left);
      heap.push(c->nodes[c->nodes[i].right].res.deviance, c->nodes[i].right);
    } else {
      ++taken;                                                                // rt.cc:78-79
    }
    release_slot(c, c->nodes[i].hist);                                        // rt.cc:83-84
  }
  return QR_OK;
}

static int fit_oblivious(qr_ctx *c) {
  const uint32_t depth = c->p.treedepth;
  std::vector<int> level{0};
  int *d_slots = nullptr;
  uint64_t *d_lcounts = nullptr, *h_lcounts = nullptr;
  const size_t maxnodes = (size_t) 1 << depth;
  QR_TRY(dev_alloc(&d_slots, maxnodes));
  QR_TRY(dev_alloc(&d_lcounts, maxnodes));
  QR_CUDA(cudaMallocHost((void **) &h_lcounts, maxnodes * sizeof(uint64_t)));
  int rc = QR_OK;
  for (uint32_t d = 0; d < depth && rc == QR_OK; ++d) {
    std::vector<int> slots;
    for (int i : level) slots.push_back(c->nodes[i].hist);
    auto body = [&]() -> int {
      {
        PhaseTimer pt(c, PH_SCAN);
        QR_CUDA(cudaMemcpyAsync(d_slots, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        if (c->exact) QR_LAUNCH(c, PH_SCAN, obv_level_kernel<true>, (unsigned) c->F, 256, 0, c->d_hist_sum, c->d_hist_cnt, c->ncells, d_slots, (uint32_t) slots.size(), c->d_thr_off, (uint32_t) c->F, c->p.minleafsupport, c->d_qexp, c->d_obv_scores);
        else QR_LAUNCH(c, PH_SCAN, obv_level_kernel<false>, (unsigned) c->F, 256, 0, c->d_hist_sum, c->d_hist_cnt, c->ncells, d_slots, (uint32_t) slots.size(), c->d_thr_off, (uint32_t) c->F, c->p.minleafsupport, c->d_qexp, c->d_obv_scores);
        QR_LAUNCH(c, PH_SCAN, obv_argmax_kernel, 1, 256, 0, c->d_obv_scores, c->d_thr_off, (uint32_t) c->F, c->d_hist_cnt, c->ncells, d_slots, (uint32_t) slots.size(), c->d_res, d_lcounts);
        QR_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, sizeof(SplitResult), cudaMemcpyDeviceToHost, c->stream));
        QR_CUDA(cudaMemcpyAsync(h_lcounts, d_lcounts, slots.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        QR_CUDA(cudaStreamSynchronize(c->stream));
      }
      return QR_OK;
    };
    rc = body();
    if (rc != QR_OK) break;
    const SplitResult best = c->h_res[0];
    if (!best.valid) break;                                   // ot.cc:96
    std::vector<int> next;
    for (size_t k = 0; k < level.size() && rc == QR_OK; ++k) {
      const int i = level[k];
      c->nodes[i].res.feature = best.feature;
      c->nodes[i].res.threshold_idx = best.threshold_idx;
      c->nodes[i].res.lcount = h_lcounts[k];
      c->nodes[i].res.valid = 1;
      if (c->comm) rc = comm_local_lcount(c, i);              // local left size for the partition
      if (rc == QR_OK) rc = split_node(c, i, d != depth - 1); // ot.cc:127
      if (rc != QR_OK) break;
      next.push_back(c->nodes[i].left);
      next.push_back(c->nodes[i].right);
      release_slot(c, c->nodes[i].hist);                      // ot.cc:157-160 (and the root's copy)
    }
    level.swap(next);
  }
  cudaFree(d_slots);
  cudaFree(d_lcounts);
  cudaFreeHost(h_lcounts);
  return rc;
}

static void collect_leaves(qr_ctx *c, int i) {
  if (c->nodes[i].is_leaf()) { c->leaves.push_back(i); return; }
  collect_leaves(c, c->nodes[i].left);      // rtnode.cc:34-46: left to right
  collect_leaves(c, c->nodes[i].right);
}

static void flatten(const qr_ctx *c, int i, qr_flat_tree *t, uint32_t *next) {
  const HostNode &nd = c->nodes[i];
  const uint32_t id = (*next)++;
  const bool leaf = nd.is_leaf();
  t->feature[id] = leaf ? -1 : (int32_t) nd.res.feature;
  t->threshold_idx[id] = leaf ? 0xffffffffu : nd.res.threshold_idx;
  t->threshold[id] = leaf ? 0.f : c->thr[nd.res.feature][nd.res.threshold_idx];   // rt.cc:317-318
  t->left[id] = t->right[id] = -1;
  if (t->value) t->value[id] = leaf ? nd.value : (nd.res.n ? nd.res.sum / (double) nd.res.n : 0.0);  // rtnode.h:105
  if (t->deviance) t->deviance[id] = nd.res.deviance;
  if (t->count) t->count[id] = nd.res.n;
  if (!leaf) {
    t->left[id] = (int32_t) *next;
    flatten(c, nd.left, t, next);
    t->right[id] = (int32_t) *next;
    flatten(c, nd.right, t, next);
  }
}

static int fit_tree(qr_ctx *c, qr_flat_tree *out) {
  // release histograms still held by the previous tree
  for (auto &nd : c->nodes) release_slot(c, nd.hist);
  c->nodes.clear();
  c->leaves.clear();
  c->rho = c->sigma = 0;
  c->nsplits = 0;
  c->has_tree = false;

  QR_TRY(prepare_fixed_point(c));
  HostNode root;
  root.lo = 0; root.n = (uint32_t) c->N; root.buf = 0;
  root.hist = alloc_slot(c);
  uint32_t n_part = 0;
  QR_TRY(build_hist(c, root.hist, true, 0, 0, (uint32_t) c->N, true, &n_part));   // mart.cc:335
  QR_TRY(reduce_hist(c, root.hist, &n_part));
  QR_TRY(finalize_nodes(c, 0, -1, root.hist, -1, n_part, 0.0));
  root.res = c->h_res[0];
  c->nodes.push_back(root);

  QR_TRY(c->oblivious ? fit_oblivious(c) : fit_leafwise(c));

  // leaves in DFS order, leaf outputs, doc -> leaf map
  collect_leaves(c, 0);
  const size_t nl = c->leaves.size();
  std::vector<LeafSeg> segs(nl);
  for (size_t k = 0; k < nl; ++k) {
    const HostNode &nd = c->nodes[c->leaves[k]];
    segs[k] = LeafSeg{nd.lo, nd.n, nd.buf, 0};
  }
  {
    PhaseTimer pt(c, PH_LEAF);
    LeafSeg *d_segs = nullptr;
    QR_TRY(dev_alloc(&d_segs, nl));
    QR_CUDA(cudaMemcpyAsync(d_segs, segs.data(), nl * sizeof(LeafSeg), cudaMemcpyHostToDevice, c->stream));
    const bool root_only = nl == 1;
    if (c->comm) {
      QR_TRY(comm_leaf_fit(c, d_segs, (uint32_t) nl, root_only));
    } else {
      QR_LAUNCH(c, PH_LEAF, leaf_fit_kernel, (unsigned) nl, 256, 0, d_segs, c->d_ids[0], c->d_ids[1], root_only,
                c->d_lambda, c->lambda ? c->d_weight : nullptr, c->exact, c->d_leafval, c->d_leaf_of_doc);
    }
    QR_CUDA(cudaMemcpyAsync(c->h_leafval, c->d_leafval, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_segs);
  }
  for (size_t k = 0; k < nl; ++k) c->nodes[c->leaves[k]].value = c->h_leafval[k];
  c->has_tree = true;

  if (out) {
    const uint32_t nn = (uint32_t) c->nodes.size();
    if (out->capacity < nn) { set_error("qr_flat_tree capacity %u < %u nodes", out->capacity, nn); return QR_EINVAL; }
    uint32_t next = 0;
    flatten(c, 0, out, &next);
    out->nnodes = nn;
    out->nleaves = (uint32_t) nl;
  }
  return QR_OK;
}

static int update_modelscores(qr_ctx *c, double weight) {
  if (!c->has_tree) { set_error("qr_update_modelscores: no fitted tree"); return QR_EINVAL; }
  PhaseTimer pt(c, PH_LEAF);
  QR_LAUNCH(c, PH_LEAF, update_scores_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_leaf_of_doc,
            c->d_leafval, weight, c->N, c->d_scores);
  c->ranking_valid = false;
  return QR_OK;
}

}  // namespace qr

using namespace qr;

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *qr_last_error(void) { return g_last_error.c_str(); }

int qr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int qr_ctx_create(const float *feat, size_t N, size_t F, const float *labels, const uint64_t *qoff,
                  size_t Q, const qr_params *params, qr_ctx **out) {
  if (out) *out = nullptr;
  int rc = ctx_create_common(feat, false, N, F, labels, qoff, Q, params, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_ctx_create_rowmajor(const float *feat, size_t N, size_t F, const float *labels,
                           const uint64_t *qoff, size_t Q, const qr_params *params, qr_ctx **out) {
  if (out) *out = nullptr;
  int rc = ctx_create_common(feat, true, N, F, labels, qoff, Q, params, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_ctx_destroy(qr_ctx *c) {
  if (!c) return QR_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  comm_destroy(c->comm);
  void *ptrs[] = {c->d_panels, c->d_thr_off, c->d_labels, c->d_gain, c->d_qoff, c->d_idcg, c->d_invlg, c->d_lg,
                  c->d_scores, c->d_lambda, c->d_weight, c->d_lamq, c->d_maxabs, c->d_qexp, c->d_rankpos,
                  c->d_qndcg, c->d_metric, c->d_ids[0], c->d_ids[1], c->d_leaf_of_doc, c->d_blockcnt,
                  c->d_partials, c->d_hist_sum, c->d_hist_cnt, c->d_fbest_score, c->d_fbest_t, c->d_res,
                  c->d_leafval, c->d_obv_scores};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (c->h_res) cudaFreeHost(c->h_res);
  if (c->h_leafval) cudaFreeHost(c->h_leafval);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return QR_OK;
}

#define QR_CHECK_CTX(c)                                                  \
  do {                                                                   \
    if (!(c)) { set_error("null context"); return QR_EINVAL; }           \
    cudaSetDevice((c)->device);                                          \
  } while (0)

int qr_get_thresholds(qr_ctx *c, size_t f, const float **thr, size_t *n) {
  QR_CHECK_CTX(c);
  if (f >= c->F || !thr || !n) { set_error("qr_get_thresholds: bad argument"); return QR_EINVAL; }
  *thr = c->thr[f].data();
  *n = c->thr[f].size();
  return QR_OK;
}

int qr_compute_pseudoresponses(qr_ctx *c) {
  QR_CHECK_CTX(c);
  return compute_pseudo(c);
}

int qr_fit_tree(qr_ctx *c, qr_flat_tree *out) {
  QR_CHECK_CTX(c);
  return fit_tree(c, out);
}

int qr_update_modelscores(qr_ctx *c, double weight) {
  QR_CHECK_CTX(c);
  return update_modelscores(c, weight);
}

int qr_apply_tree(qr_ctx *c, const qr_flat_tree *t, double weight) {
  QR_CHECK_CTX(c);
  if (!t || t->nnodes == 0) { set_error("qr_apply_tree: empty tree"); return QR_EINVAL; }
  const uint32_t n = t->nnodes;
  for (uint32_t i = 0; i < n; ++i)
    if (t->feature[i] >= 0 && ((size_t) t->feature[i] >= c->F || t->threshold_idx[i] >= c->thr[t->feature[i]].size())) {
      set_error("qr_apply_tree: node %u does not belong to this context's binning", i);
      return QR_EINVAL;
    }
  int32_t *d_feat, *d_left, *d_right; uint32_t *d_tidx; double *d_val;
  QR_TRY(dev_alloc(&d_feat, n)); QR_TRY(dev_alloc(&d_left, n)); QR_TRY(dev_alloc(&d_right, n));
  QR_TRY(dev_alloc(&d_tidx, n)); QR_TRY(dev_alloc(&d_val, n));
  QR_CUDA(cudaMemcpyAsync(d_feat, t->feature, n * 4, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaMemcpyAsync(d_left, t->left, n * 4, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaMemcpyAsync(d_right, t->right, n * 4, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaMemcpyAsync(d_tidx, t->threshold_idx, n * 4, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaMemcpyAsync(d_val, t->value, n * 8, cudaMemcpyHostToDevice, c->stream));
  DevTree dt{d_feat, d_tidx, d_left, d_right, d_val};
  int rc = dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    QR_LAUNCH(c, PH_LEAF, apply_tree_kernel<B>, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_panels, c->N, dt, weight, c->d_scores);
    return QR_OK;
  });
  cudaStreamSynchronize(c->stream);
  cudaFree(d_feat); cudaFree(d_left); cudaFree(d_right); cudaFree(d_tidx); cudaFree(d_val);
  c->ranking_valid = false;
  return rc;
}

int qr_evaluate(qr_ctx *c, double *metric) {
  QR_CHECK_CTX(c);
  return evaluate(c, metric);
}

int qr_boost_iteration(qr_ctx *c, qr_flat_tree *tree, double *metric) {
  QR_CHECK_CTX(c);
  QR_TRY(compute_pseudo(c));                       // mart.cc:331
  QR_TRY(fit_tree(c, tree));                       // mart.cc:335-339
  QR_TRY(update_modelscores(c, c->p.shrinkage));   // mart.cc:342-345
  if (metric) QR_TRY(evaluate(c, metric));         // mart.cc:347
  return QR_OK;
}

int qr_get_scores(qr_ctx *c, double *s) {
  QR_CHECK_CTX(c);
  QR_CUDA(cudaMemcpyAsync(s, c->d_scores, c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}
int qr_set_scores(qr_ctx *c, const double *s) {
  QR_CHECK_CTX(c);
  QR_CUDA(cudaMemcpyAsync(c->d_scores, s, c->N * 8, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  c->ranking_valid = false;
  return QR_OK;
}
int qr_get_pseudoresponses(qr_ctx *c, double *lam, double *w) {
  QR_CHECK_CTX(c);
  if (lam) QR_CUDA(cudaMemcpyAsync(lam, c->d_lambda, c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  if (w) QR_CUDA(cudaMemcpyAsync(w, c->d_weight, c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}
int qr_set_pseudoresponses(qr_ctx *c, const double *lam, const double *w) {
  QR_CHECK_CTX(c);
  if (lam) QR_CUDA(cudaMemcpyAsync(c->d_lambda, lam, c->N * 8, cudaMemcpyHostToDevice, c->stream));
  if (w) QR_CUDA(cudaMemcpyAsync(c->d_weight, w, c->N * 8, cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}
int qr_get_leaf_assignment(qr_ctx *c, uint32_t *leaf) {
  QR_CHECK_CTX(c);
  QR_CUDA(cudaMemcpyAsync(leaf, c->d_leaf_of_doc, c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}
int qr_get_bins(qr_ctx *c, size_t f, uint32_t *bins) {
  QR_CHECK_CTX(c);
  if (f >= c->F) { set_error("qr_get_bins: feature out of range"); return QR_EINVAL; }
  const size_t p = f / c->fpp, j = f % c->fpp;
  std::vector<unsigned char> rows(c->N * 16);
  QR_CUDA(cudaMemcpy(rows.data(), c->d_panels + p * c->N, c->N * 16, cudaMemcpyDeviceToHost));
  for (size_t d = 0; d < c->N; ++d) {
    if (c->bin_bytes == 1) bins[d] = rows[d * 16 + j];
    else { uint16_t v; memcpy(&v, &rows[d * 16 + 2 * j], 2); bins[d] = v; }
  }
  return QR_OK;
}
int qr_get_ranking(qr_ctx *c, uint32_t *pos) {
  QR_CHECK_CTX(c);
  QR_TRY(ensure_ranking(c));
  QR_CUDA(cudaMemcpyAsync(pos, c->d_rankpos, c->N * 4, cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  return QR_OK;
}
int qr_last_tree_stats(qr_ctx *c, double *rho, double *sigma, uint32_t *nsplits) {
  QR_CHECK_CTX(c);
  if (rho) *rho = c->rho;
  if (sigma) *sigma = c->sigma;
  if (nsplits) *nsplits = c->nsplits;
  return QR_OK;
}
uint64_t qr_launch_count(qr_ctx *c) { return c ? c->launches : 0; }
int qr_phase_times(qr_ctx *c, double ms[6], uint64_t launches[6], int reset) {
  QR_CHECK_CTX(c);
  for (int i = 0; i < kNumPhases; ++i) {
    if (ms) ms[i] = c->phase_ms[i];
    if (launches) launches[i] = c->phase_launches[i];
    if (reset) { c->phase_ms[i] = 0; c->phase_launches[i] = 0; }
  }
  return QR_OK;
}
int qr_timer_start(qr_ctx *c) {
  QR_CHECK_CTX(c);
  QR_CUDA(cudaEventRecord(c->ev_t0, c->stream));
  return QR_OK;
}
int qr_timer_stop(qr_ctx *c, double *ms) {
  QR_CHECK_CTX(c);
  QR_CUDA(cudaEventRecord(c->ev_t1, c->stream));
  QR_CUDA(cudaEventSynchronize(c->ev_t1));
  float f = 0;
  QR_CUDA(cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
  if (ms) *ms = f;
  return QR_OK;
}
int qr_set_profiling(qr_ctx *c, int enabled) {
  QR_CHECK_CTX(c);
  c->profiling = enabled != 0;
  return QR_OK;
}

}  // extern "C"
