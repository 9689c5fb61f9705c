// quickrank_b200 — host side of the training C ABI (include/quickrank_b200.h).
//
// The host keeps what the reference keeps on the host inside RegressionTree::fit (rt.cc:49-163):
// the max-heap of frontier nodes keyed by deviance (maxheap.h:31-106) and the node bookkeeping.
// All per-document and per-bin work runs in the kernels of qr_kernels.cuh.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <functional>
#include <limits>
#include <thread>

#include <cub/cub.cuh>

#include "qr_comm.cuh"
#include "qr_tree_kernels.cuh"

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

// launch with programmatic stream serialization: the kernel may become resident before its predecessor in the
// stream has finished; it must execute griddepcontrol.wait (pdl_wait) before touching the predecessor's output
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

#define QR_LAUNCH_PDL(ctx, phase, kernel, grid, block, smem, ...)                              \
  do {                                                                                         \
    cudaError_t _le = launch_pdl(kernel, grid, block, smem, (ctx)->stream, __VA_ARGS__);       \
    (ctx)->launches++;                                                                         \
    (ctx)->phase_launches[phase]++;                                                            \
    if (_le != cudaSuccess) {                                                                  \
      qr::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, #kernel,             \
                    cudaGetErrorString(_le));                                                  \
      return QR_ECUDA;                                                                         \
    }                                                                                          \
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
static int finish_binning(qr_ctx *c, const float *d_col, uint32_t max_bin);

static int build_binning(qr_ctx *c, const float *d_col, const qr_ctx *thr_from) {
  const size_t N = c->N, F = c->F;
  cudaStream_t st = c->stream;
  if (thr_from) {   // evaluation context: reuse the training thresholds, only the bin map is new
    c->thr = thr_from->thr;
    uint32_t max_bin = 0;
    for (size_t f = 0; f < F; ++f) max_bin = std::max<uint32_t>(max_bin, (uint32_t) c->thr[f].size() - 1);
    int *d_bad = nullptr;
    QR_TRY(dev_alloc(&d_bad, 1));
    QR_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    uint32_t *d_scratch = nullptr;
    QR_TRY(dev_alloc(&d_scratch, N));
    for (size_t f = 0; f < F; ++f)
      flip_keys_kernel<<<(unsigned) ((N + 255) / 256), 256, 0, st>>>(d_col + f * N, d_scratch, N, d_bad);
    int bad = 0;
    QR_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_bad);
    cudaFree(d_scratch);
    if (bad) { set_error("feature matrix contains NaN or infinite values"); return QR_EINVAL; }
    // a document sample holds training documents only: no bin beyond those the sampled context occupies, same bin width
    if (!c->eval_only && thr_from->bin_bytes == 1) max_bin = std::min<uint32_t>(max_bin, 255u);
    return finish_binning(c, d_col, max_bin);
  }
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

  // phase A: per feature, the distinct values of the local documents (ascending in the
  // reference's radix order) and the first / last value of that order
  struct Local { std::vector<float> uniq; int over = 0; float fmin = 0, fmax = 0; };
  std::vector<Local> loc(F);
  const size_t nth = (size_t) c->p.nthresholds;
  const unsigned blocks = (unsigned) ((N + 255) / 256);
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
    QR_CUDA(cudaMemcpyAsync(&loc[f].fmin, d_vals, sizeof(float), cudaMemcpyDeviceToHost, st));
    QR_CUDA(cudaMemcpyAsync(&loc[f].fmax, d_vals + (N - 1), sizeof(float), cudaMemcpyDeviceToHost, st));
    QR_CUDA(cudaStreamSynchronize(st));
    if (nth != 0 && (size_t) nu > nth) {
      loc[f].over = 1;   // more distinct values than thresholds already locally: equal-width mode
    } else {
      loc[f].uniq.resize((size_t) nu);
      QR_CUDA(cudaMemcpy(loc[f].uniq.data(), d_uniq, (size_t) nu * sizeof(float), cudaMemcpyDeviceToHost));
    }
  }
  int bad = 0;
  QR_CUDA(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  cudaFree(d_keys); cudaFree(d_keys_out); cudaFree(d_vals); cudaFree(d_uniq);
  cudaFree(d_flags); cudaFree(d_num); cudaFree(d_bad); cudaFree(d_tmp);

  // phase B (several ranks): every rank must end up with the thresholds of the WHOLE dataset
  if (c->comm) {
    auto key = [](float v) { uint32_t b; memcpy(&b, &v, 4); return b ^ ((uint32_t) (-(int32_t) (b >> 31)) | 0x80000000u); };
    struct Head { int bad, over, count; float fmin, fmax; };
    std::vector<Head> head(F);
    size_t total = 0;
    for (size_t f = 0; f < F; ++f) {
      head[f] = Head{bad, loc[f].over, (int) loc[f].uniq.size(), loc[f].fmin, loc[f].fmax};
      total += loc[f].uniq.size();
    }
    std::vector<unsigned char> all_head;
    QR_TRY(comm_allgather_host(c->comm, head.data(), F * sizeof(Head), &all_head, st));
    const int world = comm_world(c->comm);
    const size_t head_stride = all_head.size() / world;
    size_t max_total = 0;
    for (int r = 0; r < world; ++r) {
      const Head *h = reinterpret_cast<const Head *>(all_head.data() + r * head_stride);
      size_t t = 0;
      for (size_t f = 0; f < F; ++f) t += (size_t) h[f].count;
      max_total = std::max(max_total, t);
    }
    std::vector<float> flatv(std::max<size_t>(max_total, 1), 0.f);
    size_t o = 0;
    for (size_t f = 0; f < F; ++f) { std::copy(loc[f].uniq.begin(), loc[f].uniq.end(), flatv.begin() + o); o += loc[f].uniq.size(); }
    std::vector<unsigned char> all_vals;
    QR_TRY(comm_allgather_host(c->comm, flatv.data(), flatv.size() * sizeof(float), &all_vals, st));
    const size_t val_stride = all_vals.size() / world;
    std::vector<size_t> cursor(world, 0);
    for (size_t f = 0; f < F; ++f) {
      std::vector<float> merged;
      Local g;
      bool first = true;
      for (int r = 0; r < world; ++r) {
        const Head &h = reinterpret_cast<const Head *>(all_head.data() + r * head_stride)[f];
        const float *v = reinterpret_cast<const float *>(all_vals.data() + r * val_stride) + cursor[r];
        bad |= h.bad;
        g.over |= h.over;
        if (first || key(h.fmin) < key(g.fmin)) g.fmin = h.fmin;
        if (first || key(h.fmax) > key(g.fmax)) g.fmax = h.fmax;
        first = false;
        merged.insert(merged.end(), v, v + h.count);
        cursor[r] += (size_t) h.count;
      }
      if (!g.over) {
        std::sort(merged.begin(), merged.end(), [&](float a, float b) { return key(a) < key(b); });
        for (float v : merged)
          if (g.uniq.empty() || g.uniq.back() < v) g.uniq.push_back(v);   // mart.cc:148-151
        if (nth != 0 && g.uniq.size() > nth) { g.over = 1; g.uniq.clear(); }
      }
      loc[f] = std::move(g);
    }
  }
  if (bad) {
    set_error("feature matrix contains NaN or infinite values (the reference's bin map is undefined for them)");
    return QR_EINVAL;
  }

  // phase C: threshold lists
  c->thr.assign(F, std::vector<float>());
  uint32_t max_bin = 0;
  for (size_t f = 0; f < F; ++f) {
    std::vector<float> &t = c->thr[f];
    if (!loc[f].over) {                          // mart.cc:155-158: distinct values + FLT_MAX
      t = loc[f].uniq;
      t.push_back(FLT_MAX);
      max_bin = std::max<uint32_t>(max_bin, (uint32_t) (t.size() - 2));
    } else {                                     // mart.cc:159-169: equal width, float accumulation
      t.resize(nth + 1);
      float cur = loc[f].fmin;
      const float step = (float) std::fabs((double) (loc[f].fmax - cur)) / (float) nth;
      for (size_t j = 0; j != nth; cur += step) t[j++] = cur;
      t[nth] = FLT_MAX;
      max_bin = std::max<uint32_t>(max_bin, (uint32_t) nth);
    }
  }
  return finish_binning(c, d_col, max_bin);
}

// phase D: bin width, cell layout, device copies of the thresholds, the panel matrix
static int binning_layout(qr_ctx *c, uint32_t max_bin, float **d_thr_out);
static int finish_binning(qr_ctx *c, const float *d_col, uint32_t max_bin) {
  const size_t N = c->N, F = c->F;
  cudaStream_t st = c->stream;
  float *d_thr = nullptr;
  QR_TRY(binning_layout(c, max_bin, &d_thr));
  QR_TRY(dev_alloc(&c->d_panels, (size_t) c->npanels * N));
  // FAST mode gathers the built child's documents from a document-major copy (QR_ROW_COPY=0: from the panels)
  if (!c->exact && c->npanels > 1 && (getenv("QR_ROW_COPY") == nullptr || atoi(getenv("QR_ROW_COPY")) != 0))
    QR_TRY(dev_alloc(&c->d_rows, (size_t) c->npanels * N));
  dim3 grid((unsigned) ((N + 127) / 128), c->npanels);
  if (c->bin_bytes == 1)
    binning_kernel<uint8_t><<<grid, 128, 0, st>>>(d_col, N, (uint32_t) F, d_thr, c->d_thr_off, c->d_panels, c->npanels, c->d_rows);
  else
    binning_kernel<uint16_t><<<grid, 128, 0, st>>>(d_col, N, (uint32_t) F, d_thr, c->d_thr_off, c->d_panels, c->npanels, c->d_rows);
  QR_CUDA(cudaGetLastError());
  QR_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_thr);
  return QR_OK;
}

// bin width and cell layout from c->thr; uploads the threshold offsets (and, when asked, the thresholds)
static int binning_layout(qr_ctx *c, uint32_t max_bin, float **d_thr_out) {
  const size_t F = c->F;
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
  QR_TRY(dev_alloc(&c->d_thr_off, F + 1));
  QR_CUDA(cudaMemcpy(c->d_thr_off, c->thr_off.data(), (F + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (d_thr_out) {
    float *d_thr = nullptr;
    QR_TRY(dev_alloc(&d_thr, c->ncells));
    QR_CUDA(cudaMemcpy(d_thr, flat.data(), c->ncells * sizeof(float), cudaMemcpyHostToDevice));
    *d_thr_out = d_thr;
  }
  return QR_OK;
}

// REFERENCE mode, large nodes (qr_exact_kernels.cuh): per feature, the documents in (bin, document) order — a stable
// sort of the bin column, made once since the bins never change — and the position of every histogram cell in it.
// QR_EXACT_WALK_MIN: built children of at least this many documents are accumulated by walking those lists (default
// max(N/8, 65536); never below N/16, which bounds the large nodes of a round by kWalkMax).
static int build_exact_walk_tables(qr_ctx *c) {
  const size_t N = c->N, F = c->F;
  cudaStream_t st = c->stream;
  QR_TRY(dev_alloc((qr::SqChunk **) &c->d_sq_chunks, N / qr::kSqChunk + c->max_tasks + 2));
  QR_TRY(dev_alloc(&c->d_sq_replayed, 1));
  QR_CUDA(cudaMemset(c->d_sq_replayed, 0, sizeof(unsigned long long)));
  size_t want = std::max<size_t>(N / 8, 65536);
  if (const char *e = getenv("QR_EXACT_WALK_MIN")) want = (size_t) std::max<long long>(1, atoll(e));
  want = std::max(want, N / 16 + 1);
  c->walk_min = (uint32_t) std::min<size_t>(want, 0xffffffffu);
  if (N > 0xfffffff0u) { set_error("QR_HIST_REFERENCE: more than 2^32 documents"); return QR_ELIMIT; }
  QR_TRY(dev_alloc(&c->d_perm, F * N));
  QR_TRY(dev_alloc(&c->d_cell_pos, (size_t) c->ncells + 1));
  QR_TRY(dev_alloc(&c->d_mark, N));
  QR_CUDA(cudaMemsetAsync(c->d_mark, 0, N * sizeof(uint32_t), st));
  uint32_t *d_keys = nullptr, *d_keys_out = nullptr, *d_iota = nullptr;
  void *d_tmp = nullptr;
  QR_TRY(dev_alloc(&d_keys, N));
  QR_TRY(dev_alloc(&d_keys_out, N));
  QR_TRY(dev_alloc(&d_iota, N));
  size_t tmp_bytes = 0;
  const int bits = c->bin_bytes == 1 ? 8 : 16;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys_out, d_iota, c->d_perm, (int) N, 0, bits, st);
  QR_CUDA(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
  const unsigned blocks = (unsigned) ((N + 255) / 256);
  int rc = QR_OK;
  std::vector<uint32_t> ends(2 * F, 0u);   // first and last bin of every feature's sorted column
  for (size_t f = 0; f < F && rc == QR_OK; ++f) {
    if (c->bin_bytes == 1) qr::bin_column_kernel<uint8_t><<<blocks, 256, 0, st>>>(c->d_panels, N, (uint32_t) f, d_keys, d_iota);
    else qr::bin_column_kernel<uint16_t><<<blocks, 256, 0, st>>>(c->d_panels, N, (uint32_t) f, d_keys, d_iota);
    size_t tb = tmp_bytes;
    if (cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys_out, d_iota, c->d_perm + f * N, (int) N, 0, bits, st) != cudaSuccess) {
      set_error("sorting the bin column of feature %zu failed", f);
      rc = QR_ECUDA;
      break;
    }
    const uint32_t cells = c->thr_off[f + 1] - c->thr_off[f];
    qr::cell_pos_kernel<<<(cells + 255) / 256, 256, 0, st>>>(d_keys_out, N, (uint32_t) f, c->thr_off[f], cells, c->d_cell_pos);
    cudaMemcpyAsync(&ends[2 * f], d_keys_out, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&ends[2 * f + 1], d_keys_out + (N - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("building the sorted document lists failed: %s", cudaGetErrorString(cudaGetLastError())); rc = QR_ECUDA; }
  }
  const unsigned long long end = (unsigned long long) F * N;
  if (rc == QR_OK && cudaMemcpy(c->d_cell_pos + c->ncells, &end, sizeof(end), cudaMemcpyHostToDevice) != cudaSuccess) rc = QR_ECUDA;
  // single-bin features (every document in one bin) cannot be split on once a leaf needs >= 1 document, and feature 0
  // apart (node totals, rtnode.h:99-104) nothing reads their sums: they are not accumulated
  if (rc == QR_OK && c->p.minleafsupport >= 1 && (getenv("QR_EXACT_SKIP") == nullptr || atoi(getenv("QR_EXACT_SKIP")) != 0)) {
    std::vector<uint8_t> skip(F, 0);
    bool any = false;
    for (size_t f = 1; f < F; ++f) if (ends[2 * f] == ends[2 * f + 1]) { skip[f] = 1; any = true; }
    if (any) {
      rc = dev_alloc(&c->d_fskip, F);
      if (rc == QR_OK && cudaMemcpy(c->d_fskip, skip.data(), F, cudaMemcpyHostToDevice) != cudaSuccess) rc = QR_ECUDA;
    }
  }
  cudaFree(d_keys); cudaFree(d_keys_out); cudaFree(d_iota); cudaFree(d_tmp);
  return rc;
}

static int init_root_counts(qr_ctx *c);

// QR_INIT_TIMING=1: wall-clock breakdown of qr_ctx_create on stderr (development aid)
struct InitClock {
  bool on;
  std::chrono::steady_clock::time_point t0;
  InitClock() : on(getenv("QR_INIT_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char *what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[qr init] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

static int ctx_create_state(qr_ctx *c, const float *labels, const uint64_t *qoffsets, const qr_params *params, InitClock &clk);

// The bins of a document sample taken from the sampled context on the device: same thresholds, same bin width and
// panel layout, rows src[0..N) of its panels — no feature values cross the bus and nothing is binned again.
static int gather_sample_bins(qr_ctx *c, const qr_ctx *from, const uint32_t *src_host) {
  const size_t N = c->N;
  QR_CUDA(cudaMemcpyAsync(c->d_src_doc, src_host, N * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  QR_CUDA(cudaStreamSynchronize(from->stream));
  sample_gather_panels_kernel<<<dim3((unsigned) ((N + 255) / 256), c->npanels), 256, 0, c->stream>>>(
      from->d_panels, from->N, c->d_src_doc, N, c->npanels, c->d_panels, c->d_rows);
  QR_CUDA(cudaGetLastError());
  QR_CUDA(cudaStreamSynchronize(c->stream));   // (src_host is pageable memory of the caller)
  return QR_OK;
}

static int gather_binning(qr_ctx *c, const qr_ctx *from, const uint32_t *src_host) {
  const size_t N = c->N, F = c->F;
  c->thr = from->thr;
  uint32_t max_bin = 0;
  for (size_t f = 0; f < F; ++f) max_bin = std::max<uint32_t>(max_bin, (uint32_t) c->thr[f].size() - 1);
  if (from->bin_bytes == 1) max_bin = std::min<uint32_t>(max_bin, 255u);
  QR_TRY(binning_layout(c, max_bin, nullptr));
  if (c->bin_bytes != from->bin_bytes || c->npanels != from->npanels || c->ncells != from->ncells) {
    set_error("internal: the sample's bin layout differs from the sampled context's");
    return QR_ECUDA;
  }
  const size_t cap = std::max(c->cap_N, N);
  QR_TRY(dev_alloc(&c->d_src_doc, cap));
  QR_TRY(dev_alloc(&c->d_panels, (size_t) c->npanels * cap));
  if (!c->exact && c->npanels > 1 && (getenv("QR_ROW_COPY") == nullptr || atoi(getenv("QR_ROW_COPY")) != 0))
    QR_TRY(dev_alloc(&c->d_rows, (size_t) c->npanels * cap));
  return gather_sample_bins(c, from, src_host);
}

static int ctx_create_common(const float *feat, bool rowmajor, size_t N, size_t F, const float *labels,
                             const uint64_t *qoffsets, size_t Q, const qr_params *params,
                             const unsigned char *comm_id, int rank, int world, const qr_ctx *thr_from,
                             qr_ctx **out, bool sample = false, const uint32_t *gather_src = nullptr) {
  if ((!feat && !gather_src) || !labels || !qoffsets || !params || !out || N == 0 || F == 0 || Q == 0) {
    set_error("qr_ctx_create: null or empty argument");
    return QR_EINVAL;
  }
  if (N >= ((size_t) 1 << 31)) { set_error("N >= 2^31 documents per GPU is not supported"); return QR_ELIMIT; }
  if (qoffsets[0] != 0 || qoffsets[Q] != N) { set_error("query offsets must start at 0 and end at N"); return QR_EINVAL; }
  if (params->algo > QR_ALGO_OBVLAMBDAMART) { set_error("unknown algo %u", params->algo); return QR_EINVAL; }
  InitClock clk;
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
  if (c->oblivious && (params->treedepth == 0 || params->treedepth > 15)) {
    delete c; set_error("treedepth must be in 1..15"); return QR_EINVAL;
  }
  if (!c->oblivious && params->nleaves < 1) { delete c; set_error("nleaves must be >= 1"); return QR_EINVAL; }
  *out = c;  // destroyed by the caller on failure paths below via qr_ctx_destroy
  QR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  clk.lap("cuda context + stream");
  QR_CUDA(cudaEventCreate(&c->ev0));
  QR_CUDA(cudaEventCreate(&c->ev1));
  QR_CUDA(cudaEventCreate(&c->ev_t0));
  QR_CUDA(cudaEventCreate(&c->ev_t1));
  QR_CUDA(cudaEventCreate(&c->ev_k0));
  QR_CUDA(cudaEventCreate(&c->ev_k1));

  if (comm_id != nullptr && world > 1) {
    if (c->exact) { set_error("reference-order accumulation (QR_HIST_REFERENCE) is single-GPU only"); return QR_EINVAL; }
    QR_TRY(comm_create(comm_id, rank, world, &c->comm));
    unsigned long long nq[2] = {(unsigned long long) N, (unsigned long long) Q};
    unsigned long long *d_nq = nullptr;
    QR_TRY(dev_alloc(&d_nq, 2));
    QR_CUDA(cudaMemcpyAsync(d_nq, nq, sizeof(nq), cudaMemcpyHostToDevice, c->stream));
    QR_TRY(comm_allreduce_sum_u64(c->comm, d_nq, 2, c->stream));
    QR_CUDA(cudaMemcpyAsync(nq, d_nq, sizeof(nq), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_nq);
    c->N_global = (size_t) nq[0];
    c->Q_global = (size_t) nq[1];
    // the largest shard: histogram slices are laid out from quantities every rank knows, so that all ranks
    // use the same layout (the fused exchange reads the peers' per-slice squares partials)
    unsigned long long nmax = (unsigned long long) N, *d_nmax = nullptr;
    QR_TRY(dev_alloc(&d_nmax, 1));
    QR_CUDA(cudaMemcpyAsync(d_nmax, &nmax, sizeof(nmax), cudaMemcpyHostToDevice, c->stream));
    QR_TRY(comm_allreduce_max_u64(c->comm, d_nmax, 1, c->stream));
    QR_CUDA(cudaMemcpyAsync(&nmax, d_nmax, sizeof(nmax), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_nmax);
    c->N_local_max = (size_t) nmax;
  }

  if (gather_src) {   // a document sample cut from the bins of `thr_from`, sized for all of its documents
    c->eval_only = false;
    c->cap_N = thr_from->N; c->cap_Q = thr_from->Q; c->cap_maxlen = thr_from->maxlen;
    QR_TRY(gather_binning(c, thr_from, gather_src));
    clk.lap("bins gathered from the sampled context");
    return ctx_create_state(c, labels, qoffsets, params, clk);
  }
  // features -> device column-major (VerticalDataset layout), then bins; floats are released
  float *d_col = nullptr;
  QR_TRY(dev_alloc(&d_col, N * F));
  if (rowmajor) {
    float *d_row = nullptr;
    QR_TRY(dev_alloc(&d_row, N * F));
    QR_CUDA(cudaMemcpyAsync(d_row, feat, N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned) ((N + 31) / 32), (unsigned) ((F + 31) / 32));   // documents on x: grid.y stops at 65535
    transpose_kernel<<<grid, dim3(32, 8), 0, c->stream>>>(d_row, d_col, N, F);
    QR_CUDA(cudaGetLastError());
    QR_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_row);
  } else {
    QR_CUDA(cudaMemcpyAsync(d_col, feat, N * F * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  c->eval_only = thr_from != nullptr && !sample;   // (a document sample is binned with the thresholds of the whole set and trained on)
  clk.lap("feature upload (+transpose)");
  int rc = build_binning(c, d_col, thr_from);
  cudaFree(d_col);
  if (rc != QR_OK) return rc;
  clk.lap("thresholds + binning");
  return ctx_create_state(c, labels, qoffsets, params, clk);
}

// Contents of the per-document / per-query tables for the context's current N and Q: labels, gains, query offsets,
// ideal DCG per query (host glibc), zeroed state arrays (mart.cc:121-122, lambdamart.cc:38).  The buffers exist and hold
// at least cap_N / cap_Q entries.  Used by context creation and by qr_sample_redraw.
static int fill_query_tables(qr_ctx *c, const float *labels, const std::vector<uint32_t> &qoff, const std::vector<double> &lg) {
  const size_t N = c->N, Q = c->Q;
  std::vector<double> gain(N), idcg(Q);
  // Queries are independent: a few host threads each take a range of them (and of their documents).  pow(2, label)
  // (dcg.cc:37) goes through a per-thread memo of the few distinct label values: the same glibc results without a
  // pow() call per document.
  const size_t cutoff = c->cutoff;
  auto fill = [&](size_t q0, size_t q1) {
    float memo_label[8];
    double memo_gain[8];
    int memo_n = 0;
    auto gain_of = [&](float label) {
      for (int k = 0; k < memo_n; ++k)
        if (memo_label[k] == label) return memo_gain[k];
      const double g = std::pow(2.0, (double) label);
      if (memo_n < 8) { memo_label[memo_n] = label; memo_gain[memo_n] = g; ++memo_n; }
      return g;
    };
    for (size_t i = qoff[q0]; i < qoff[q1]; ++i) gain[i] = gain_of(labels[i]);
    std::vector<float> tmp;
    for (size_t q = q0; q < q1; ++q) {                                                       // ndcg.cc:35-47
      tmp.assign(labels + qoff[q], labels + qoff[q + 1]);
      std::sort(tmp.begin(), tmp.end(), std::greater<int>());
      const size_t size = std::min(cutoff, tmp.size());
      double dcg = 0.0;
      for (size_t i = 0; i < size; ++i) dcg += (gain_of(tmp[i]) - 1.0) / lg[i];
      idcg[q] = dcg;
    }
  };
  const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t) std::thread::hardware_concurrency(), (size_t) 16, N / 50000 + 1, Q}));
  if (nthreads <= 1) {
    fill(0, Q);
  } else {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; ++t) pool.emplace_back(fill, Q * t / nthreads, Q * (t + 1) / nthreads);
    for (auto &t : pool) t.join();
  }
  QR_CUDA(cudaMemcpy(c->d_labels, labels, N * sizeof(float), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_gain, gain.data(), N * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_qoff, qoff.data(), (Q + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_idcg, idcg.data(), Q * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemset(c->d_scores, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_lambda, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_weight, 0, N * sizeof(double)));
  QR_CUDA(cudaMemset(c->d_leaf_of_doc, 0, N * sizeof(uint32_t)));
  QR_CUDA(cudaMemset(c->d_qexp, 0, 2 * sizeof(int)));
  return QR_OK;
}

// query offsets as 32-bit device offsets, and the longest query
static int check_query_offsets(const uint64_t *qoffsets, size_t Q, std::vector<uint32_t> &qoff, uint32_t *maxlen_out) {
  qoff.resize(Q + 1);
  uint32_t maxlen = 0;
  for (size_t q = 0; q <= Q; ++q) {
    if (q && qoffsets[q] < qoffsets[q - 1]) { set_error("query offsets must be non-decreasing"); return QR_EINVAL; }
    qoff[q] = (uint32_t) qoffsets[q];
    if (q) maxlen = std::max<uint32_t>(maxlen, qoff[q] - qoff[q - 1]);
  }
  *maxlen_out = maxlen;
  return QR_OK;
}

// discount tables (host glibc): lg[i] = log2((float) i + 2) (dcg.cc:37), invlg[i] = 1 / log2(i + 2) (ndcg.cc:79)
static void discount_tables(uint32_t n, std::vector<double> &lg, std::vector<double> &invlg) {
  lg.resize(n + 1);
  invlg.resize(n + 1);
  for (uint32_t i = 0; i <= n; ++i) {
    lg[i] = std::log2((double) ((float) i + 2.0f));
    invlg[i] = 1.0 / std::log2((double) (i + 2));
  }
}

// second half of context creation: per-query tables, state arrays, histogram pool, kernel attributes, peers
static int ctx_create_state(qr_ctx *c, const float *labels, const uint64_t *qoffsets, const qr_params *params, InitClock &clk) {
  const size_t F = c->F;
  std::vector<uint32_t> qoff;
  uint32_t maxlen = 0;
  QR_TRY(check_query_offsets(qoffsets, c->Q, qoff, &maxlen));
  c->maxlen = (maxlen + 3u) & ~3u;
  // everything below that scales with the documents / queries / longest query is sized by these (see qr_internal.cuh)
  c->cap_N = std::max(c->cap_N, c->N);
  c->cap_Q = std::max(c->cap_Q, c->Q);
  c->cap_maxlen = std::max(c->cap_maxlen, c->maxlen);
  const size_t N = c->cap_N, Q = c->cap_Q;
  std::vector<double> lg, invlg;
  discount_tables(c->cap_maxlen, lg, invlg);
  QR_TRY(dev_alloc(&c->d_labels, N));
  QR_TRY(dev_alloc(&c->d_gain, N));
  QR_TRY(dev_alloc(&c->d_qoff, Q + 1));
  QR_TRY(dev_alloc(&c->d_idcg, Q));
  QR_TRY(dev_alloc(&c->d_lg, c->cap_maxlen + 1));
  QR_TRY(dev_alloc(&c->d_invlg, c->cap_maxlen + 1));
  QR_CUDA(cudaMemcpy(c->d_lg, lg.data(), (c->cap_maxlen + 1) * sizeof(double), cudaMemcpyHostToDevice));
  QR_CUDA(cudaMemcpy(c->d_invlg, invlg.data(), (c->cap_maxlen + 1) * sizeof(double), cudaMemcpyHostToDevice));
  QR_TRY(dev_alloc(&c->d_scores, N));
  QR_TRY(dev_alloc(&c->d_lambda, N));
  QR_TRY(dev_alloc(&c->d_weight, N));
  QR_TRY(dev_alloc(&c->d_lamq, N));
  QR_TRY(dev_alloc(&c->d_maxabs, 2));
  QR_TRY(dev_alloc(&c->d_qexp, 2));
  QR_TRY(dev_alloc(&c->d_rankpos, N));
  QR_TRY(dev_alloc(&c->d_qndcg, Q));
  QR_TRY(dev_alloc(&c->d_metric, 1));
  QR_TRY(dev_alloc(&c->d_leaf_of_doc, N));
  QR_TRY(fill_query_tables(c, labels, qoff, lg));
  clk.lap("per-query tables + state arrays");

  const size_t maxleaves = c->oblivious ? ((size_t) 1 << params->treedepth) : std::max<size_t>(params->nleaves, 1);
  c->max_tasks = (uint32_t) maxleaves + 1;
  if (c->eval_only) {   // no tree growth on this context: no histogram pool
    QR_TRY(dev_alloc(&c->d_leafval, maxleaves + 1));
    QR_CUDA(cudaGetLastError());
    return QR_OK;
  }
  // every expansion holds two child histograms until the node is popped or the tree is finished;
  // speculative expansions (qr_tree_host.cuh) can double the number of live nodes
  c->nslots = (int) (4 * maxleaves + 8);
  // node ids (one per expansion child) are 16 bits in node_of_doc
  const size_t max_nodes = c->oblivious ? ((size_t) 2 << params->treedepth) : 4 * maxleaves + 16;
  if (!c->exact && max_nodes > 65535) {
    set_error("trees of more than %s are not supported by the fixed-point growth path (node ids are 16 bits)",
              c->oblivious ? "depth 14" : "16379 leaves");
    return QR_ELIMIT;
  }
  c->max_nodes = (uint32_t) max_nodes;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  // behind the pool: sharded training, two sets of staging slots (local histograms of a round's built
  // children); one GPU, fixed-point mode: one raw slot per task of a round (kept clear by the fused split scan)
  c->stage_slot0 = c->nslots;
  const size_t hist_smem_bytes = (size_t) c->fpp * c->max_thr * 12;
  (void) hist_smem_bytes;
  c->fused_scan = !c->exact && !c->comm && F <= 65535;   // one GPU: scan_pub_kernel (its records pack the feature in 16 bits)
  const size_t pool_slots = (size_t) c->nslots + (c->comm ? 2 * (maxleaves + 1) : (c->fused_scan ? maxleaves + 1 : 0));
  const size_t hist_bytes = pool_slots * c->ncells * 12;
  if (hist_bytes + (64u << 20) > free_b) {
    set_error("histogram pool needs %zu MB (%d nodes x %u cells); not enough device memory — bound the "
              "bin count with --num-thresholds", hist_bytes >> 20, (int) pool_slots, c->ncells);
    return QR_ENOMEM;
  }
  // sharded training exports the two pools over CUDA IPC: at least 2 MB each, so that each is an allocation
  // of its own and not a piece of a page shared with other buffers
  const size_t pool_cells = std::max<size_t>(pool_slots * c->ncells, c->comm ? ((size_t) 2 << 20) / 4 : 1);
  QR_TRY(dev_alloc(&c->d_hist_sum, pool_cells));
  QR_TRY(dev_alloc(&c->d_hist_cnt, pool_cells));
  for (int i = c->nslots - 1; i >= 0; --i) c->free_slots.push_back(i);
  const size_t mt = c->max_tasks;
  if (c->fused_scan) {
    QR_CUDA(cudaMemset(c->d_hist_sum + (size_t) c->nslots * c->ncells, 0, mt * c->ncells * sizeof(unsigned long long)));
    QR_CUDA(cudaMemset(c->d_hist_cnt + (size_t) c->nslots * c->ncells, 0, mt * c->ncells * sizeof(uint32_t)));
    QR_TRY(dev_alloc(&c->d_sq_acc, mt));
    QR_CUDA(cudaMemset(c->d_sq_acc, 0, mt * sizeof(ulonglong2)));
  }
  c->pub_ok = !c->exact && F <= 65535;   // scan_pub_kernel's records pack the feature in 16 bits
  if (c->pub_ok) {
    QR_CUDA(cudaHostAlloc((void **) &c->h_out, mt * 2 * sizeof(ChildOut), cudaHostAllocMapped));
    QR_CUDA(cudaHostGetDevicePointer((void **) &c->d_out_mapped, c->h_out, 0));
    memset(c->h_out, 0, mt * 2 * sizeof(ChildOut));   // (tag 0 = not written)
    QR_TRY(dev_alloc(&c->d_cand, mt * 2 * F));
    QR_TRY(dev_alloc(&c->d_noderec, mt * 2));
  }
  QR_TRY(dev_alloc(&c->d_partials, mt));
  QR_TRY(dev_alloc(&c->d_sq_built, mt));
  c->max_slices = 4096 + (uint32_t) mt;
  QR_TRY(dev_alloc(&c->d_sq128, std::max<size_t>(c->max_slices, c->comm ? ((size_t) 2 << 20) / sizeof(ulonglong2) : 1)));   // (exported over IPC)
  QR_TRY(dev_alloc(&c->d_task_done, mt));
  QR_CUDA(cudaMemset(c->d_task_done, 0, mt * sizeof(uint32_t)));
  QR_TRY(dev_alloc(&c->d_tasks, mt));
  QR_CUDA(cudaMallocHost((void **) &c->h_tasks, mt * sizeof(NodeTask)));
  QR_TRY(dev_alloc(&c->d_lcount, mt));
  if (const char *e = getenv("QR_PEER_FUSED")) c->peer_fused = atoi(e) != 0;
  if (const char *e = getenv("QR_PEER_ONESHOT_MAX")) c->oneshot_max = (uint32_t) std::max(0, atoi(e));
  QR_TRY(dev_alloc(&c->d_fbest_score, mt * 2 * F));
  QR_TRY(dev_alloc(&c->d_fbest_t, mt * 2 * F));
  QR_TRY(dev_alloc(&c->d_fbest_lc, mt * 2 * F));
  QR_TRY(dev_alloc(&c->d_totals, mt * 2));
  QR_TRY(dev_alloc(&c->d_res, mt * 2));
  QR_CUDA(cudaHostAlloc((void **) &c->h_res, mt * 2 * sizeof(SplitResult), cudaHostAllocMapped));
  QR_CUDA(cudaHostGetDevicePointer((void **) &c->d_res_mapped, c->h_res, 0));
  QR_CUDA(cudaHostAlloc((void **) &c->h_flags, mt * sizeof(uint32_t), cudaHostAllocMapped));
  QR_CUDA(cudaHostGetDevicePointer((void **) &c->d_flags_mapped, c->h_flags, 0));
  memset(c->h_flags, 0, mt * sizeof(uint32_t));
  QR_CUDA(cudaHostAlloc((void **) &c->h_err, sizeof(uint32_t), cudaHostAllocMapped));
  QR_CUDA(cudaHostGetDevicePointer((void **) &c->d_err_mapped, c->h_err, 0));
  *c->h_err = 0u;
  QR_TRY(dev_alloc(&c->d_leafsum, maxleaves + 1));
  QR_TRY(dev_alloc(&c->d_leafval, maxleaves + 1));
  QR_CUDA(cudaMallocHost((void **) &c->h_leafval, (maxleaves + 1) * sizeof(double)));
  QR_TRY(dev_alloc(&c->d_obv_scores, c->ncells));
  QR_TRY(dev_alloc(&c->d_obv_slots, mt));
  QR_TRY(dev_alloc(&c->d_obv_lcounts, mt));
  QR_CUDA(cudaMallocHost((void **) &c->h_obv_lcounts, mt * sizeof(uint64_t)));
  if (c->exact) {
    // REFERENCE mode: sample-id lists cut by a stable partition, leaves summed in list order
    QR_TRY(dev_alloc(&c->d_ids[0], N));
    QR_TRY(dev_alloc(&c->d_ids[1], N));
    QR_TRY(dev_alloc(&c->d_blockcnt, (N + kPartItems - 1) / kPartItems + mt + 1));
    QR_TRY(dev_alloc(&c->d_segs, maxleaves + 1));
    QR_CUDA(cudaMallocHost((void **) &c->h_segs, (maxleaves + 1) * sizeof(LeafSeg)));
    QR_TRY(build_exact_walk_tables(c));
  } else {
    // FAST mode: node_of_doc + the compact lists of a round's built children.  One GPU: the built child is the
    // smaller one, so a round appends at most N/2 documents.  Several ranks: a task's region must hold whatever
    // share of the (globally smaller) child is local: min(global size, N) each, N_global/2 + ... in total.
    QR_TRY(dev_alloc(&c->d_node, N + 8));
    QR_CUDA(cudaMemset(c->d_node, 0, (N + 8) * sizeof(uint16_t)));
    c->compact_cap = (c->comm ? c->N_global / 2 : N / 2) + 8;
    QR_TRY(dev_alloc(&c->d_cids, c->compact_cap));
    QR_TRY(dev_alloc(&c->d_clamq, c->compact_cap));
    QR_TRY(dev_alloc(&c->d_counts, 2 * mt));
    QR_CUDA(cudaMemset(c->d_counts, 0, 2 * mt * sizeof(uint32_t)));
    c->leafn_off = ((size_t) c->max_nodes * 2 + 15) & ~(size_t) 15;
    const size_t meta = c->leafn_off + (maxleaves + 1) * sizeof(unsigned long long);
    QR_CUDA(cudaMallocHost((void **) &c->h_leafmeta, meta));
    QR_TRY(dev_alloc(&c->d_leafmeta, meta));
    QR_CUDA(cudaFuncSetAttribute(route_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
    QR_CUDA(cudaFuncSetAttribute(route_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 132 * 1024));
  }

  // opt in to large dynamic shared memory where needed
  const size_t hist_smem = (size_t) c->fpp * c->max_thr * 12;
  if (hist_smem <= 200 * 1024) {
    const int bytes = (int) hist_smem;
    cudaFuncSetAttribute(hist_limb_kernel<uint8_t, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint16_t, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint8_t, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint16_t, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint8_t, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint16_t, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint8_t, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(hist_limb_kernel<uint16_t, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  }
  QR_CUDA(cudaGetLastError());
  clk.lap("state + pools");
  QR_TRY(comm_setup_peers(c));
  {
    const char *e = getenv("QR_PEER_SLICED");
    c->sliced = c->comm && comm_transport(c->comm) == 2 && c->pub_ok && c->peer_fused && !c->oblivious &&
                F >= (size_t) comm_world(c->comm) && comm_world(c->comm) > 1 && (e == nullptr || atoi(e) != 0);
  }
  QR_TRY(init_root_counts(c));
  QR_CUDA(cudaGetLastError());
  clk.lap("root counts");
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
  const size_t smem = (size_t) kRankWarps * rank_smem_per_warp(c->maxlen);
  if (smem > 200 * 1024) { set_error("longest query (%u documents) exceeds the ranking kernel's shared-memory budget", c->maxlen); return QR_ELIMIT; }
  static bool attr_set = false;
  if (!attr_set || smem > 48 * 1024) {
    cudaFuncSetAttribute(rank_kernel<kRankWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
    attr_set = true;
  }
  const unsigned grid = (unsigned) ((c->Q + kRankWarps - 1) / kRankWarps);
  // (a document sample ranks by its gathered keys: lambdamart.cc:94, see qr_ctx_create_sample)
  QR_LAUNCH(c, PH_RANK, rank_kernel<kRankWarps>, grid, kRankWarps * 32, smem,
            (const double *) (c->d_rankkey ? c->d_rankkey : c->d_scores), c->d_labels,
            c->d_gain, c->d_qoff, c->d_idcg, c->d_lg, (uint32_t) c->Q, c->maxlen, c->cutoff,
            c->d_rankpos, c->d_qndcg);
  c->ranking_valid = true;
  return QR_OK;
}

// NDCG@k of `nvec` score vectors over this context's documents (scores_dev[v * N + doc]) in one ranking launch and
// one launch of sequential means: the line search's candidates (qr_linesearch.cu).  REFERENCE-mode contexts only.
int evaluate_vectors(qr_ctx *c, const double *scores_dev, uint32_t nvec, double *metrics_host) {
  if (!c->exact || nvec == 0) { set_error("internal: evaluate_vectors needs a REFERENCE-mode context"); return QR_EINVAL; }
  const size_t smem = (size_t) kRankWarps * rank_smem_per_warp(c->maxlen);
  if (smem > 200 * 1024) { set_error("longest query (%u documents) exceeds the ranking kernel's shared-memory budget", c->maxlen); return QR_ELIMIT; }
  cudaFuncSetAttribute(rank_kernel<kRankWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
  if (c->vec_cap < nvec) {
    if (c->d_vec_qndcg) cudaFree(c->d_vec_qndcg);
    if (c->d_vec_metric) cudaFree(c->d_vec_metric);
    c->d_vec_qndcg = nullptr; c->d_vec_metric = nullptr;
    c->vec_cap = std::max<uint32_t>(nvec, 32);
    QR_TRY(dev_alloc(&c->d_vec_qndcg, (size_t) c->vec_cap * c->Q));
    QR_TRY(dev_alloc(&c->d_vec_metric, c->vec_cap));
  }
  const dim3 grid((unsigned) ((c->Q + kRankWarps - 1) / kRankWarps), nvec);
  // (a one-vector grid takes the kernel's single-vector path, which also writes rankpos: the cached ranking is dropped)
  QR_LAUNCH(c, PH_RANK, rank_kernel<kRankWarps>, grid, kRankWarps * 32, smem, scores_dev, c->d_labels, c->d_gain, c->d_qoff,
            c->d_idcg, c->d_lg, (uint32_t) c->Q, c->maxlen, c->cutoff, c->d_rankpos, c->d_vec_qndcg);
  c->ranking_valid = false;
  QR_LAUNCH(c, PH_RANK, ndcg_mean_kernel, nvec, 32, 0, c->d_vec_qndcg, (uint32_t) c->Q, (uint32_t) c->Q, true, 0, c->d_vec_metric);
  QR_CUDA(cudaMemcpyAsync(metrics_host, c->d_vec_metric, nvec * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
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
  const uint32_t hg = (uint32_t) std::min<size_t>(kHG, std::max<size_t>(1, std::min<size_t>(c->cutoff, c->maxlen)));
  const size_t smem = (size_t) kLambdaWarps * lambda_smem_per_warp(c->maxlen, hg);
  if (smem > 200 * 1024) { set_error("longest query (%u documents) exceeds the lambda kernel's shared-memory budget", c->maxlen); return QR_ELIMIT; }
  cudaFuncSetAttribute(lambda_kernel<kLambdaWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024));
  const unsigned grid = (unsigned) ((c->Q + kLambdaWarps - 1) / kLambdaWarps);
  QR_LAUNCH(c, PH_PSEUDO, lambda_kernel<kLambdaWarps>, grid, kLambdaWarps * 32, smem, c->d_scores,
            c->d_labels, c->d_gain, c->d_qoff, c->d_idcg, c->d_invlg, c->d_rankpos, (uint32_t) c->Q,
            c->maxlen, c->cutoff, hg, c->d_lambda, c->d_weight);
  return QR_OK;
}

static int evaluate(qr_ctx *c, double *metric) {
  QR_TRY(ensure_ranking(c));
  // FAST: the per-query values are added as fixed-point integers (exact, so the mean does not depend on how the
  // queries are spread over blocks or GPUs); REFERENCE: the sequential sum in query order (metric.h:96-105)
  const int qshift = 62 - (ceil_log2(c->Q_global ? c->Q_global : c->Q) + 1);
  {
    PhaseTimer pt(c, PH_RANK);
    QR_LAUNCH(c, PH_RANK, ndcg_mean_kernel, 1, c->exact ? 32 : 1024, 0, c->d_qndcg, (uint32_t) c->Q, (uint32_t) c->Q,
              c->exact, qshift, c->d_metric);
    if (c->comm) QR_TRY(comm_allreduce_sum_u64(c->comm, reinterpret_cast<unsigned long long *>(c->d_metric), 1, c->stream));
  }
  double m = 0;
  QR_CUDA(cudaMemcpyAsync(&m, c->d_metric, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  if (!c->exact) {
    long long q;
    memcpy(&q, &m, sizeof(q));
    const size_t nq = c->comm ? c->Q_global : c->Q;
    m = nq ? std::ldexp((double) q, -qshift) / (double) nq : 0.0;
  }
  if (metric) *metric = m;
  return QR_OK;
}

#include "qr_tree_host.cuh"

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
  int rc = ctx_create_common(feat, false, N, F, labels, qoff, Q, params, nullptr, 0, 1, nullptr, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_ctx_create_rowmajor(const float *feat, size_t N, size_t F, const float *labels,
                           const uint64_t *qoff, size_t Q, const qr_params *params, qr_ctx **out) {
  if (out) *out = nullptr;
  int rc = ctx_create_common(feat, true, N, F, labels, qoff, Q, params, nullptr, 0, 1, nullptr, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_ctx_create_sharded(const float *feat, int rowmajor, size_t N, size_t F, const float *labels,
                          const uint64_t *qoff, size_t Q, const qr_params *params,
                          const unsigned char id[QR_COMM_ID_BYTES], int rank, int world, qr_ctx **out) {
  if (out) *out = nullptr;
  if (!id) { set_error("qr_ctx_create_sharded: null communicator id"); return QR_EINVAL; }
  int rc = ctx_create_common(feat, rowmajor != 0, N, F, labels, qoff, Q, params, id, rank, world, nullptr, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_ctx_comm_transport(const qr_ctx *c) { return c ? comm_transport(c->comm) : 0; }

int qr_ctx_create_eval(qr_ctx *train, const float *feat_rowmajor, size_t N, size_t F, const float *labels,
                       const uint64_t *qoff, size_t Q, qr_ctx **out) {
  if (out) *out = nullptr;
  if (!train) { set_error("qr_ctx_create_eval: null training context"); return QR_EINVAL; }
  if (F != train->F) { set_error("dataset has %zu features, the training set has %zu", F, train->F); return QR_EINVAL; }
  qr_params p = train->p;
  p.device = train->device;
  int rc = ctx_create_common(feat_rowmajor, true, N, F, labels, qoff, Q, &p, nullptr, 0, 1, train, out);
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

// A document sample of `full` as a training context of its own (LambdaMartSelective::learn,
// lambdamartselective.cc:185-206: pseudo-responses, root histogram and tree fit run over `sampleids` only).
int qr_ctx_create_sample(qr_ctx *full, const float *feat_rowmajor, size_t N, size_t F, const float *labels,
                         const uint64_t *qoff, size_t Q, const uint32_t *src_doc, const uint32_t *key_doc, qr_ctx **out) {
  if (out) *out = nullptr;
  if (!full || !src_doc) { set_error("qr_ctx_create_sample: null argument"); return QR_EINVAL; }
  if (full->eval_only || full->comm) { set_error("qr_ctx_create_sample: the sampled context must be a single-GPU training context"); return QR_EINVAL; }
  if (F != full->F) { set_error("the sample has %zu features, the training set has %zu", F, full->F); return QR_EINVAL; }
  for (size_t i = 0; i < N; ++i)
    if (src_doc[i] >= full->N || (key_doc && key_doc[i] >= full->N)) {
      set_error("qr_ctx_create_sample: document index out of range at %zu", i);
      return QR_EINVAL;
    }
  qr_params p = full->p;
  p.device = full->device;
  p.hist_mode = QR_HIST_FAST;   // (sums over a sample have no reference order to follow: fixed point)
  // feat_rowmajor == NULL: the sample's bins are gathered on the device from the panels of `full`
  int rc = ctx_create_common(feat_rowmajor, true, N, F, labels, qoff, Q, &p, nullptr, 0, 1, full, out, true,
                             feat_rowmajor ? nullptr : src_doc);
  if (rc == QR_OK) {
    qr_ctx *c = *out;
    c->sample_of_N = full->N;
    if (!c->d_src_doc) {
      rc = dev_alloc(&c->d_src_doc, N);
      if (rc == QR_OK && cudaMemcpy(c->d_src_doc, src_doc, N * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = QR_ECUDA;
    }
    if (rc == QR_OK && key_doc) {
      rc = dev_alloc(&c->d_key_doc, std::max(c->cap_N, N));
      if (rc == QR_OK) rc = dev_alloc(&c->d_rankkey, std::max(c->cap_N, N));
      if (rc == QR_OK && cudaMemcpy(c->d_key_doc, key_doc, N * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) rc = QR_ECUDA;
    }
    if (rc == QR_ECUDA) set_error("qr_ctx_create_sample: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc != QR_OK && out && *out) { std::string keep = g_last_error; qr_ctx_destroy(*out); *out = nullptr; g_last_error = keep; }
  return rc;
}

int qr_sample_redraw(qr_ctx *c, qr_ctx *full, size_t N, const float *labels, const uint64_t *qoffsets, size_t Q,
                     const uint32_t *src_doc, const uint32_t *key_doc) {
  if (!c || !full || !labels || !qoffsets || !src_doc || N == 0 || Q == 0) { set_error("qr_sample_redraw: null or empty argument"); return QR_EINVAL; }
  if (!c->d_src_doc || c->sample_of_N != full->N || c->device != full->device || c->F != full->F) {
    set_error("qr_sample_redraw: not a sample of this context");
    return QR_EINVAL;
  }
  if ((key_doc != nullptr) != (c->d_key_doc != nullptr)) { set_error("qr_sample_redraw: the sample was created %s ranking keys", c->d_key_doc ? "with" : "without"); return QR_EINVAL; }
  if (N > c->cap_N || Q > c->cap_Q || qoffsets[0] != 0 || qoffsets[Q] != N) {
    set_error("qr_sample_redraw: %zu documents / %zu queries do not fit a sample context created for %zu / %zu "
              "(create it with feat_rowmajor == NULL), or bad query offsets", N, Q, c->cap_N, c->cap_Q);
    return QR_EINVAL;
  }
  for (size_t i = 0; i < N; ++i)
    if (src_doc[i] >= full->N || (key_doc && key_doc[i] >= full->N)) { set_error("qr_sample_redraw: document index out of range at %zu", i); return QR_EINVAL; }
  std::vector<uint32_t> qoff;
  uint32_t maxlen = 0;
  QR_TRY(check_query_offsets(qoffsets, Q, qoff, &maxlen));
  if (((maxlen + 3u) & ~3u) > c->cap_maxlen) { set_error("qr_sample_redraw: a query of %u documents exceeds the longest query of the sampled context", maxlen); return QR_EINVAL; }
  cudaSetDevice(c->device);
  QR_CUDA(cudaStreamSynchronize(c->stream));
  // the last tree of the previous sample goes: its nodes' histogram slots, its document -> node map
  for (auto &nd : c->nodes) release_slot(c, nd.hist);
  c->nodes.clear();
  c->leaves.clear();
  c->has_tree = false;
  c->ranking_valid = false;
  c->N = c->N_global = N;
  c->Q = c->Q_global = Q;
  c->maxlen = (maxlen + 3u) & ~3u;
  QR_TRY(gather_sample_bins(c, full, src_doc));
  if (key_doc) QR_CUDA(cudaMemcpy(c->d_key_doc, key_doc, N * sizeof(uint32_t), cudaMemcpyHostToDevice));
  std::vector<double> lg, invlg;
  discount_tables(c->maxlen, lg, invlg);
  QR_TRY(fill_query_tables(c, labels, qoff, lg));
  QR_CUDA(cudaMemset(c->d_node, 0, (c->cap_N + 8) * sizeof(uint16_t)));
  QR_TRY(init_root_counts(c));
  return QR_OK;
}

int qr_sample_pull_scores(qr_ctx *c, qr_ctx *full) {
  if (!c || !full || !c->d_src_doc || c->sample_of_N != full->N || c->device != full->device) {
    set_error("qr_sample_pull_scores: not a sample of this context");
    return QR_EINVAL;
  }
  cudaSetDevice(c->device);
  QR_CUDA(cudaStreamSynchronize(full->stream));   // (the two contexts run on streams of their own)
  QR_LAUNCH(c, PH_RANK, sample_gather_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, (const double *) full->d_scores,
            (const uint32_t *) c->d_src_doc, (const uint32_t *) c->d_key_doc, c->N, c->d_scores, c->d_rankkey);
  c->ranking_valid = false;
  return QR_OK;
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
                  c->d_leafval, c->d_obv_scores, c->d_tasks, c->d_lcount, c->d_segs, c->d_leaf_partials,
                  c->d_leafsum, c->d_obv_slots, c->d_obv_lcounts, c->d_sq128, c->d_task_done,
                  c->d_root_cnt, c->d_fbest_lc, c->d_totals, c->d_node, c->d_cids, c->d_clamq, c->d_counts,
                  c->d_sq_built, c->d_leafmeta, c->d_sq_acc, c->d_cand, c->d_noderec, c->d_kspan, c->d_rows,
                  c->d_perm, c->d_cell_pos, c->d_mark, c->d_sq_chunks, c->d_sq_replayed, c->d_fskip,
                  c->d_vec_qndcg, c->d_vec_metric, c->d_src_doc, c->d_key_doc, c->d_rankkey};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (c->h_res) cudaFreeHost(c->h_res);
  if (c->h_leafval) cudaFreeHost(c->h_leafval);
  if (c->h_tasks) cudaFreeHost(c->h_tasks);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->h_segs) cudaFreeHost(c->h_segs);
  if (c->h_obv_lcounts) cudaFreeHost(c->h_obv_lcounts);
  if (c->h_leafmeta) cudaFreeHost(c->h_leafmeta);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->d_apply) cudaFree(c->d_apply);
  if (c->h_apply) cudaFreeHost(c->h_apply);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_k0) cudaEventDestroy(c->ev_k0);
  if (c->ev_k1) cudaEventDestroy(c->ev_k1);
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

#define QR_CHECK_TRAINABLE(c)                                                                        \
  do {                                                                                              \
    if ((c)->eval_only) { set_error("this is an evaluation-only context (qr_ctx_create_eval)"); return QR_EINVAL; } \
  } while (0)

int qr_compute_pseudoresponses(qr_ctx *c) {
  QR_CHECK_CTX(c);
  QR_CHECK_TRAINABLE(c);
  return compute_pseudo(c);
}

int qr_fit_tree(qr_ctx *c, qr_flat_tree *out) {
  QR_CHECK_CTX(c);
  QR_CHECK_TRAINABLE(c);
  return fit_tree(c, out);
}

int qr_update_modelscores(qr_ctx *c, double weight) {
  QR_CHECK_CTX(c);
  return update_modelscores(c, weight);
}

int qr_apply_trees(qr_ctx *c, const qr_flat_tree *trees, const double *weights, size_t ntrees);

// one tree = a set of one (the staging buffers of qr_apply_trees are reused: no allocation and no stream
// synchronisation per call, the validation set takes one of these per boosting iteration)
int qr_apply_tree(qr_ctx *c, const qr_flat_tree *t, double weight) {
  QR_CHECK_CTX(c);
  if (!t || t->nnodes == 0) { set_error("qr_apply_tree: empty tree"); return QR_EINVAL; }
  return qr_apply_trees(c, t, &weight, 1);
}

// packs `ntrees` trees (and their weights, when given) into the context's staging buffer and copies it to the device:
// weights | nodes | roots
struct StagedTrees { const double *w; const qr::PackedNode *nodes; const uint32_t *roots; size_t total; };
static int stage_trees(qr_ctx *c, const char *who, const qr_flat_tree *trees, const double *weights, size_t ntrees, StagedTrees *st) {
  size_t total = 0;
  for (size_t t = 0; t < ntrees; ++t) {
    const qr_flat_tree &ft = trees[t];
    if (ft.nnodes == 0) { set_error("%s: empty tree %zu", who, t); return QR_EINVAL; }
    for (uint32_t i = 0; i < ft.nnodes; ++i)
      if (ft.feature[i] >= 0 && ((size_t) ft.feature[i] >= c->F || ft.threshold_idx[i] >= c->thr[ft.feature[i]].size())) {
        set_error("%s: node %u of tree %zu does not belong to this context's binning", who, i, t);
        return QR_EINVAL;
      }
    total += ft.nnodes;
  }
  const size_t bytes = total * sizeof(PackedNode) + ntrees * sizeof(uint32_t) + ntrees * sizeof(double) + 16;
  if (bytes > c->apply_cap) {
    if (c->d_apply) cudaFree(c->d_apply);
    if (c->h_apply) cudaFreeHost(c->h_apply);
    c->apply_cap = std::max<size_t>(bytes * 2, 1 << 16);
    QR_CUDA(cudaMalloc(&c->d_apply, c->apply_cap));
    QR_CUDA(cudaMallocHost(&c->h_apply, c->apply_cap));
  } else {
    QR_CUDA(cudaStreamSynchronize(c->stream));   // the previous pass may still be reading the staging copy
  }
  unsigned char *h = (unsigned char *) c->h_apply;
  double *hw = (double *) h;                                           // weights first (8-byte aligned)
  PackedNode *hn = (PackedNode *) (h + ntrees * sizeof(double));
  uint32_t *hr = (uint32_t *) (hn + total);
  size_t o = 0;
  for (size_t t = 0; t < ntrees; ++t) {
    const qr_flat_tree &ft = trees[t];
    hr[t] = (uint32_t) o;
    hw[t] = weights ? weights[t] : 1.0;
    for (uint32_t i = 0; i < ft.nnodes; ++i)
      hn[o + i] = PackedNode{ft.feature[i], ft.feature[i] >= 0 ? ft.threshold_idx[i] : 0u, ft.left[i], ft.right[i], ft.value[i]};
    o += ft.nnodes;
  }
  QR_CUDA(cudaMemcpyAsync(c->d_apply, c->h_apply, bytes, cudaMemcpyHostToDevice, c->stream));
  unsigned char *d = (unsigned char *) c->d_apply;
  st->w = (const double *) d;
  st->nodes = (const PackedNode *) (d + ntrees * sizeof(double));
  st->roots = (const uint32_t *) (st->nodes + total);
  st->total = total;
  return QR_OK;
}

int qr_apply_trees(qr_ctx *c, const qr_flat_tree *trees, const double *weights, size_t ntrees) {
  QR_CHECK_CTX(c);
  if (ntrees == 0) return QR_OK;
  if (!trees || !weights) { set_error("qr_apply_trees: null argument"); return QR_EINVAL; }
  StagedTrees st{};
  QR_TRY(stage_trees(c, "qr_apply_trees", trees, weights, ntrees, &st));
  int rc = dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    QR_LAUNCH(c, PH_LEAF, apply_trees_kernel<B>, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_panels, c->N, st.nodes, st.roots, st.w,
              (uint32_t) ntrees, c->d_scores);
    return QR_OK;
  });
  c->ranking_valid = false;
  return rc;
}

// Dart::update_contribution_scores (dart.cc:689-706) for `ntrees` trees in one pass over the documents:
// contribution[t] = mean over the (global) dataset of |tree_t(doc)|, the unweighted leaf output.
int qr_tree_contributions(qr_ctx *c, const qr_flat_tree *trees, size_t ntrees, double *contribution) {
  QR_CHECK_CTX(c);
  if (ntrees == 0) return QR_OK;
  if (!trees || !contribution) { set_error("qr_tree_contributions: null argument"); return QR_EINVAL; }
  StagedTrees st{};
  QR_TRY(stage_trees(c, "qr_tree_contributions", trees, nullptr, ntrees, &st));
  const unsigned nblocks = (unsigned) ((c->N + kContribThreads - 1) / kContribThreads);
  double *d_part = nullptr, *d_out = nullptr;
  QR_TRY(dev_alloc(&d_part, (size_t) nblocks * ntrees));
  if (dev_alloc(&d_out, ntrees) != QR_OK) { cudaFree(d_part); return QR_ECUDA; }
  int rc = dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    QR_LAUNCH(c, PH_LEAF, tree_contrib_kernel<B>, nblocks, kContribThreads, 0, c->d_panels, c->N, st.nodes, st.roots,
              (uint32_t) ntrees, d_part);
    QR_LAUNCH(c, PH_LEAF, contrib_reduce_kernel, (unsigned) ntrees, 32, 0, d_part, nblocks, (uint32_t) ntrees, d_out);
    return QR_OK;
  });
  if (rc == QR_OK && c->comm) rc = comm_allreduce_sum_f64(c->comm, d_out, ntrees, c->stream);
  if (rc == QR_OK) {
    cudaError_t e = cudaMemcpyAsync(contribution, d_out, ntrees * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { set_error("qr_tree_contributions: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  }
  cudaFree(d_part);
  cudaFree(d_out);
  if (rc == QR_OK) for (size_t t = 0; t < ntrees; ++t) contribution[t] /= (double) c->N_global;
  return rc;
}

int qr_evaluate(qr_ctx *c, double *metric) {
  QR_CHECK_CTX(c);
  return evaluate(c, metric);
}

int qr_boost_iteration(qr_ctx *c, qr_flat_tree *tree, double *metric) {
  QR_CHECK_CTX(c);
  QR_CHECK_TRAINABLE(c);
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
int qr_last_tree_rounds(qr_ctx *c, uint32_t *rounds, double *beta) {
  QR_CHECK_CTX(c);
  if (rounds) *rounds = c->nrounds;
  if (beta) *beta = c->beta;
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
int qr_hist_kernel_time(qr_ctx *c, double *ms, uint64_t *launches, double *docs, int reset) {
  QR_CHECK_CTX(c);
  if (ms) *ms = c->histk_ms;
  if (launches) *launches = c->histk_launches;
  if (docs) *docs = c->histk_docs;
  if (reset) { c->histk_ms = 0; c->histk_launches = 0; c->histk_docs = 0; }
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

// Self-test of the ordered squares sum (qr_exact_kernels.cuh): the parallel scheme and the plain chain on the same
// values (as one node whose list is the identity), so that a test can require bit equality on adversarial inputs.
int qr_selftest_ordered_squares(const double *values, size_t n, int fused, int device, double *parallel, double *serial,
                                uint64_t *replayed_chunks) {
  if (!values || n == 0 || n > 0xfffffff0u || !parallel || !serial) { set_error("qr_selftest_ordered_squares: bad arguments"); return QR_EINVAL; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_error("no CUDA device"); return QR_ECUDA; }
  QR_CUDA(cudaSetDevice(device < 0 ? 0 : device));
  double *d_v = nullptr, *d_out = nullptr;
  NodeTask *d_t = nullptr;
  SqChunk *d_ch = nullptr;
  unsigned long long *d_rep = nullptr;
  const uint32_t nch = (uint32_t) ((n + kSqChunk - 1) / kSqChunk);
  QR_TRY(dev_alloc(&d_v, n));
  QR_TRY(dev_alloc(&d_out, 2));
  QR_TRY(dev_alloc(&d_t, 2));
  QR_TRY(dev_alloc(&d_ch, nch + 1));
  QR_TRY(dev_alloc(&d_rep, 1));
  NodeTask t[2];
  memset(t, 0, sizeof(t));
  for (int i = 0; i < 2; ++i) { t[i].n = (uint32_t) n; t[i].src = 2; t[i].whole = 1; t[i].build_left = 1; t[i].fused_sq = fused ? 1u : 0u; t[i].sq0 = (uint32_t) i; }
  int rc = QR_OK;
  auto ck = [&](cudaError_t e) { if (e != cudaSuccess && rc == QR_OK) { set_error("qr_selftest_ordered_squares: %s", cudaGetErrorString(e)); rc = QR_ECUDA; } };
  ck(cudaMemcpy(d_v, values, n * sizeof(double), cudaMemcpyHostToDevice));
  ck(cudaMemcpy(d_t, t, sizeof(t), cudaMemcpyHostToDevice));
  ck(cudaMemset(d_out, 0, 2 * sizeof(double)));
  ck(cudaMemset(d_rep, 0, sizeof(unsigned long long)));
  if (rc == QR_OK) {
    squares_exact_kernel<<<1, 32>>>(d_t + 1, nullptr, d_v, nullptr, nullptr, d_out, 1);
    if (n > kSqSerialMax) {
      ordered_squares_sums_kernel<<<(nch + 3) / 4, 128>>>(d_t, 1, nullptr, d_v, nullptr, nullptr, d_ch, nch);
      ordered_squares_binade_kernel<<<1, 256>>>(d_t, nullptr, d_ch);
      ordered_squares_pairs_kernel<<<(nch + 3) / 4, 128>>>(d_t, 1, nullptr, d_v, nullptr, nullptr, d_ch, nch);
      ordered_squares_resolve_kernel<<<1, 32>>>(d_t, nullptr, d_v, nullptr, nullptr, d_ch, d_out, d_rep);
    } else {
      squares_exact_kernel<<<1, 32>>>(d_t, nullptr, d_v, nullptr, nullptr, d_out, 1);
    }
    ck(cudaGetLastError());
    ck(cudaDeviceSynchronize());
  }
  double out[2] = {0, 0};
  unsigned long long rep = 0;
  ck(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
  ck(cudaMemcpy(&rep, d_rep, sizeof(rep), cudaMemcpyDeviceToHost));
  cudaFree(d_v); cudaFree(d_out); cudaFree(d_t); cudaFree(d_ch); cudaFree(d_rep);
  *parallel = out[0]; *serial = out[1];
  if (replayed_chunks) *replayed_chunks = rep;
  return rc;
}

}  // extern "C"
