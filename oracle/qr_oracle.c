/* TEST INFRASTRUCTURE — CPU restatement of the reference's hot path.  See qr_oracle.h.
 *
 * Written from the behaviour of hpclab/quickrank @ c569a59 (citations are file:line into
 * /root/reference) as compiled by g++ 13.3 with the flags in oracle/Makefile.  Where that build
 * fuses a multiply-add (checked in the disassembly of oracle/_ref/obj) this file calls fma()
 * explicitly and is itself compiled with -ffp-contract=off, so that the restatement is
 * bit-identical to oracle/_ref and does not depend on compiler contraction choices:
 *   lambdamart.cc:135-140  p[j]=fma(rho,d,p[j]); p[k]=fma(-rho,d,p[k]); w[.]=fma(rho*(1-rho),d,w[.])
 *   mart.cc:466            scores[i]=fma(shrinkage,leaf,scores[i])
 *   rtnode_histogram.cc:65-69 (child ctor)  squares_sum=fma(l,l,squares_sum)
 *   rtnode_histogram.cc:199-203 (update)    squares_sum+=l*l            (NOT fused in that build)
 *   ensemble.cc:116        sum += leaf*weight                           (NOT fused in that build)
 */
#include "qr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * libstdc++ std::sort (bits/stl_algo.h: __sort, __introsort_loop, __unguarded_partition_pivot,
 * __move_median_to_first, __unguarded_partition, __final_insertion_sort, __partial_sort via
 * bits/stl_heap.h), restated over an array of uint32 "positions" with a pluggable strict
 * comparator.  The sequence of moves depends only on comparator outcomes, so it reproduces the
 * permutation std::sort yields for ties (SURVEY.md section 7.1 "Sort").
 * ------------------------------------------------------------------------------------------ */
typedef int (*qro_less)(uint32_t a, uint32_t b, const void *ctx);

static void adjust_heap(uint32_t *first, long hole, long len, uint32_t value, qro_less comp,
                        const void *ctx) {
  const long top = hole;
  long child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp(first[child], first[child - 1], ctx)) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  /* __push_heap */
  long parent = (hole - 1) / 2;
  while (hole > top && comp(first[parent], value, ctx)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void heap_sort_range(uint32_t *first, uint32_t *last, qro_less comp, const void *ctx) {
  /* __partial_sort(first, last, last): __heap_select == __make_heap, then __sort_heap */
  long len = last - first;
  if (len >= 2) {
    long parent = (len - 2) / 2;
    for (;;) {
      uint32_t v = first[parent];
      adjust_heap(first, parent, len, v, comp, ctx);
      if (parent == 0) break;
      parent--;
    }
  }
  while (last - first > 1) {
    --last;
    uint32_t v = *last;
    *last = *first;
    adjust_heap(first, 0, last - first, v, comp, ctx);
  }
}

static void move_median_to_first(uint32_t *result, uint32_t *a, uint32_t *b, uint32_t *c,
                                 qro_less comp, const void *ctx) {
  uint32_t *pick;
  if (comp(*a, *b, ctx)) {
    if (comp(*b, *c, ctx)) pick = b;
    else if (comp(*a, *c, ctx)) pick = c;
    else pick = a;
  } else if (comp(*a, *c, ctx)) pick = a;
  else if (comp(*b, *c, ctx)) pick = c;
  else pick = b;
  uint32_t t = *result;
  *result = *pick;
  *pick = t;
}

static uint32_t *unguarded_partition(uint32_t *first, uint32_t *last, uint32_t *pivot,
                                     qro_less comp, const void *ctx) {
  for (;;) {
    while (comp(*first, *pivot, ctx)) ++first;
    --last;
    while (comp(*pivot, *last, ctx)) --last;
    if (!(first < last)) return first;
    uint32_t t = *first;
    *first = *last;
    *last = t;
    ++first;
  }
}

static void introsort_loop(uint32_t *first, uint32_t *last, long depth_limit, qro_less comp,
                           const void *ctx) {
  while (last - first > 16) {
    if (depth_limit == 0) {
      heap_sort_range(first, last, comp, ctx);
      return;
    }
    --depth_limit;
    uint32_t *mid = first + (last - first) / 2;
    move_median_to_first(first, first + 1, mid, last - 1, comp, ctx);
    uint32_t *cut = unguarded_partition(first + 1, last, first, comp, ctx);
    introsort_loop(cut, last, depth_limit, comp, ctx);
    last = cut;
  }
}

static void unguarded_linear_insert(uint32_t *last, qro_less comp, const void *ctx) {
  uint32_t val = *last;
  uint32_t *next = last - 1;
  while (comp(val, *next, ctx)) {
    *last = *next;
    last = next;
    --next;
  }
  *last = val;
}

static void insertion_sort(uint32_t *first, uint32_t *last, qro_less comp, const void *ctx) {
  if (first == last) return;
  for (uint32_t *i = first + 1; i != last; ++i) {
    if (comp(*i, *first, ctx)) {
      uint32_t val = *i;
      memmove(first + 1, first, (size_t) (i - first) * sizeof(uint32_t));
      *first = val;
    } else {
      unguarded_linear_insert(i, comp, ctx);
    }
  }
}

static void std_sort(uint32_t *first, uint32_t *last, qro_less comp, const void *ctx) {
  if (first == last) return;
  long n = last - first, lg = 0;
  for (long t = n; t > 1; t >>= 1) ++lg; /* std::__lg */
  introsort_loop(first, last, 2 * lg, comp, ctx);
  if (last - first > 16) {
    insertion_sort(first, first + 16, comp, ctx);
    for (uint32_t *i = first + 16; i != last; ++i) unguarded_linear_insert(i, comp, ctx);
  } else {
    insertion_sort(first, last, comp, ctx);
  }
}

/* queryresults.cc:37-45: comp(i, j) = values[i] > values[j] */
static int score_greater(uint32_t a, uint32_t b, const void *ctx) {
  const double *s = (const double *) ctx;
  return s[a] > s[b];
}
/* ndcg.cc:40-41: std::greater<int>() applied to float labels (truncating conversion) */
static int label_int_greater(uint32_t a, uint32_t b, const void *ctx) {
  const float *l = (const float *) ctx;
  return (int) l[a] > (int) l[b];
}

void qro_sort_desc(const double *scores, size_t n, uint32_t *idx) {
  for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t) i; /* queryresults.cc:50-51 */
  std_sort(idx, idx + n, score_greater, scores);         /* queryresults.cc:52 */
}

/* ------------------------------------------------------------------------------------------ */
/* dcg.cc:33-39.  log2(i + 2.0f) resolves to the double overload (the argument is a float that
 * holds an exact small integer), pow(2.0, label) is glibc pow. */
static inline size_t norm_cutoff(size_t k) { return k == 0 ? QRO_NO_CUTOFF : k; } /* metric.h:65-67 */

double qro_dcg_labels(const float *labels, size_t len, size_t cutoff) {
  cutoff = norm_cutoff(cutoff);
  const size_t size = cutoff < len ? cutoff : len;
  double dcg = 0.0;
  for (size_t i = 0; i < size; ++i)
    dcg += (pow(2.0, (double) labels[i]) - 1.0) / log2((double) ((float) i + 2.0f));
  return dcg;
}

/* ndcg.cc:35-47 */
double qro_idcg(const float *labels, size_t n, size_t cutoff) {
  uint32_t *idx = (uint32_t *) malloc((n ? n : 1) * sizeof(uint32_t));
  float *sorted = (float *) malloc((n ? n : 1) * sizeof(float));
  for (size_t i = 0; i < n; ++i) idx[i] = (uint32_t) i;
  std_sort(idx, idx + n, label_int_greater, labels);
  for (size_t i = 0; i < n; ++i) sorted[i] = labels[idx[i]];
  double r = qro_dcg_labels(sorted, n, cutoff);
  free(idx);
  free(sorted);
  return r;
}

/* dcg.cc:41-57 (+ queryresults.cc:55-62) */
double qro_dcg_query(const float *labels, const double *scores, size_t n, size_t cutoff) {
  cutoff = norm_cutoff(cutoff);
  const size_t size = cutoff < n ? cutoff : n;
  if (size == 0) return 0.0;
  uint32_t *idx = (uint32_t *) malloc(n * sizeof(uint32_t));
  float *sorted = (float *) malloc(size * sizeof(float));
  qro_sort_desc(scores, n, idx);
  for (size_t i = 0; i < size; ++i) sorted[i] = labels[idx[i]];
  double r = qro_dcg_labels(sorted, size, cutoff);
  free(idx);
  free(sorted);
  return r;
}

/* ndcg.cc:49-58 */
double qro_ndcg_query(const float *labels, const double *scores, size_t n, size_t cutoff) {
  if (n == 0) return 0.0;
  const double idcg = qro_idcg(labels, n, cutoff);
  if (idcg > 0) return qro_dcg_query(labels, scores, n, cutoff) / idcg;
  return 0.0;
}

/* metric.h:93-106: sequential sum over queries, then one division */
double qro_ndcg_dataset(const float *labels, const double *scores, const uint64_t *qoff, size_t Q,
                        size_t cutoff) {
  if (Q == 0) return 0.0;
  double avg = 0.0;
  for (size_t q = 0; q < Q; ++q) {
    size_t o = qoff[q], n = qoff[q + 1] - qoff[q];
    avg += qro_ndcg_query(labels + o, scores + o, n, cutoff);
  }
  return avg / (double) Q;
}

/* ndcg.cc:72-88 */
double qro_delta_ndcg(const float *sl, size_t n, size_t cutoff, double idcg, size_t i, size_t j) {
  cutoff = norm_cutoff(cutoff);
  const size_t size = cutoff < n ? cutoff : n;
  if (idcg <= 0.0 || i >= size || sl[i] == sl[j]) return 0.0;
  const double gain = pow(2.0, (double) sl[i]) - pow(2.0, (double) sl[j]);
  if (j < size)
    return (1.0 / log2((double) (j + 2)) - 1.0 / log2((double) (i + 2))) * gain / idcg;
  return (-1.0 / log2((double) (i + 2))) * gain / idcg;
}

/* lambdamart.cc:62-152, sample_presence == NULL branch */
void qro_lambdas(const double *scores, const float *labels, const uint64_t *qoff, size_t Q,
                 size_t cutoff, double *lambdas, double *weights) {
  cutoff = norm_cutoff(cutoff);
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t q = 0; q < Q; ++q) {
    const size_t off = qoff[q], n = qoff[q + 1] - qoff[q];
    for (size_t j = off; j < off + n; ++j) lambdas[j] = weights[j] = 0.0; /* :77-78 */
    if (n == 0) continue;
    uint32_t *unmap = (uint32_t *) malloc(n * sizeof(uint32_t));
    float *sl = (float *) malloc(n * sizeof(float));
    qro_sort_desc(scores + off, n, unmap);              /* rankedresults.cc:32 */
    for (size_t i = 0; i < n; ++i) sl[i] = labels[off + unmap[i]]; /* rankedresults.cc:36-39 */
    const double idcg = qro_idcg(sl, n, cutoff);        /* ndcg.cc:68 */
    for (size_t j = 0; j < n; ++j) {                    /* :114 */
      const float jl = sl[j];
      const size_t ja = off + unmap[j];
      for (size_t k = 0; k < n; ++k) {                  /* :119 */
        if (k == j) continue;
        if (j >= cutoff && k >= cutoff) break;          /* :125-126 */
        const float kl = sl[k];
        if (jl > kl) {                                  /* :129 */
          const size_t ka = off + unmap[k];
          /* jacobian->at(j,k): symmetric, stored for i<j only (symmatrix.h:60-67) */
          const double d = fabs(j < k ? qro_delta_ndcg(sl, n, cutoff, idcg, j, k)
                                      : qro_delta_ndcg(sl, n, cutoff, idcg, k, j));
          const double rho = 1.0 / (1.0 + exp(scores[ja] - scores[ka])); /* :132-134 */
          const double t = (1.0 - rho) * rho;
          lambdas[ja] = fma(rho, d, lambdas[ja]);       /* :137 */
          lambdas[ka] = fma(-rho, d, lambdas[ka]);      /* :138 */
          weights[ja] = fma(t, d, weights[ja]);         /* :139 */
          weights[ka] = fma(t, d, weights[ka]);         /* :140 */
        }
      }
    }
    free(unmap);
    free(sl);
  }
}

/* lambdamart.cc:62-152, sample_presence != NULL branch (the document-sampling trainers: lambdamartselective.cc:194,
 * stochasticnegative.cc:197).  A query is compacted to its present documents (:84-98); the compacted list is RANKED by
 * scores_on_training_[d] with d the document's position inside its query — the query offset is not added at :94 —
 * while rho (:132-134) uses the documents' own scores.  Absent documents keep zero lambdas and weights. */
void qro_lambdas_masked(const double *scores, const float *labels, const uint64_t *qoff, size_t Q, size_t cutoff,
                        const uint8_t *presence, double *lambdas, double *weights) {
  cutoff = norm_cutoff(cutoff);
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t q = 0; q < Q; ++q) {
    const size_t off = qoff[q], n = qoff[q + 1] - qoff[q];
    for (size_t j = off; j < off + n; ++j) lambdas[j] = weights[j] = 0.0; /* :77-78 */
    if (n == 0) continue;
    size_t *map = (size_t *) malloc(n * sizeof(size_t));       /* map_from_cleaned */
    double *keys = (double *) malloc(n * sizeof(double));      /* training_scores_cleaned */
    float *lab = (float *) malloc(n * sizeof(float));          /* labels_cleaned */
    size_t m = 0;
    for (size_t d = 0; d < n; ++d)
      if (presence[off + d]) { map[m] = d; lab[m] = labels[off + d]; keys[m] = scores[d]; ++m; }   /* :91-96 */
    uint32_t *unmap = (uint32_t *) malloc((m ? m : 1) * sizeof(uint32_t));
    float *sl = (float *) malloc((m ? m : 1) * sizeof(float));
    qro_sort_desc(keys, m, unmap);
    for (size_t i = 0; i < m; ++i) sl[i] = lab[unmap[i]];
    const double idcg = qro_idcg(sl, m, cutoff);
    for (size_t j = 0; j < m; ++j) {
      const float jl = sl[j];
      const size_t ja = off + map[unmap[j]];                   /* :117 */
      for (size_t k = 0; k < m; ++k) {
        if (k == j) continue;
        if (j >= cutoff && k >= cutoff) break;
        const float kl = sl[k];
        if (jl > kl) {
          const size_t ka = off + map[unmap[k]];               /* :121 */
          const double d = fabs(j < k ? qro_delta_ndcg(sl, m, cutoff, idcg, j, k)
                                      : qro_delta_ndcg(sl, m, cutoff, idcg, k, j));
          const double rho = 1.0 / (1.0 + exp(scores[ja] - scores[ka]));
          const double t = (1.0 - rho) * rho;
          lambdas[ja] = fma(rho, d, lambdas[ja]);
          lambdas[ka] = fma(-rho, d, lambdas[ka]);
          weights[ja] = fma(t, d, weights[ja]);
          weights[ka] = fma(t, d, weights[ka]);
        }
      }
    }
    free(map); free(keys); free(lab); free(unmap); free(sl);
  }
}

/* mart.cc:418-431: float label minus double score */
void qro_mart_pseudo(const double *scores, const float *labels, size_t N, double *pseudo) {
  for (size_t i = 0; i < N; ++i) pseudo[i] = (double) labels[i] - scores[i];
}

/* ------------------------------------------------------------------------------------------ */
/* radix.cc:28-73: stable LSD argsort on sign-flipped float bits */
static inline uint32_t flip_bits(uint32_t x) { return x ^ ((uint32_t) (-(int32_t) (x >> 31)) | 0x80000000u); }

void qro_radix_argsort(const float *v, size_t n, uint64_t *idx) {
  uint32_t *key = (uint32_t *) malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *tmp = (uint32_t *) malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *ord = (uint32_t *) malloc((n ? n : 1) * sizeof(uint32_t));
  size_t *cnt = (size_t *) calloc(65537, sizeof(size_t));
  for (size_t i = 0; i < n; ++i) {
    uint32_t b;
    memcpy(&b, v + i, 4);
    key[i] = flip_bits(b);
  }
  for (size_t i = 0; i < n; ++i) cnt[(key[i] & 0xFFFF) + 1]++;
  for (size_t i = 0; i < 65536; ++i) cnt[i + 1] += cnt[i];
  for (size_t i = 0; i < n; ++i) tmp[cnt[key[i] & 0xFFFF]++] = (uint32_t) i;
  memset(cnt, 0, 65537 * sizeof(size_t));
  for (size_t i = 0; i < n; ++i) cnt[(key[i] >> 16) + 1]++;
  for (size_t i = 0; i < 65536; ++i) cnt[i + 1] += cnt[i];
  for (size_t i = 0; i < n; ++i) ord[cnt[key[tmp[i]] >> 16]++] = tmp[i];
  for (size_t i = 0; i < n; ++i) idx[i] = ord[i];
  free(key);
  free(tmp);
  free(ord);
  free(cnt);
}

/* mart.cc:117-176 (thresholds) + rtnode_histogram.cc:227-253 (stmap) */
qro_bins *qro_binning(const float *colmajor, size_t N, size_t F, size_t nthresholds) {
  qro_bins *b = (qro_bins *) calloc(1, sizeof(qro_bins));
  b->N = N;
  b->F = F;
  b->colmajor = colmajor;
  b->thr = (float **) calloc(F ? F : 1, sizeof(float *));
  b->thr_size = (size_t *) calloc(F ? F : 1, sizeof(size_t));
  b->bins = (uint32_t *) malloc((N && F ? N * F : 1) * sizeof(uint32_t));
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t f = 0; f < F; ++f) {
    const float *x = colmajor + f * N;
    uint64_t *idx = (uint64_t *) malloc((N ? N : 1) * sizeof(uint64_t));
    qro_radix_argsort(x, N, idx);
    /* distinct values, early stop once nthresholds+1 were seen (mart.cc:144-152) */
    size_t cap = (nthresholds == 0 ? N + 1 : nthresholds + 1);
    float *uniqs = (float *) malloc((cap + 1) * sizeof(float));
    size_t nu = 0;
    uniqs[nu++] = x[idx[0]];
    for (size_t j = 1; j < N && (nthresholds == 0 || nu != nthresholds + 1); ++j) {
      const float v = x[idx[j]];
      if (uniqs[nu - 1] < v) uniqs[nu++] = v;
    }
    if (nu <= nthresholds || nthresholds == 0) { /* mart.cc:155-158 */
      uniqs[nu++] = FLT_MAX;
      b->thr[f] = uniqs;
      b->thr_size[f] = nu;
    } else {                                     /* mart.cc:159-169: equal-width, float accumulate */
      free(uniqs);
      float *t = (float *) malloc((nthresholds + 1) * sizeof(float));
      float cur = x[idx[0]];
      const float step = (float) fabs((double) (x[idx[N - 1]] - cur)) / (float) nthresholds;
      for (size_t j = 0; j != nthresholds; cur += step) t[j++] = cur;
      t[nthresholds] = FLT_MAX;
      b->thr[f] = t;
      b->thr_size[f] = nthresholds + 1;
    }
    /* stmap: walk sorted docs against thresholds (rtnode_histogram.cc:241-251) */
    uint32_t *bin = b->bins + f * N;
    size_t last = (size_t) -1, j;
    for (size_t t = 0; t < b->thr_size[f]; ++t) {
      for (j = last + 1; j < N; ++j) {
        size_t k = idx[j];
        if (x[k] > b->thr[f][t]) break;
        bin[k] = (uint32_t) t;
      }
      last = j - 1;
    }
    free(idx);
  }
  return b;
}

void qro_bins_free(qro_bins *b) {
  if (!b) return;
  for (size_t f = 0; f < b->F; ++f) free(b->thr[f]);
  free(b->thr);
  free(b->thr_size);
  free(b->bins);
  free(b);
}

/* ------------------------------------------------------------------------------------------ */
/* Tree engine */
typedef struct {
  double **sum;     /* sumlbl[f][t], cumulative over t (rtnode_histogram.cc:59-62) */
  size_t **cnt;     /* count[f][t], cumulative */
  double squares;   /* squares_sum_ */
} hist_t;

typedef struct node_s {
  size_t *ids;
  size_t n;
  float threshold;
  uint32_t threshold_idx;
  double deviance, avglabel;
  struct node_s *left, *right;
  hist_t *hist;
  size_t feature; /* (size_t)-1 for leaves */
} node_t;

static hist_t *hist_alloc(const qro_bins *b) {
  hist_t *h = (hist_t *) malloc(sizeof(hist_t));
  h->sum = (double **) malloc(b->F * sizeof(double *));
  h->cnt = (size_t **) malloc(b->F * sizeof(size_t *));
  for (size_t f = 0; f < b->F; ++f) {
    h->sum[f] = (double *) calloc(b->thr_size[f], sizeof(double));
    h->cnt[f] = (size_t *) calloc(b->thr_size[f], sizeof(size_t));
  }
  h->squares = 0.0;
  return h;
}
static void hist_free(const qro_bins *b, hist_t *h) {
  if (!h) return;
  for (size_t f = 0; f < b->F; ++f) {
    free(h->sum[f]);
    free(h->cnt[f]);
  }
  free(h->sum);
  free(h->cnt);
  free(h);
}

/* RTNodeHistogram::update(labels, nsampleids, sampleids) (rtnode_histogram.cc:172-204) when
 * root != 0, RTNodeHistogram(parent, sampleids, n, labels) (rtnode_histogram.cc:41-70) otherwise.
 * The two differ only in how squares_sum_ is rounded in the oracle/_ref build (see file header). */
static hist_t *hist_from_samples(const qro_bins *b, const size_t *ids, size_t n, const double *lab,
                                 int root) {
  hist_t *h = hist_alloc(b);
#pragma omp parallel for schedule(static)
  for (size_t f = 0; f < b->F; ++f) {
    const uint32_t *bin = b->bins + f * b->N;
    double *s = h->sum[f];
    size_t *c = h->cnt[f];
    for (size_t i = 0; i < n; ++i) {
      const size_t d = ids[i];
      const uint32_t t = bin[d];
      s[t] += lab[d];
      c[t]++;
    }
    for (size_t t = 1; t < b->thr_size[f]; ++t) {
      s[t] += s[t - 1];
      c[t] += c[t - 1];
    }
  }
  double sq = 0.0;
  if (root)
    for (size_t i = 0; i < n; ++i) sq += lab[ids[i]] * lab[ids[i]];
  else
    for (size_t i = 0; i < n; ++i) sq = fma(lab[ids[i]], lab[ids[i]], sq);
  h->squares = sq;
  return h;
}

/* RTNodeHistogram(parent, left) (rtnode_histogram.cc:72-87) / transform_intorightchild (:206-217) */
static hist_t *hist_right(const qro_bins *b, const hist_t *parent, const hist_t *left) {
  hist_t *h = hist_alloc(b);
  for (size_t f = 0; f < b->F; ++f)
    for (size_t t = 0; t < b->thr_size[f]; ++t) {
      h->sum[f][t] = parent->sum[f][t] - left->sum[f][t];
      h->cnt[f][t] = parent->cnt[f][t] - left->cnt[f][t];
    }
  h->squares = parent->squares - left->squares;
  return h;
}

/* RTNode(sampleids, hist) (rtnode.h:97-107) */
static node_t *node_new(const qro_bins *b, size_t *ids, hist_t *h) {
  node_t *nd = (node_t *) calloc(1, sizeof(node_t));
  const size_t last = b->thr_size[0] - 1;
  nd->hist = h;
  nd->ids = ids;
  nd->n = h->cnt[0][last];
  const double sumlabel = h->sum[0][last];
  nd->avglabel = nd->n ? sumlabel / (double) nd->n : 0.0;
  nd->deviance = h->squares - sumlabel * sumlabel / (double) nd->n;
  nd->feature = (size_t) -1;
  nd->threshold_idx = UINT32_MAX;
  return nd;
}

/* RegressionTree::split scan (rt.cc:257-312): strict '>' from -1 in ascending (f,t) order */
static int best_split(const qro_bins *b, const hist_t *h, size_t minls, size_t *bf, size_t *bt) {
  double best = -1.0;
  size_t best_f = (size_t) -1, best_t = (size_t) -1;
  for (size_t f = 0; f < b->F; ++f) {
    const double *sl = h->sum[f];
    const size_t *sc = h->cnt[f];
    const size_t ts = b->thr_size[f];
    const double s = sl[ts - 1];
    const size_t c = sc[ts - 1];
    for (size_t t = 0; t < ts; ++t) {
      const size_t lc = sc[t], rc = c - lc;
      if (lc >= minls && rc >= minls) {
        const double ls = sl[t], rs = s - ls;
        const double score = ls * ls / (double) lc + rs * rs / (double) rc;
        if (score > best) {
          best = score;
          best_f = f;
          best_t = t;
        }
      }
    }
  }
  if (best == -1.0) return 0;
  *bf = best_f;
  *bt = best_t;
  return 1;
}

/* partition + child histograms (rt.cc:314-358) */
static void split_at(const qro_bins *b, node_t *nd, const double *lab, size_t f, size_t t,
                     int build_hists) {
  const float thr = b->thr[f][t];
  const size_t last = b->thr_size[f] - 1;
  const size_t cnt = nd->hist->cnt[f][last], lc = nd->hist->cnt[f][t], rc = cnt - lc;
  size_t *ls = (size_t *) malloc((lc ? lc : 1) * sizeof(size_t)), ln = 0;
  size_t *rs = (size_t *) malloc((rc ? rc : 1) * sizeof(size_t)), rn = 0;
  const float *x = b->colmajor + f * b->N;
  for (size_t i = 0; i < nd->n; ++i) {
    const size_t d = nd->ids[i];
    if (x[d] <= thr) ls[ln++] = d;
    else rs[rn++] = d;
  }
  nd->feature = f;
  nd->threshold = thr;
  nd->threshold_idx = (uint32_t) t;
  if (build_hists) {
    hist_t *lh = hist_from_samples(b, ls, ln, lab, 0);
    hist_t *rh = hist_right(b, nd->hist, lh);
    nd->left = node_new(b, ls, lh);
    nd->right = node_new(b, rs, rh);
  } else { /* ot.cc:142-149: last oblivious level, leaves carry sum/size */
    const double lsum = nd->hist->sum[f][t];
    const double rsum = nd->hist->sum[f][last] - lsum;
    node_t *l = (node_t *) calloc(1, sizeof(node_t)), *r = (node_t *) calloc(1, sizeof(node_t));
    l->ids = ls; l->n = ln; l->avglabel = lsum / (double) ln; l->feature = (size_t) -1;
    r->ids = rs; r->n = rn; r->avglabel = rsum / (double) rn; r->feature = (size_t) -1;
    l->threshold_idx = r->threshold_idx = UINT32_MAX;
    nd->left = l;
    nd->right = r;
  }
}

/* MaxHeap<RTNode*> (maxheap.h:31-106) */
typedef struct { double key; node_t *val; } heap_item;
typedef struct { heap_item *arr; size_t size, cap; } heap_t;
static void heap_init(heap_t *h, size_t init) {
  h->cap = init + 2;
  h->arr = (heap_item *) malloc(h->cap * sizeof(heap_item));
  h->size = 0;
  h->arr[0].key = DBL_MAX;
  h->arr[0].val = NULL;
}
static void heap_push(heap_t *h, double key, node_t *v) {
  if (++h->size == h->cap) {
    h->cap = 2 * h->cap + 1;
    h->arr = (heap_item *) realloc(h->arr, h->cap * sizeof(heap_item));
  }
  size_t p = h->size;
  while (key > h->arr[p >> 1].key) {
    h->arr[p] = h->arr[p >> 1];
    p >>= 1;
  }
  h->arr[p].key = key;
  h->arr[p].val = v;
}
static void heap_pop(heap_t *h) {
  const heap_item last = h->arr[h->size--];
  size_t child, p = 1;
  while (p << 1 <= h->size) {
    child = p << 1;
    if (child < h->size && h->arr[child + 1].key > h->arr[child].key) ++child;
    if (last.key < h->arr[child].key) h->arr[p] = h->arr[child];
    else break;
    p = child;
  }
  h->arr[p] = last;
}

static void free_nodes(const qro_bins *b, node_t *nd, node_t *root, hist_t *root_hist) {
  if (!nd) return;
  free_nodes(b, nd->left, root, root_hist);
  free_nodes(b, nd->right, root, root_hist);
  if (nd->hist && nd->hist != root_hist) hist_free(b, nd->hist);
  if (nd != root) free(nd->ids);
  free(nd);
}

static void count_nodes(const node_t *nd, uint32_t *nn, uint32_t *nl) {
  (*nn)++;
  if (nd->feature == (size_t) -1) { (*nl)++; return; }
  count_nodes(nd->left, nn, nl);
  count_nodes(nd->right, nn, nl);
}

static int32_t flatten(const node_t *nd, qro_tree *t, uint32_t *next) {
  const int32_t id = (int32_t) (*next)++;
  t->feature[id] = nd->feature == (size_t) -1 ? -1 : (int32_t) nd->feature;
  t->threshold_idx[id] = nd->threshold_idx;
  t->threshold[id] = nd->threshold;
  t->value[id] = nd->avglabel;
  t->deviance[id] = nd->deviance;
  t->count[id] = nd->n;
  t->left[id] = t->right[id] = -1;
  if (nd->feature != (size_t) -1) {
    t->left[id] = flatten(nd->left, t, next);
    t->right[id] = flatten(nd->right, t, next);
  }
  return id;
}

/* RTNode::save_leaves order (rtnode.cc:34-46) + update_output (rt.cc:165-207) */
static void fit_leaves(node_t *nd, const double *lab, const double *w, uint32_t *leaf_of_doc,
                       uint32_t *next_leaf) {
  if (nd->feature != (size_t) -1) {
    fit_leaves(nd->left, lab, w, leaf_of_doc, next_leaf);
    fit_leaves(nd->right, lab, w, leaf_of_doc, next_leaf);
    return;
  }
  const uint32_t li = (*next_leaf)++;
  if (w) {
    double s1 = 0.0, s2 = 0.0;
    for (size_t j = 0; j < nd->n; ++j) {
      s1 += lab[nd->ids[j]];
      s2 += w[nd->ids[j]];
    }
    nd->avglabel = s2 >= DBL_EPSILON ? s1 / s2 : 0.0; /* rt.cc:200 */
  } else {
    double ps = 0.0;
    for (size_t j = 0; j < nd->n; ++j) ps += lab[nd->ids[j]];
    nd->avglabel = ps / (double) nd->n;                /* rt.cc:178 */
  }
  if (leaf_of_doc)
    for (size_t j = 0; j < nd->n; ++j) leaf_of_doc[nd->ids[j]] = li;
}

static qro_tree *fit_tree_on(const qro_bins *b, const double *lab, const double *w, size_t *ids, size_t N,
                             size_t nleaves, size_t minls, size_t depth, uint32_t *leaf_of_doc);

qro_tree *qro_fit_tree(const qro_bins *b, const double *lab, const double *w, size_t nleaves,
                       size_t minls, size_t depth, uint32_t *leaf_of_doc) {
  const size_t N = b->N;
  size_t *ids = (size_t *) malloc((N ? N : 1) * sizeof(size_t));
  for (size_t i = 0; i < N; ++i) ids[i] = i;
  return fit_tree_on(b, lab, w, ids, N, nleaves, minls, depth, leaf_of_doc);
}

/* The document-sampling trainers (lambdamartselective.cc:199-206, stochasticnegative.cc:201-205): the root histogram
 * is refreshed over sampleids[0 .. nsampleids) in that order (RTNodeHistogram::update(labels, nsampleids, sampleids),
 * rtnode_histogram.cc:172-204) and the tree is grown and its leaf outputs fitted on those documents only.
 * leaf_of_doc is written for the sampled documents. */
qro_tree *qro_fit_tree_sampled(const qro_bins *b, const double *lab, const double *w, const uint64_t *sampleids,
                               size_t nsampleids, size_t nleaves, size_t minls, size_t depth, uint32_t *leaf_of_doc) {
  size_t *ids = (size_t *) malloc((nsampleids ? nsampleids : 1) * sizeof(size_t));
  for (size_t i = 0; i < nsampleids; ++i) ids[i] = (size_t) sampleids[i];
  return fit_tree_on(b, lab, w, ids, nsampleids, nleaves, minls, depth, leaf_of_doc);
}

static qro_tree *fit_tree_on(const qro_bins *b, const double *lab, const double *w, size_t *ids, size_t N,
                             size_t nleaves, size_t minls, size_t depth, uint32_t *leaf_of_doc) {
  hist_t *root_hist = hist_from_samples(b, ids, N, lab, 1); /* mart.cc:335 */
  node_t *root = node_new(b, ids, root_hist);

  if (depth == 0) {
    /* RegressionTree::fit (rt.cc:49-84) */
    heap_t heap;
    heap_init(&heap, nleaves);
    size_t taken = 0, bf, bt;
    if (root->deviance > 0.0 && best_split(b, root->hist, minls, &bf, &bt)) {
      split_at(b, root, lab, bf, bt, 1);
      heap_push(&heap, root->left->deviance, root->left);
      heap_push(&heap, root->right->deviance, root->right);
    }
    while (heap.size != 0 && (nleaves == 0 || taken + heap.size < nleaves)) {
      node_t *nd = heap.arr[1].val;
      heap_pop(&heap);
      if (nd->deviance > 0.0 && best_split(b, nd->hist, minls, &bf, &bt)) {
        split_at(b, nd, lab, bf, bt, 1);
        heap_push(&heap, nd->left->deviance, nd->left);
        heap_push(&heap, nd->right->deviance, nd->right);
      } else {
        ++taken;
      }
      if (nd->hist != root_hist) hist_free(b, nd->hist);
      nd->hist = NULL;
    }
    free(heap.arr);
  } else {
    /* ObliviousRT::fit (ot.cc:32-175) */
    const size_t nn = (size_t) 1 << (depth + 1);
    node_t **arr = (node_t **) calloc(nn, sizeof(node_t *));
    arr[0] = root;
    double **ss = (double **) malloc(b->F * sizeof(double *));
    for (size_t f = 0; f < b->F; ++f) ss[f] = (double *) malloc(b->thr_size[f] * sizeof(double));
    const double invalid = -DBL_MAX;
    for (size_t d = 0; d < depth; ++d) {
      const size_t lbegin = ((size_t) 1 << d) - 1, lend = ((size_t) 1 << (d + 1)) - 1;
      for (size_t f = 0; f < b->F; ++f)
        for (size_t t = 0; t < b->thr_size[f]; ++t) ss[f][t] = 0.0;
      for (size_t i = lbegin; i < lend; ++i) { /* fill (ot.cc:177-201) */
        const hist_t *h = arr[i]->hist;
        for (size_t f = 0; f < b->F; ++f) {
          const size_t ts = b->thr_size[f];
          const double s = h->sum[f][ts - 1];
          const size_t c = h->cnt[f][ts - 1];
          for (size_t t = 0; t < ts; ++t)
            if (ss[f][t] != invalid) {
              const size_t lc = h->cnt[f][t], rc = c - lc;
              if (lc >= minls && rc >= minls) {
                const double ls = h->sum[f][t], rs = s - ls;
                ss[f][t] += ls * ls / (double) lc + rs * rs / (double) rc;
              } else {
                ss[f][t] = invalid;
              }
            }
        }
      }
      double best = 0.0; /* ot.cc:72-92: strict '>' from 0.0 */
      size_t bf = (size_t) -1, bt = (size_t) -1;
      for (size_t f = 0; f < b->F; ++f)
        for (size_t t = 0; t < b->thr_size[f]; ++t)
          if (ss[f][t] != invalid && ss[f][t] > best) {
            best = ss[f][t];
            bf = f;
            bt = t;
          }
      if (best == invalid || best == 0.0) break; /* ot.cc:96 */
      for (size_t i = lbegin; i < lend; ++i) {
        node_t *nd = arr[i];
        split_at(b, nd, lab, bf, bt, d != depth - 1);
        arr[2 * i + 1] = nd->left;
        arr[2 * i + 2] = nd->right;
        if (d) { /* ot.cc:157-160 */
          if (nd->hist != root_hist) hist_free(b, nd->hist);
          nd->hist = NULL;
        }
      }
    }
    for (size_t f = 0; f < b->F; ++f) free(ss[f]);
    free(ss);
    free(arr);
  }

  uint32_t next_leaf = 0;
  fit_leaves(root, lab, w, leaf_of_doc, &next_leaf);

  qro_tree *t = (qro_tree *) calloc(1, sizeof(qro_tree));
  count_nodes(root, &t->nnodes, &t->nleaves);
  t->feature = (int32_t *) malloc(t->nnodes * sizeof(int32_t));
  t->threshold_idx = (uint32_t *) malloc(t->nnodes * sizeof(uint32_t));
  t->threshold = (float *) malloc(t->nnodes * sizeof(float));
  t->left = (int32_t *) malloc(t->nnodes * sizeof(int32_t));
  t->right = (int32_t *) malloc(t->nnodes * sizeof(int32_t));
  t->value = (double *) malloc(t->nnodes * sizeof(double));
  t->deviance = (double *) malloc(t->nnodes * sizeof(double));
  t->count = (uint64_t *) malloc(t->nnodes * sizeof(uint64_t));
  uint32_t next = 0;
  flatten(root, t, &next);

  free_nodes(b, root, root, root_hist);
  hist_free(b, root_hist);
  free(ids);
  return t;
}

void qro_split_scores(const qro_bins *b, const double *lab, const uint64_t *ids64, size_t n,
                      size_t minls, const uint32_t *cand_f, const uint32_t *cand_t, size_t ncand,
                      double *out) {
  for (size_t k = 0; k < ncand; ++k) {
    const size_t f = cand_f[k], t = cand_t[k], ts = b->thr_size[f];
    const uint32_t *bin = b->bins + f * b->N;
    double *s = (double *) calloc(ts, sizeof(double));
    size_t *c = (size_t *) calloc(ts, sizeof(size_t));
    for (size_t i = 0; i < n; ++i) {
      const size_t d = (size_t) ids64[i];
      s[bin[d]] += lab[d];
      c[bin[d]]++;
    }
    for (size_t j = 1; j < ts; ++j) {
      s[j] += s[j - 1];
      c[j] += c[j - 1];
    }
    const size_t lc = c[t], rc = c[ts - 1] - lc;
    out[k] = -1.0;
    if (t < ts && lc >= minls && rc >= minls) {
      const double ls = s[t], rs = s[ts - 1] - ls;
      out[k] = ls * ls / (double) lc + rs * rs / (double) rc;
    }
    free(s);
    free(c);
  }
}

void qro_tree_free(qro_tree *t) {
  if (!t) return;
  free(t->feature);
  free(t->threshold_idx);
  free(t->threshold);
  free(t->left);
  free(t->right);
  free(t->value);
  free(t->deviance);
  free(t->count);
  free(t);
}

/* rtnode.h:134-152 on a flat tree; stride = distance between features of one document */
static inline double walk(const qro_tree *t, const float *d, size_t stride) {
  int32_t n = 0;
  while (t->feature[n] >= 0)
    n = d[(size_t) t->feature[n] * stride] <= t->threshold[n] ? t->left[n] : t->right[n];
  return t->value[n];
}

/* mart.cc:459-468 (fused multiply-add in the oracle/_ref build) */
void qro_update_scores(const qro_tree *t, const float *colmajor, size_t N, double shrinkage,
                       double *scores) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) scores[i] = fma(shrinkage, walk(t, colmajor + i, N), scores[i]);
}

/* ltr_algorithm.cc:44-52 -> ensemble.cc:111-118 (product and sum rounded separately) */
void qro_score_dataset(const qro_tree *const *trees, const double *weights, size_t ntrees,
                       const float *rowmajor, size_t N, size_t F, double *scores) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < N; ++i) {
    const float *d = rowmajor + i * F;
    double sum = 0.0;
    for (size_t k = 0; k < ntrees; ++k) {
      const double p = walk(trees[k], d, 1) * weights[k];
      sum += p;
    }
    scores[i] = sum;
  }
}

/* mart.cc:307-381 without validation/subsampling */
int qro_train(int algo, const float *colmajor, const float *labels, const uint64_t *qoff, size_t N,
              size_t F, size_t Q, size_t ntrees, double shrinkage, size_t nthresholds,
              size_t nleaves, size_t depth, size_t minls, size_t cutoff, qro_tree **out_trees,
              double *out_metric, double *out_scores) {
  const int lambda = (algo == 1 || algo == 3);
  const int obv = (algo == 2 || algo == 3);
  cutoff = norm_cutoff(cutoff);
  qro_bins *b = qro_binning(colmajor, N, F, nthresholds);
  double *scores = (double *) calloc(N ? N : 1, sizeof(double));
  double *lam = (double *) calloc(N ? N : 1, sizeof(double));
  double *w = lambda ? (double *) calloc(N ? N : 1, sizeof(double)) : NULL;
  for (size_t m = 0; m < ntrees; ++m) {
    if (lambda) qro_lambdas(scores, labels, qoff, Q, cutoff, lam, w);
    else qro_mart_pseudo(scores, labels, N, lam);
    qro_tree *t = qro_fit_tree(b, lam, w, obv ? ((size_t) 1 << depth) : nleaves, minls,
                               obv ? depth : 0, NULL);
    qro_update_scores(t, colmajor, N, shrinkage, scores);
    if (out_metric) out_metric[m] = qro_ndcg_dataset(labels, scores, qoff, Q, cutoff);
    if (out_trees) out_trees[m] = t;
    else qro_tree_free(t);
  }
  if (out_scores) memcpy(out_scores, scores, N * sizeof(double));
  free(scores);
  free(lam);
  free(w);
  qro_bins_free(b);
  return 0;
}
