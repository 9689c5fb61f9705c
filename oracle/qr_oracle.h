/* TEST INFRASTRUCTURE — CPU restatement of the reference's hot path (hpclab/quickrank).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this library, and only as the checker.  The product (quickrank_b200/, host/) never links,
 * imports or executes it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_reference.py)
 * against the unmodified reference sources compiled into oracle/_ref/libqr_ref.so, against the
 * reference's own known-answer unit tests (catch-unit-tests/metric/ir/test-{dcg,ndcg}.cc), and
 * against golden vectors generated from oracle/_ref and committed under tests/golden/.
 *
 * All file:line citations are into /root/reference (commit c569a59).
 */
#ifndef QR_ORACLE_H
#define QR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRO_NO_CUTOFF ((size_t) -1)   /* metric.h:46 */

/* Flat pre-order (left child first) tree; leaves have feature == -1.  Arrays are malloc'ed by
 * the oracle (capacity nnodes) and released with qro_tree_free. */
typedef struct {
  uint32_t nnodes, nleaves;
  int32_t *feature;        /* featureidx (rtnode.h:50), -1 for leaves */
  uint32_t *threshold_idx; /* best_thresholdid (rt.cc:289) */
  float *threshold;        /* thresholds[f][t] (rt.cc:317-318) */
  int32_t *left, *right;   /* node indices */
  double *value;           /* avglabel: leaf output after update_output, node mean otherwise */
  double *deviance;        /* rtnode.h:106 */
  uint64_t *count;         /* nsampleids */
} qro_tree;

typedef struct {
  size_t N, F;
  float **thr;         /* thresholds_[f] (mart.cc:136-170) */
  size_t *thr_size;    /* thresholds_size_[f] */
  uint32_t *bins;      /* stmap[f][doc] (rtnode_histogram.cc:227-253), feature-major [F][N] */
  const float *colmajor; /* borrowed: VerticalDataset::data_ (vertical_dataset.h:66) */
} qro_bins;

/* --- per-query sort: std::sort(idx, comp = score[i] > score[j]) with libstdc++ introsort
 *     (queryresults.cc:37-53; libstdc++ bits/stl_algo.h __introsort_loop/__final_insertion_sort) */
void qro_sort_desc(const double *scores, size_t n, uint32_t *idx);

/* --- metric (dcg.cc:33-57, ndcg.cc:35-58, metric.h:93-106) */
double qro_dcg_labels(const float *labels, size_t len, size_t cutoff);
double qro_idcg(const float *labels, size_t n, size_t cutoff);
double qro_dcg_query(const float *labels, const double *scores, size_t n, size_t cutoff);
double qro_ndcg_query(const float *labels, const double *scores, size_t n, size_t cutoff);
double qro_ndcg_dataset(const float *labels, const double *scores, const uint64_t *qoff, size_t Q,
                        size_t cutoff);
/* Ndcg::jacobian entry (i,j), i<j, ranks in sorted order (ndcg.cc:60-92) */
double qro_delta_ndcg(const float *sorted_labels, size_t n, size_t cutoff, double idcg, size_t i,
                      size_t j);

/* --- pseudo-responses */
void qro_lambdas(const double *scores, const float *labels, const uint64_t *qoff, size_t Q,
                 size_t cutoff, double *lambdas, double *weights);
/* the same with sample_presence (LambdaMartSelective / StochasticNegative): presence[doc] != 0 = sampled */
void qro_lambdas_masked(const double *scores, const float *labels, const uint64_t *qoff, size_t Q,
                        size_t cutoff, const uint8_t *presence, double *lambdas, double *weights); /* lambdamart.cc:62-152 */
void qro_mart_pseudo(const double *scores, const float *labels, size_t N,
                     double *pseudo); /* mart.cc:418-431 */

/* --- init: argsort, thresholds, bin map (mart.cc:117-176, radix.cc:35-73,
 *     rtnode_histogram.cc:227-253) */
void qro_radix_argsort(const float *v, size_t n, uint64_t *idx);
qro_bins *qro_binning(const float *colmajor, size_t N, size_t F, size_t nthresholds);
void qro_bins_free(qro_bins *b);

/* --- tree fit.  lambdas = pseudoresponses_, weights = instance_weights_ (NULL => MART mean
 *     leaves, rt.cc:165-184; else Newton leaves, rt.cc:186-207).  depth == 0: leaf-wise
 *     RegressionTree (rt.cc:49-163); depth > 0: ObliviousRT (ot.cc:32-175).
 *     leaf_of_doc (optional, [N]) receives the DFS leaf index of every document. */
qro_tree *qro_fit_tree(const qro_bins *b, const double *lambdas, const double *weights,
                       size_t nleaves, size_t minls, size_t depth, uint32_t *leaf_of_doc);
/* the same on sampleids[0 .. nsampleids) only, in that order (the document-sampling trainers) */
qro_tree *qro_fit_tree_sampled(const qro_bins *b, const double *lambdas, const double *weights, const uint64_t *sampleids,
                               size_t nsampleids, size_t nleaves, size_t minls, size_t depth, uint32_t *leaf_of_doc);
void qro_tree_free(qro_tree *t);

/* Tie audit: the reference's split score lsum^2/lcnt + rsum^2/rcnt (rt.cc:272-283) of `ncand`
 * candidate (feature, threshold index) pairs on the node made of documents ids[0..n) (ascending),
 * with the histogram built from those samples in list order (rtnode_histogram.cc:51-63).
 * out[k] = -1 when the candidate violates the minimum leaf support. */
void qro_split_scores(const qro_bins *b, const double *lambdas, const uint64_t *ids, size_t n,
                      size_t minls, const uint32_t *cand_f, const uint32_t *cand_t, size_t ncand,
                      double *out);

/* scores[i] += shrinkage * tree(doc_i), walking the raw float columns (mart.cc:459-468) */
void qro_update_scores(const qro_tree *t, const float *colmajor, size_t N, double shrinkage,
                       double *scores);

/* --- ensemble scoring: sum_t weight_t * leaf_t(doc), row-major docs
 *     (ltr_algorithm.cc:44-52, ensemble.cc:111-118, rtnode.h:134-152) */
void qro_score_dataset(const qro_tree *const *trees, const double *weights, size_t ntrees,
                       const float *rowmajor, size_t N, size_t F, double *scores);

/* --- full boosting loop (mart.cc:307-381 without validation): algo 0 MART, 1 LAMBDAMART,
 *     2 OBVMART, 3 OBVLAMBDAMART.  Returns trees via out_trees[ntrees] (caller frees each) and
 *     the per-iteration training metric. */
int qro_train(int algo, const float *colmajor, const float *labels, const uint64_t *qoff, size_t N,
              size_t F, size_t Q, size_t ntrees, double shrinkage, size_t nthresholds,
              size_t nleaves, size_t depth, size_t minls, size_t cutoff, qro_tree **out_trees,
              double *out_metric, double *out_scores);

#ifdef __cplusplus
}
#endif
#endif
