/* quickrank_b200 — C ABI of the B200-native LambdaMART / GBRT hot path.
 *
 * This is the drop-in boundary: everything the reference (hpclab/quickrank @ c569a59) does
 * inside its per-iteration loop and its ensemble-scoring loop is reachable through these entry
 * points, with plain pointers and sizes only.  The reference has no FFI of its own; its extension
 * seam is the set of protected virtual hooks of `quickrank::learning::forests::Mart`
 * (include/learning/forests/mart.h:118-147) plus `LTR_Algorithm::score_dataset`
 * (include/learning/ltr_algorithm.h:73) and the link symbol `double ranker(float*)`
 * (src/quickscore.cc:62).  Each function below names the reference interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the subclass a QuickRank
 * maintainer would add on top of this header.
 *
 * Conventions (same as the reference's hooks): single caller thread per
 * context, not re-entrant; all pointers are HOST pointers unless the name says `_device`; every
 * function returns 0 on success and a QR_E* code otherwise, qr_last_error() then holds a message
 * (the reference prints to std::cerr and exits — the C++ host in host/ does exactly that with the
 * message).  There is no CPU fallback: without a CUDA device every compute entry fails with
 * QR_ENODEVICE.
 */
#ifndef QUICKRANK_B200_H
#define QUICKRANK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QR_OK 0
#define QR_EINVAL 1     /* bad argument */
#define QR_ENODEVICE 2  /* no CUDA device / CUDA runtime failure at start-up */
#define QR_ECUDA 3      /* CUDA error during execution */
#define QR_ELIMIT 4     /* input exceeds a documented limit of this build */
#define QR_ENOMEM 5
#define QR_ECOMM 6      /* NCCL / multi-GPU failure */

/* --algo names of the reference (src/learning/ltr_algorithm_factory.cc:67-219) */
enum qr_algo {
  QR_ALGO_MART = 0,           /* forests::Mart */
  QR_ALGO_LAMBDAMART = 1,     /* forests::LambdaMart */
  QR_ALGO_OBVMART = 2,        /* forests::ObliviousMart */
  QR_ALGO_OBVLAMBDAMART = 3   /* forests::ObliviousLambdaMart */
};

/* How per-bin gradient sums are accumulated.
 *  FAST       64-bit fixed-point integer accumulation in shared memory: order-independent, so the
 *             histogram (and every split) is deterministic, identical for duplicated columns and
 *             identical for any number of GPUs.
 *  REFERENCE  every per-bin sum, prefix sum, squares sum and leaf sum is accumulated in FP64 in the
 *             reference's own order (ascending document id per bin, then ascending bin:
 *             src/learning/tree/rtnode_histogram.cc:51-63), which makes split indices bit-exact
 *             with the reference even where its own rounding breaks a tie.  Parity mode. */
enum qr_hist_mode { QR_HIST_FAST = 0, QR_HIST_REFERENCE = 1 };

typedef struct {
  uint32_t algo;             /* enum qr_algo */
  uint32_t nleaves;          /* --num-leaves (leaf-wise algos) */
  uint32_t treedepth;        /* --tree-depth (oblivious algos; leaves = 1 << depth) */
  uint32_t minleafsupport;   /* --min-leaf-support */
  uint64_t nthresholds;      /* --num-thresholds, 0 = one per distinct value (mart.cc:144-158) */
  uint64_t ndcg_cutoff;      /* NDCG@k; 0 = no cutoff (metric.h:65-67) */
  double shrinkage;          /* --shrinkage */
  uint32_t hist_mode;        /* enum qr_hist_mode */
  int32_t device;            /* CUDA device ordinal, -1 = current device */
} qr_params;

/* Flat pre-order (left child first) regression tree; leaves have feature == -1.  Replaces the
 * RTNode pointer graph (include/learning/tree/rtnode.h:38-168).  Arrays are caller-owned with
 * room for `capacity` nodes (2*leaves-1 suffices). */
typedef struct {
  uint32_t capacity;
  uint32_t nnodes;
  uint32_t nleaves;
  int32_t *feature;         /* RTNode::featureidx (0-based); XML featureid = feature+1 (rt.cc:350-352) */
  uint32_t *threshold_idx;  /* index into the feature's threshold list (rt.cc:289) */
  float *threshold;         /* RTNode::threshold (rt.cc:317-318) */
  int32_t *left;            /* node index of the `<=` child */
  int32_t *right;
  double *value;            /* RTNode::avglabel: leaf output for leaves, node mean otherwise */
  double *deviance;         /* RTNode::deviance (rtnode.h:106) */
  uint64_t *count;          /* RTNode::nsampleids */
} qr_flat_tree;

typedef struct qr_ctx qr_ctx;        /* one training run on one GPU (Mart::init .. Mart::clear) */
typedef struct qr_scorer qr_scorer;  /* one ensemble resident on one GPU */

const char *qr_last_error(void);
/* Number of visible CUDA devices (0 if none); never fails. */
int qr_device_count(void);

/* ---- training ------------------------------------------------------------------------------ */

/* Replaces Mart::init / LambdaMart::init (mart.cc:117-176, lambdamart.cc:35-39) together with
 * RTRootHistogram::RTRootHistogram (rtnode_histogram.cc:227-253): per-feature argsort, threshold
 * lists, bin map; uploads labels/query offsets; zeroes scores.  `feat_colmajor` is
 * VerticalDataset::data_ (include/data/vertical_dataset.h:66): feat[f*N + doc]. */
int qr_ctx_create(const float *feat_colmajor, size_t N, size_t F, const float *labels,
                  const uint64_t *qoffsets, size_t Q, const qr_params *params, qr_ctx **out);
/* Same, from the row-major layout of data::Dataset (include/data/dataset.h:65-66); the transpose
 * done by VerticalDataset's constructor (vertical_dataset.cc:29-70) happens on the device. */
int qr_ctx_create_rowmajor(const float *feat_rowmajor, size_t N, size_t F, const float *labels,
                           const uint64_t *qoffsets, size_t Q, const qr_params *params,
                           qr_ctx **out);
/* Evaluation-only context for a second dataset (validation / test set): its documents are binned
 * with the THRESHOLDS OF `train`, so that qr_apply_tree walks the same splits the float test
 * `x <= threshold` would (SURVEY.md section 7.1 "Bins").  Replaces scores_on_validation_ together with
 * Mart::update_modelscores(Dataset...) (mart.cc:447-457) and Metric::evaluate_dataset(Dataset...)
 * (metric.h:77-91) of the validation branch of Mart::learn (mart.cc:354-359): use qr_apply_tree,
 * qr_evaluate, qr_get_scores / qr_set_scores on it.  Row-major features. */
int qr_ctx_create_eval(qr_ctx *train, const float *feat_rowmajor, size_t N, size_t F, const float *labels,
                       const uint64_t *qoffsets, size_t Q, qr_ctx **out);
/* A document SAMPLE of `full` as a training context of its own: what LambdaMartSelective::learn
 * (lambdamartselective.cc:185-206) and StochasticNegative::learn (stochasticnegative.cc:188-206) do
 * with `sampleids` / `sample_presence` — pseudo-responses over the sampled documents of each query
 * (LambdaMart::compute_pseudoresponses with sample_presence, lambdamart.cc:84-105), root histogram over
 * the sample (RTNodeHistogram::update(labels, nsampleids, sampleids), rtnode_histogram.cc:172-204), tree
 * fit and leaf outputs on it.  The rows are the sampled documents in ascending order within each query
 * (the "cleaned" order of lambdamart.cc:90-98; queries left empty are dropped), binned with the
 * thresholds of `full`; sums are fixed-point (QR_HIST_FAST) whatever the mode of `full`.
 * feat_rowmajor = the sampled rows [N][F], or NULL: the bins are then gathered on the device from those `full`
 * already holds (no feature values are uploaded and nothing is binned again) — what the host trainers do.
 * src_doc[i] = the document of `full` row i is; key_doc[i] = the document of `full` whose score row i is
 * RANKED by (lambdamart.cc:94 reads scores_on_training_[d] with d relative to the query, so
 * key_doc[i] = src_doc[i] - offset(query of i); NULL: rank by the document's own score).
 * Per iteration: qr_sample_pull_scores, qr_compute_pseudoresponses and qr_fit_tree on the sample, then
 * qr_apply_tree(full, tree, shrinkage) and qr_evaluate(full) (update_modelscores and the metric run
 * over ALL documents, lambdamartselective.cc:211-215). */
int qr_ctx_create_sample(qr_ctx *full, const float *feat_rowmajor, size_t N, size_t F, const float *labels,
                         const uint64_t *qoffsets, size_t Q, const uint32_t *src_doc, const uint32_t *key_doc,
                         qr_ctx **out);
/* A new draw into an existing sample context (created with feat_rowmajor == NULL: it is sized for every document of
 * `full`): same arguments as above, nothing is allocated or freed — the bins of the new documents are gathered, the
 * per-query tables and per-bin counts rebuilt, the state of the previous sample (scores, last tree) dropped.  What
 * LambdaMartSelective::learn does every `sampling_iterations` trees (lambdamartselective.cc:170-192). */
int qr_sample_redraw(qr_ctx *sample, qr_ctx *full, size_t N, const float *labels, const uint64_t *qoffsets, size_t Q,
                     const uint32_t *src_doc, const uint32_t *key_doc);
/* Copies the current scores of `full` into the sample (own scores and ranking keys). */
int qr_sample_pull_scores(qr_ctx *sample, qr_ctx *full);
/* Replaces Mart::clear (mart.cc:178-206). */
int qr_ctx_destroy(qr_ctx *ctx);

/* thresholds_[f] / thresholds_size_[f] (mart.h:150-151); pointer valid until qr_ctx_destroy. */
int qr_get_thresholds(qr_ctx *ctx, size_t f, const float **thresholds, size_t *n);

/* Replaces LambdaMart::compute_pseudoresponses (lambdamart.cc:62-152) or, for the MART algos,
 * Mart::compute_pseudoresponses (mart.cc:418-431).  Result stays on the device. */
int qr_compute_pseudoresponses(qr_ctx *ctx);

/* Replaces hist_->update (mart.cc:335, rtnode_histogram.cc:172-204) + fit_regressor_on_gradient
 * (mart.cc:433-445 / lambdamart.cc:47-60 / obliviousmart.cc, obliviouslambdamart.cc:55-66):
 * RegressionTree::fit or ObliviousRT::fit followed by update_output.  `out` may be NULL. */
int qr_fit_tree(qr_ctx *ctx, qr_flat_tree *out);

/* Replaces Mart::update_modelscores(VerticalDataset...) (mart.cc:459-468) for the tree just
 * fitted: scores[i] += weight * leaf(doc_i). */
int qr_update_modelscores(qr_ctx *ctx, double weight);

/* scores[i] += weight * tree(doc_i) for an arbitrary tree of this context's binning (DART's
 * add/subtract passes, dart.cc:634-687; weight carries the sign). */
int qr_apply_tree(qr_ctx *ctx, const qr_flat_tree *tree, double weight);
/* The same for a set of trees in ONE pass over the documents: for every document the trees are
 * applied in array order, scores[i] = fma(weights[t], tree_t(doc_i), scores[i]) — the per-document
 * sequence of operations of Dart::update_modelscores' loop over `trees_to_update` (dart.cc:634-650),
 * which visits the documents once per tree instead. */
int qr_apply_trees(qr_ctx *ctx, const qr_flat_tree *trees, const double *weights, size_t ntrees);
/* Replaces Dart::update_contribution_scores (src/learning/forests/dart.cc:689-706), for `ntrees` trees in one pass:
 * contribution[t] = mean over the dataset (all ranks' documents) of |tree_t(doc)|, the unweighted leaf output
 * (RTNode::score_instance), the quantity DART's CONTR / WCONTR sampling and normalisation read
 * (dart.cc:774-850, 917-940). */
int qr_tree_contributions(qr_ctx *ctx, const qr_flat_tree *trees, size_t ntrees, double *contribution);

/* Replaces Metric::evaluate_dataset(VerticalDataset, scores) for Ndcg (metric.h:93-106,
 * ndcg.cc:49-58) on the training scores held by the context. */
int qr_evaluate(qr_ctx *ctx, double *metric);

/* One whole iteration of Mart::learn's loop body (mart.cc:331-347) without returning to the
 * host in between; tree and metric may be NULL. */
int qr_boost_iteration(qr_ctx *ctx, qr_flat_tree *tree, double *metric);

/* parity taps (device <-> host copies of the state arrays named in mart.h:153-160) */
int qr_get_scores(qr_ctx *ctx, double *scores);
int qr_set_scores(qr_ctx *ctx, const double *scores);
int qr_get_pseudoresponses(qr_ctx *ctx, double *lambdas, double *weights /* may be NULL */);
int qr_set_pseudoresponses(qr_ctx *ctx, const double *lambdas, const double *weights);
int qr_get_leaf_assignment(qr_ctx *ctx, uint32_t *leaf_of_doc); /* DFS leaf index per doc */
int qr_get_bins(qr_ctx *ctx, size_t f, uint32_t *bins);         /* stmap[f][doc] */
int qr_get_ranking(qr_ctx *ctx, uint32_t *position_of_rank);    /* per query, as RankedResults::pos_of_rank */
/* measured per-tree quantities for the roofline formula:
 * rho = sum over splits of n_left / N, sigma = sum over splits of n_node / N */
int qr_last_tree_stats(qr_ctx *ctx, double *rho, double *sigma, uint32_t *nsplits);
/* growth rounds of the last tree (batched node expansions, see quickrank_b200/csrc/qr_tree_host.cuh)
 * and beta = documents whose histogram rows were actually accumulated / N (root included): with the
 * smaller-child rule of the fixed-point mode this is below 1 + rho */
int qr_last_tree_rounds(qr_ctx *ctx, uint32_t *rounds, double *beta);
/* number of kernel launches issued by this context so far */
uint64_t qr_launch_count(qr_ctx *ctx);
/* device time (ms, CUDA events on the context's stream) spent per phase since the last reset:
 * [0] pseudo-responses, [1] histogram kernels, [2] split scan, [3] partition, [4] leaf fit +
 * score update, [5] ranking/NDCG; also returns per-phase launch counts (either may be NULL). */
int qr_phase_times(qr_ctx *ctx, double ms[6], uint64_t launches[6], int reset);
int qr_set_profiling(qr_ctx *ctx, int enabled);
/* Self-test of QR_HIST_REFERENCE's squares sum (RTNodeHistogram's squares_sum_, rtnode_histogram.cc:65-69 fused,
 * :199-203 multiply then add): sum of values[i]^2 in index order, once by the parallel scheme the mode uses for long
 * lists (quickrank_b200/csrc/qr_exact_kernels.cuh) and once by the plain sequential chain; the two must be
 * bit-identical.  replayed_chunks (may be NULL): 256-addend chunks the parallel scheme had to replay one by one. */
int qr_selftest_ordered_squares(const double *values, size_t n, int fused, int device, double *parallel,
                                double *serial, uint64_t *replayed_chunks);
/* while profiling is enabled every launch of the histogram kernel (the dominant kernel of the path) is
 * bracketed by its own pair of CUDA events: accumulated device time (ms), number of launches and
 * number of documents whose bin rows those launches accumulated, since the last reset */
int qr_hist_kernel_time(qr_ctx *ctx, double *ms, uint64_t *launches, double *docs, int reset);
/* CUDA-event stopwatch on the context's stream (the stream every kernel of the context is
 * launched on): start records an event, stop records a second one, synchronises and returns the
 * elapsed device time in milliseconds. */
int qr_timer_start(qr_ctx *ctx);
int qr_timer_stop(qr_ctx *ctx, double *ms);

/* ---- multi-GPU (one process per GPU; documents sharded by query) ------ */
#define QR_COMM_ID_BYTES 128
/* rank 0 creates the id, the host broadcasts it by any means (torch.distributed, MPI, a file). */
int qr_comm_unique_id(unsigned char id[QR_COMM_ID_BYTES]);
/* Creates the training context of rank `rank` of `world`: this process holds a contiguous range of
 * whole queries (its documents only); thresholds are computed over the union of all ranks' feature
 * values, and afterwards histograms, squares, leaf sums and the metric are all-reduced (the per-round
 * histograms through peer memory over NVLink when possible, everything else over NCCL) so that every
 * rank grows the identical tree.  `rowmajor` selects the feature layout (0: column-major
 * as qr_ctx_create, 1: row-major as qr_ctx_create_rowmajor).  Collective: every rank must call it. */
int qr_ctx_create_sharded(const float *feat, int rowmajor, size_t N, size_t F, const float *labels,
                          const uint64_t *qoffsets, size_t Q, const qr_params *params,
                          const unsigned char id[QR_COMM_ID_BYTES], int rank, int world, qr_ctx **out);

/* How the per-round histogram exchange of this context travels: 0 = single GPU (no exchange),
 * 1 = NCCL all-reduces, 2 = one peer-memory kernel per round (every rank maps every other rank's
 * histogram pool over CUDA IPC and reduces in place through NVLink; chosen automatically when all
 * ranks can map each other, QR_PEER_REDUCE=0 forces NCCL).  Results are identical either way:
 * the exchanged sums are integers. */
int qr_ctx_comm_transport(const qr_ctx *ctx);

/* ---- scoring ------------------------------------------------------------------------------- */

/* Uploads an ensemble (replaces building Ensemble from XML, ensemble.cc / mart.cc:37-89). */
int qr_scorer_create(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F,
                     int device, qr_scorer **out);
/* The same with options.  QR_SCORER_CONDOP_WEIGHTS: every tree weight is replaced by the one a `ranker()` emitted by
 * the reference's conditional-operator generator uses — the weight as a float printed with three decimals and an `f`
 * suffix (src/io/generate_conditional_operators.cc:95-105) — so that the scores equal that generated code's
 * (compiled without floating-point contraction) bit for bit. */
#define QR_SCORER_CONDOP_WEIGHTS 1u
int qr_scorer_create_ex(const qr_flat_tree *trees, const double *weights, size_t ntrees, size_t F,
                        int device, unsigned flags, qr_scorer **out);
int qr_scorer_destroy(qr_scorer *s);
/* Replaces LTR_Algorithm::score_dataset (ltr_algorithm.cc:44-52): scores[i] = sum_t w_t*leaf_t(doc_i)
 * for row-major documents; host buffers, copies included. */
int qr_score_dataset(qr_scorer *s, const float *docs_rowmajor, size_t N, size_t F, double *scores);
/* Same with device-resident documents and scores (no copies). */
int qr_score_dataset_device(qr_scorer *s, const float *docs_rowmajor_device, size_t N, size_t F,
                            double *scores_device);
/* The per-tree score matrix partial[doc][tree] = (float) (weight_tree * leaf_tree(doc)) for row-major documents:
 * Ensemble::partial_scores_instance (src/learning/tree/ensemble.cc:121-131) for every document, cast to Feature as
 * Driver::extract_partial_scores does (src/driver/driver.cc:411-446) — the input of CLEAVER and of the line search.
 * ignore_weights = a scorer created with unit weights.  `scores` (may be NULL) also receives the ensemble scores. */
int qr_score_partial(qr_scorer *s, const float *docs_rowmajor, size_t N, size_t F, float *partial, double *scores);
/* Same with device-resident documents and outputs (asynchronous; `scores_device` may be NULL). */
int qr_score_partial_device(qr_scorer *s, const float *docs_rowmajor_device, size_t N, size_t F,
                            float *partial_device, double *scores_device);
/* ---- Line search over a score matrix (src/learning/linear/line_search.cc, the optimiser behind CLEAVER,
 * src/optimization/post_learning/cleaver/cleaver.cc:166-330) ----
 * The matrix is row-major [N][T]: one column per tree (qr_score_partial) or per feature.  The device does the passes
 * over documents — LineSearch::score / preCompute (line_search.cc:447-482), the candidate score vectors of step 1
 * (:252-272) and step 2 (:303-316) in the reference's own arithmetic, and Metric::evaluate_dataset (NDCG@cutoff,
 * metric.h:77-92) of each; the search itself stays on the host (quickrank_b200/linesearch.py). */
typedef struct qr_linesearch qr_linesearch;
int qr_ls_create(const float *x_rowmajor, size_t N, size_t T, const float *labels, const uint64_t *qoffsets, size_t Q,
                 uint32_t ndcg_cutoff, int device, qr_linesearch **out);
int qr_ls_destroy(qr_linesearch *ls);
/* metric of the ranking by sum_f weights[f] * x[.][f] (line_search.cc:215-219) */
int qr_ls_evaluate(qr_linesearch *ls, const double *weights, double *metric);
/* step 1 for column f: metrics[p] = metric with weights[f] replaced by points[p] (line_search.cc:252-281) */
int qr_ls_feature_points(qr_linesearch *ls, const double *weights, uint32_t f, const double *points, uint32_t npoints,
                         double *metrics);
/* step 2: metrics[p] = metric of weights + p * step, p = 0 .. npoints-1 (line_search.cc:303-325) */
int qr_ls_line_points(qr_linesearch *ls, const double *weights, const double *step, uint32_t npoints, double *metrics);
/* CLEAVER's quality-loss passes (quality_loss_pruning.cc:59-70, quality_loss_adv_pruning.cc:60-81): metrics[c] = metric
 * of the ensemble without column cols[c], i.e. of sum_f weights[f] * x[.][f] - weights[col] * x[.][col] */
int qr_ls_drop_points(qr_linesearch *ls, const double *weights, const uint32_t *cols, uint32_t ncols, double *metrics);
/* QUALITY_LOSS_ADV's running scores (quality_loss_adv_pruning.cc:88-92): the cached weighted sums of `weights` lose
 * column f in place (sum -= weights[f] * x[.][f]) and from now on stand for `weights` with weights[f] = 0 — the
 * sequentially updated vector the reference carries from one greedy step to the next, not a fresh sum */
int qr_ls_drop_column(qr_linesearch *ls, const double *weights, uint32_t f);
/* ScoreLossPruning (score_loss_pruning.cc:58-63): loss[f] = sum over documents, in document order, of
 * weights[f] * x[s][f] / (sum_g weights[g] * x[s][g]); loss has T entries */
int qr_ls_score_loss(qr_linesearch *ls, const double *weights, double *loss);
uint64_t qr_ls_launch_count(qr_linesearch *ls);

/* Waits for the scorer's stream (qr_score_dataset_device is asynchronous). */
int qr_scorer_sync(qr_scorer *s);
/* kernel launches issued by this scorer so far */
uint64_t qr_scorer_launch_count(qr_scorer *s);
/* CUDA-event stopwatch on the scorer's stream: stop == 0 records the start event, stop != 0 records
 * the end event, synchronises and returns the elapsed device time in milliseconds */
int qr_scorer_timer(qr_scorer *s, int stop, double *ms);
/* Replaces `double ranker(float *v)` (src/scoring/ranker.cc:23-25, called by quickscore.cc:103). */
int qr_score_document(qr_scorer *s, const float *doc, size_t F, double *score);

#ifdef __cplusplus
}
#endif
#endif
