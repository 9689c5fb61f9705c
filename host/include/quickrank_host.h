// quickrank_b200 host layer — the C++ side that sits above the C ABI (include/quickrank_b200.h).
//
// It mirrors the slice of the reference's API that surrounds the hot path (hpclab/quickrank @
// c569a59; paths below are relative to the reference root) with the same names, argument meaning
// and error behaviour (message on std::cerr, then exit(EXIT_FAILURE)):
//   quickrank::data::Dataset / VerticalDataset        include/data/dataset.h, vertical_dataset.h
//   quickrank::metric::ir::Metric / Dcg / Ndcg         include/metric/ir/{metric,dcg,ndcg}.h
//   quickrank::learning::LTR_Algorithm                 include/learning/ltr_algorithm.h
//   quickrank::learning::forests::Mart / LambdaMart / ObliviousMart / ObliviousLambdaMart
//                                                      include/learning/forests/*.h
//   RTNode / RegressionTree / Ensemble                 include/learning/tree/*.h
//   quickrank::io::Svml                                include/io/svml.h
// The bodies of the protected Mart hooks (init, clear, compute_pseudoresponses,
// fit_regressor_on_gradient, update_modelscores) and of score_dataset call the CUDA library; the
// boosting loop, the ensemble, early stopping, checkpoints and the XML model format stay on the
// host exactly as in the reference.  Written from scratch; nothing here is reference code.
#ifndef QUICKRANK_B200_HOST_H
#define QUICKRANK_B200_HOST_H

#include <cstddef>
#include <cstdint>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "quickrank_b200.h"

namespace quickrank {

typedef float Label;        // include/types.h:28-32
typedef double Score;
typedef float Feature;
typedef size_t QueryID;
typedef double MetricScore;

namespace data {

class QueryResults {
 public:
  QueryResults(size_t n_results, Label *labels, Feature *features)
      : num_results_(n_results), labels_(labels), features_(features) {}
  size_t num_results() const { return num_results_; }
  Label *labels() const { return labels_; }
  Feature *features() const { return features_; }
  // std::sort of positions by score, descending (queryresults.cc:47-53)
  void indexing_of_sorted_labels(const Score *scores, size_t *dest) const;
  void sorted_labels(const Score *scores, Label *dest, size_t cutoff) const;

 private:
  size_t num_results_;
  Label *labels_;
  Feature *features_;
};

// Row-major (documents x features) dataset; documents of a query are contiguous.
class Dataset {
 public:
  Dataset(size_t n_instances, size_t n_features);
  virtual ~Dataset();
  Dataset(const Dataset &) = delete;
  Dataset &operator=(const Dataset &) = delete;

  Feature *at(size_t document_id, size_t feature_id) { return data_ + document_id * num_features_ + feature_id; }
  const Feature *data() const { return data_; }
  const Label *labels() const { return labels_; }
  Label getLabel(size_t document_id) const { return labels_[document_id]; }
  size_t offset(size_t i) const { return offsets_[i]; }
  const std::vector<uint64_t> &offsets() const { return offsets_; }
  std::unique_ptr<QueryResults> getQueryResults(size_t i) const;
  void addInstance(QueryID q_id, Label i_label, const std::vector<Feature> &i_features);
  // bulk counterpart of addInstance for readers that already know the layout (binary cache): labels of all
  // max_instances rows and the query offsets; the feature matrix is then filled through at()
  void set_structure(const Label *labels, const std::vector<uint64_t> &offsets);
  size_t num_features() const { return num_features_; }
  size_t num_queries() const { return num_queries_; }
  size_t num_instances() const { return num_instances_; }

 private:
  size_t num_features_, num_queries_ = 0, num_instances_ = 0, max_instances_;
  Feature *data_ = nullptr;
  Label *labels_ = nullptr;
  std::vector<uint64_t> offsets_;
  QueryID last_instance_id_ = 0;
};

// Column-major view used during training.  The reference materialises the transpose on the host
// (vertical_dataset.cc:29-70); here the device does it, so this class only carries the shape, the
// labels and the query offsets, and produces a host transpose lazily if someone asks for at().
class VerticalDataset {
 public:
  explicit VerticalDataset(std::shared_ptr<Dataset> h_dataset);
  size_t num_features() const { return src_->num_features(); }
  size_t num_queries() const { return src_->num_queries(); }
  size_t num_instances() const { return src_->num_instances(); }
  size_t offset(size_t i) const { return src_->offset(i); }
  Label getLabel(size_t document_id) const { return src_->getLabel(document_id); }
  // vertical_dataset.cc:80-88: the query's labels and a pointer into the COLUMN-major matrix at its first
  // document (feature f of its document d is features()[f * num_instances() + d])
  std::unique_ptr<QueryResults> getQueryResults(size_t i);
  Feature *at(size_t document_id, size_t feature_id);   // data_[feature_id * N + document_id]
  std::shared_ptr<Dataset> horizontal() const { return src_; }

 private:
  std::shared_ptr<Dataset> src_;
  std::vector<Feature> col_;
};

}  // namespace data

namespace metric {
namespace ir {

class Metric {
 public:
  static const size_t NO_CUTOFF = SIZE_MAX;
  explicit Metric(size_t k = NO_CUTOFF) { set_cutoff(k); }
  virtual ~Metric() {}
  virtual std::string name() const = 0;
  size_t cutoff() const { return cutoff_; }
  void set_cutoff(size_t k) { cutoff_ = k == 0 ? NO_CUTOFF : k; }
  virtual MetricScore evaluate_result_list(const data::QueryResults *rl, const Score *scores) const = 0;
  // mean over queries, sequential (metric.h:77-106)
  virtual MetricScore evaluate_dataset(const std::shared_ptr<data::Dataset> dataset, const Score *scores) const;
  friend std::ostream &operator<<(std::ostream &os, const Metric &m) { return m.put(os); }

 private:
  size_t cutoff_;
  virtual std::ostream &put(std::ostream &os) const = 0;
};

class Dcg : public Metric {
 public:
  explicit Dcg(size_t k = NO_CUTOFF) : Metric(k) {}
  std::string name() const override { return "DCG"; }
  MetricScore evaluate_result_list(const data::QueryResults *rl, const Score *scores) const override;
  MetricScore compute_dcg(const Label *labels, size_t len) const;

 private:
  std::ostream &put(std::ostream &os) const override;
};

class Ndcg : public Dcg {
 public:
  explicit Ndcg(size_t k = NO_CUTOFF) : Dcg(k) {}
  std::string name() const override { return "NDCG"; }
  MetricScore evaluate_result_list(const data::QueryResults *rl, const Score *scores) const override;
  MetricScore compute_idcg(const data::QueryResults *rl) const;

 private:
  std::ostream &put(std::ostream &os) const override;
};

}  // namespace ir
}  // namespace metric

namespace io {

class Svml {
 public:
  // SVMLight / LETOR text: "<label> qid:<id> <fid>:<val> ... [# comment]" (svml.cc:38-161).
  // QR_SVML_CACHE=1: the parsed dataset is also written next to the text as <filename>.qrb and read back
  // from there (one sequential read instead of a parse) as long as the text's size and mtime are unchanged.
  std::unique_ptr<data::Dataset> read_horizontal(const std::string &filename);
  void write(std::shared_ptr<data::Dataset> dataset, const std::string &filename);
};

// XML model -> C source of `double ranker(float *v)`, the function quickscore times
// (src/scoring/ranker.cc is the stub it replaces at link time; quickscore.cc:100-106).
// generate_conditional_operators.cc:28-115: one nested `v[f] <= thr ? left : right` expression per tree,
// tree weights printed as float with 3 decimals.
class GenOpCond {
 public:
  void generate_conditional_operators_code(const std::string model_filename, const std::string code_filename);
};
// generate_oblivious.cc:137-330: per-tree arrays (weights, leaf outputs, feature ids, thresholds) sorted by
// tree depth and a `leaf_id` that packs the `v[f] > thr` bits MSB-first.
class GenOblivious {
 public:
  void generate_oblivious_code(const std::string model_filename, const std::string code_filename);
};
// generate_vpred.cc:92-172: the model as VPRED's text input — number of trees, then per tree its depth and a
// breadth-first list of "root" / "node" / "leaf" records (ids in visiting order, leaf outputs multiplied
// by the model's shrinkage), closed by "end".
class GenVpred {
 public:
  void generate_vpred_input(const std::string &ensemble_file, const std::string &output_file);
};

}  // namespace io

// Multi-GPU training (one process per GPU): every process loads the dataset, keeps
// a contiguous range of whole queries (lambdas need all documents of a query, lambdamart.cc:71-151) and
// creates its training context with qr_ctx_create_sharded; afterwards every process runs the SAME
// Mart::learn loop — the all-reduced histograms make every decision identical — and rank 0 reports.
namespace host {
struct Sharding {
  int rank = 0, world = 1;
  int local_rank = 0;            // CUDA device of this process
  std::string addr = "127.0.0.1";
  int port = 0;                  // TCP port rank 0 serves the communicator id on
};
void set_sharding(const Sharding &s);
const Sharding &sharding();
// Contiguous query ranges [q_begin, q_end) per rank, balanced by document count; the rule of
// quickrank_b200/sharding.py (boundary closest to N*r/world, at least one query per rank).
std::vector<std::pair<size_t, size_t>> query_shards(const uint64_t *offsets, size_t num_queries, int world);
// Rank 0 sends `id` (nbytes) to every other rank over TCP (addr:port); the others receive it.
// Returns false on failure (message on stderr).
bool exchange_bytes(unsigned char *id, size_t nbytes, const Sharding &s, int timeout_s = 120);
}  // namespace host
}  // namespace quickrank

// ---- tree structures (global namespace, as in the reference) -------------------------------------

static const size_t uint_max = (size_t) -1;

class RTNode {
 public:
  float threshold = 0.0f;
  double deviance = 0.0;
  double avglabel = 0.0;
  size_t nsampleids = 0;
  RTNode *left = nullptr;
  RTNode *right = nullptr;

  explicit RTNode(double prediction) { avglabel = prediction; }
  RTNode(float new_threshold, size_t new_featureidx, size_t new_featureid, RTNode *new_left, RTNode *new_right)
      : threshold(new_threshold), left(new_left), right(new_right), featureidx(new_featureidx),
        featureid(new_featureid) {}
  ~RTNode() { delete left; delete right; }
  void set_feature(size_t fidx, size_t fid) { featureidx = fidx; featureid = fid; }
  size_t get_feature_id() const { return featureid; }
  size_t get_feature_idx() const { return featureidx; }
  bool is_leaf() const { return featureidx == uint_max; }
  // rtnode.h:134-152
  quickrank::Score score_instance(const quickrank::Feature *d, const size_t next_fx_offset) const {
    return featureidx == uint_max ? avglabel
                                  : (d[featureidx * next_fx_offset] <= threshold ? left->score_instance(d, next_fx_offset)
                                                                                 : right->score_instance(d, next_fx_offset));
  }
  size_t count_nodes() const { return is_leaf() ? 1 : 1 + left->count_nodes() + right->count_nodes(); }

 private:
  size_t featureidx = uint_max;
  size_t featureid = uint_max;
};

// The product of fit_regressor_on_gradient: owns nothing but the root pointer until the ensemble
// takes it (mart.cc:342), like the reference's RegressionTree.
class RegressionTree {
 public:
  RegressionTree() {}
  explicit RegressionTree(RTNode *r) : root(r) {}
  RTNode *get_proot() const { return root; }
  // flat pre-order form <-> pointer graph
  static RTNode *from_flat(const qr_flat_tree &t);
  static void to_flat(const RTNode *root, std::vector<int32_t> &feature, std::vector<float> &threshold,
                      std::vector<int32_t> &left, std::vector<int32_t> &right, std::vector<double> &value);

 private:
  RTNode *root = nullptr;
};

class Ensemble {
 public:
  Ensemble() {}
  ~Ensemble();
  Ensemble(const Ensemble &) = delete;
  // ensemble.cc:53-62: move assignment, used by import_model_state to take over a loaded model's trees
  Ensemble &operator=(Ensemble &&other);
  void set_capacity(size_t n) { trees_.reserve(n); }
  void push(RTNode *root, double weight, float maxlabel);
  void pop();
  size_t get_size() const { return trees_.size(); }
  bool is_notempty() const { return !trees_.empty(); }
  RTNode *getTree(int index) const { return trees_[index].root; }
  double getWeight(int index) const { return trees_[index].weight; }
  // ensemble.cc:111-118
  quickrank::Score score_instance(const quickrank::Feature *d, size_t offset = 1) const;
  std::vector<double> get_weights() const;
  bool update_ensemble_weights(std::vector<double> &weights);
  // ensemble.cc:149-188: set the weights, optionally dropping the trees whose weight became 0
  bool update_ensemble_weights(std::vector<double> &weights, bool remove);
  bool filter_out_zero_weighted_trees();
  void write_xml(std::ostream &os, int indent) const;   // <ensemble>...</ensemble> (ensemble.cc:133-147)

 private:
  struct weighted_tree { RTNode *root; double weight; float maxlabel; };
  std::vector<weighted_tree> trees_;
};

namespace quickrank {
namespace learning {

class LTR_Algorithm {
 public:
  LTR_Algorithm() {}
  virtual ~LTR_Algorithm() {}
  LTR_Algorithm(const LTR_Algorithm &) = delete;
  virtual std::string name() const = 0;
  virtual void learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
                     std::shared_ptr<metric::ir::Metric> metric, size_t partial_save,
                     const std::string model_filename) = 0;
  // fills scores[N] for a row-major dataset (ltr_algorithm.cc:44-52)
  virtual void score_dataset(std::shared_ptr<data::Dataset> dataset, Score *scores) const;
  virtual Score score_document(const Feature *d) const = 0;
  // <name>.T<iter>.xml for partial saves (ltr_algorithm.cc:54-65)
  virtual void save(std::string model_filename, int suffix = -1) const;
  static std::shared_ptr<LTR_Algorithm> load_model_from_file(std::string model_filename);
  // --restart-train (ltr_algorithm_factory.cc:249-257): take over the trees of a loaded model when its
  // parameters are compatible with this object's; false = not compatible
  virtual bool import_model_state(LTR_Algorithm &other) { (void) other; return false; }
  virtual void write_xml_model(std::ostream &os) const = 0;
  friend std::ostream &operator<<(std::ostream &os, const LTR_Algorithm &a) { return a.put(os); }

 private:
  virtual std::ostream &put(std::ostream &os) const = 0;
};

namespace forests {

struct XmlModel;  // parsed <ranker> document (quickrank_host.cc)

class Mart : public LTR_Algorithm {
 public:
  // same parameter list as the reference (mart.h:52-66)
  Mart(size_t ntrees, double shrinkage, size_t nthresholds, size_t ntreeleaves, size_t minleafsupport,
       float subsample, float max_features, size_t valid_iterations, float collapse_leaves_factor)
      : ntrees_(ntrees), shrinkage_(shrinkage), nthresholds_(nthresholds), nleaves_(ntreeleaves),
        minleafsupport_(minleafsupport), subsample_(subsample), max_features_(max_features),
        valid_iterations_(valid_iterations), collapse_leaves_factor_(collapse_leaves_factor) {}
  explicit Mart(const XmlModel &model);
  virtual ~Mart();

  void learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
             std::shared_ptr<metric::ir::Metric> training_metric, size_t partial_save,
             const std::string output_basename) override;
  Score score_document(const Feature *d) const override { return ensemble_model_.score_instance(d, 1); }
  void score_dataset(std::shared_ptr<data::Dataset> dataset, Score *scores) const override;
  std::string name() const override { return NAME_; }
  void write_xml_model(std::ostream &os) const override;
  bool import_model_state(LTR_Algorithm &other) override;   // mart.cc:493-518
  std::vector<double> get_weights() const { return ensemble_model_.get_weights(); }
  const Ensemble &ensemble() const { return ensemble_model_; }
  // histogram accumulation mode of the CUDA library (QR_HIST_FAST / QR_HIST_REFERENCE)
  void set_hist_mode(uint32_t m) { hist_mode_ = m; }
  void set_device(int d) { device_ = d; }
  static const std::string NAME_;

 protected:
  // the hooks of mart.h:118-147
  virtual void init(std::shared_ptr<data::VerticalDataset> training_dataset);
  virtual void clear(size_t num_features);
  virtual void compute_pseudoresponses(std::shared_ptr<data::VerticalDataset> training_dataset,
                                       metric::ir::Metric *metric, bool *sample_presence);
  virtual std::unique_ptr<RegressionTree> fit_regressor_on_gradient(
      std::shared_ptr<data::VerticalDataset> training_dataset, size_t *sampleids);
  virtual void update_modelscores(std::shared_ptr<data::Dataset> dataset, Score *scores, RegressionTree *tree);
  virtual void update_modelscores(std::shared_ptr<data::VerticalDataset> dataset, Score *scores, RegressionTree *tree);
  virtual MetricScore evaluate_training(metric::ir::Metric *metric);
  // the two device calls behind the hooks above, on a context of the caller's choice (validation set, document sample)
  std::unique_ptr<RegressionTree> fit_tree_on(qr_ctx *ctx);
  void apply_tree_on(qr_ctx *target, RegressionTree *tree);
  virtual uint32_t algo_id() const { return QR_ALGO_MART; }
  virtual size_t tree_depth() const { return 0; }
  virtual void write_xml_info(std::ostream &os) const;
  std::ostream &put(std::ostream &os) const override;
  void die(const char *what) const;   // cerr + exit(EXIT_FAILURE), the reference's error convention

  Ensemble ensemble_model_;
  size_t ntrees_;
  double shrinkage_;
  size_t nthresholds_;
  size_t nleaves_;
  size_t minleafsupport_;
  float subsample_;
  float max_features_;
  size_t valid_iterations_;
  float collapse_leaves_factor_;
  MetricScore best_metric_on_training_ = 0, best_metric_on_validation_ = 0;
  size_t best_model_ = 0;
  uint32_t hist_mode_ = QR_HIST_FAST;
  int device_ = -1;
  size_t metric_cutoff_ = 10;
  qr_ctx *ctx_ = nullptr;         // training set on the GPU (Mart::init .. Mart::clear)
  qr_ctx *valid_ctx_ = nullptr;   // validation set binned with the training thresholds
  size_t shard_d0_ = 0;           // first document of this rank's shard (sharded training), set by init()
};

class LambdaMart : public Mart {
 public:
  using Mart::Mart;
  std::string name() const override { return NAME_; }
  static const std::string NAME_;

 protected:
  uint32_t algo_id() const override { return QR_ALGO_LAMBDAMART; }
};

class ObliviousMart : public Mart {
 public:
  ObliviousMart(size_t ntrees, double shrinkage, size_t nthresholds, size_t treedepth, size_t minleafsupport,
                float subsample, float max_features, size_t esr, float collapse_leaves_factor)
      : Mart(ntrees, shrinkage, nthresholds, (size_t) 1 << treedepth, minleafsupport, subsample, max_features, esr,
             collapse_leaves_factor), treedepth_(treedepth) {}
  explicit ObliviousMart(const XmlModel &model);
  bool import_model_state(LTR_Algorithm &other) override;   // obliviousmart.cc:88-106
  std::string name() const override { return NAME_; }
  static const std::string NAME_;

 protected:
  uint32_t algo_id() const override { return QR_ALGO_OBVMART; }
  size_t tree_depth() const override { return treedepth_; }
  void write_xml_info(std::ostream &os) const override;
  std::ostream &put(std::ostream &os) const override;
  size_t treedepth_;
};

class ObliviousLambdaMart : public ObliviousMart {
 public:
  using ObliviousMart::ObliviousMart;
  std::string name() const override { return NAME_; }
  static const std::string NAME_;

 protected:
  uint32_t algo_id() const override { return QR_ALGO_OBVLAMBDAMART; }
};

// LambdaMART on a per-query document sample that is redrawn as training goes on: the common loop of
// LambdaMartSelective::learn (lambdamartselective.cc:46-313) and StochasticNegative::learn
// (stochasticnegative.cc:46-283).  What touches documents runs on the GPU: the sample is a training context of
// its own (qr_ctx_create_sample) in which pseudo-responses, root histogram, tree fit and leaf outputs see the
// sampled documents only; the new tree is then applied to, and NDCG evaluated on, ALL documents in the full
// context.  Choosing the sample (sorts, quotas, shuffles) is host logic, in host/src/sampled_trainers.cc.
class SampledLambdaMart : public LambdaMart {
 public:
  using LambdaMart::LambdaMart;
  ~SampledLambdaMart() override;
  void learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
             std::shared_ptr<metric::ir::Metric> training_metric, size_t partial_save,
             const std::string output_basename) override;

 protected:
  // does this configuration sample at all (if not, the loop is plain LambdaMART on the full context)
  virtual bool sampling_enabled() const = 0;
  virtual void check_supported() const = 0;
  // is a new sample drawn before iteration m
  virtual bool resample_due(size_t m) const = 0;
  // draws a sample: `ids` arrives as 0..N-1 and leaves permuted with the sample in front; returns its size.
  // `scores` = the current model's scores of all training documents.
  virtual size_t draw_sample(const data::Dataset &dataset, const std::vector<Score> &scores,
                             const std::vector<size_t> &npositives, std::vector<size_t> &ids) = 0;
  virtual void before_training() {}
  virtual void after_iteration(size_t m, bool is_best) { (void) m; (void) is_best; }
  void clear(size_t num_features) override;

 private:
  void build_sample_context(const data::Dataset &dataset, const std::vector<size_t> &ids, size_t n);
  qr_ctx *sample_ctx_ = nullptr;
};

// lambdamartselective.h:34-107
class LambdaMartSelective : public SampledLambdaMart {
 public:
  LambdaMartSelective(size_t ntrees, double shrinkage, size_t nthresholds, size_t ntreeleaves, size_t minleafsupport,
                      float subsample, float max_features, size_t esr, float collapse_leaves_factor,
                      int sampling_iterations, float max_sampling_factor, float random_sampling_factor,
                      float normalization_factor, std::string adaptive_strategy, std::string negative_strategy)
      : SampledLambdaMart(ntrees, shrinkage, nthresholds, ntreeleaves, minleafsupport, subsample, max_features, esr,
                          collapse_leaves_factor),
        sampling_iterations(sampling_iterations), rank_sampling_factor(max_sampling_factor),
        random_sampling_factor(random_sampling_factor), normalization_factor(normalization_factor),
        adaptive_strategy(std::move(adaptive_strategy)), negative_strategy(std::move(negative_strategy)) {}
  explicit LambdaMartSelective(const XmlModel &model) : SampledLambdaMart(model) {}
  std::string name() const override { return NAME_; }
  static const std::string NAME_;
  // LambdaMartSelective::sampling_query_level (lambdamartselective.cc:326-493), public for host/selective_check.cc
  size_t sampling_query_level(const data::Dataset &dataset, const std::vector<Score> &scores,
                              const std::vector<size_t> &npositives, std::vector<size_t> &ids, float adapt_factor);
  // rand() calls the draws of this process have made so far (a later draw continues the same stream)
  static size_t rand_calls();

 protected:
  bool sampling_enabled() const override { return rank_sampling_factor > 0 || random_sampling_factor > 0; }
  void check_supported() const override;
  bool resample_due(size_t m) const override { return m > 0 && m % (size_t) sampling_iterations == 0; }
  size_t draw_sample(const data::Dataset &dataset, const std::vector<Score> &scores,
                     const std::vector<size_t> &npositives, std::vector<size_t> &ids) override {
    return sampling_query_level(dataset, scores, npositives, ids, adapt_factor_);
  }
  void before_training() override;
  void after_iteration(size_t m, bool is_best) override;
  std::ostream &put(std::ostream &os) const override;

 private:
  int sampling_iterations = 0;
  float rank_sampling_factor = 1.0f, random_sampling_factor = 0.0f, normalization_factor = 100.0f;
  std::string adaptive_strategy = "NO", negative_strategy = "RATIO";
  std::vector<bool> improvements_;
  float adapt_factor_ = 1.0f;
};

// stochasticnegative.h: every positive document and, per query, a random share `subsample` of the negatives,
// redrawn at every iteration.  The reference seeds the shuffle from the wall clock (stochasticnegative.cc:315-316),
// so two of its own runs differ; here the stream is std::default_random_engine seeded by --seed (default 0) + the
// number of draws so far, which keeps a run reproducible.
class StochasticNegative : public SampledLambdaMart {
 public:
  using SampledLambdaMart::SampledLambdaMart;
  std::string name() const override { return NAME_; }
  static const std::string NAME_;
  void set_seed(unsigned long long s) { seed_ = s; }

 protected:
  bool sampling_enabled() const override { return subsample_ != 1.0f; }
  void check_supported() const override;
  bool resample_due(size_t m) const override { return m > 0; }
  size_t draw_sample(const data::Dataset &dataset, const std::vector<Score> &scores,
                     const std::vector<size_t> &npositives, std::vector<size_t> &ids) override;

 private:
  unsigned long long seed_ = 0, draws_ = 0;
};

// DART on top of LambdaMART (dart.h:34-199, dart.cc).  The dropout selection, the weight
// normalisation and the bookkeeping of Dart::learn (dart.cc:172-602) are host logic and are
// replicated here, std::rand() stream included; everything that touches documents — pseudo-responses,
// tree fit, adding / subtracting the contribution of a set of trees to all scores
// (Dart::update_modelscores, dart.cc:634-687), NDCG — runs on the GPU through the C ABI.
// Supported: sample type UNIFORM / TOP_FIFTY, normalisation TREE / NONE / WEIGHTED / FOREST /
// TREE_BOOST3, adaptive type FIXED (the reference's defaults and BASELINE.json config 5); the
// contribution-based and line-search variants are rejected at construction.
class Dart : public LambdaMart {
 public:
  enum class SamplingType { UNIFORM, WEIGHTED, WEIGHTED_INV, COUNT2, COUNT3, COUNT2N, COUNT3N, TOP_FIFTY, CONTR,
                            CONTR_INV, WCONTR, WCONTR_INV, TOP_WCONTR, LESS_WCONTR };
  enum class NormalizationType { TREE, NONE, WEIGHTED, FOREST, TREE_ADAPTIVE, LINESEARCH, TREE_BOOST3, CONTR, WCONTR,
                                 LMART_ADAPTIVE };
  enum class AdaptiveType { FIXED, PLUS1_DIV2, PLUSHALF_DIV2, PLUSONETHIRD_DIV2, PLUSHALF_RESET, PLUSHALF_RESET_LB1_UB5,
                            PLUSHALF_RESET_LB1_UB10, PLUSHALF_RESET_LB1_UBRD };
  // same parameter list as the reference (dart.h:62-70)
  Dart(size_t ntrees, double shrinkage, size_t nthresholds, size_t ntreeleaves, size_t minleafsupport, float subsample,
       float max_features, size_t valid_iterations, float collapse_leaves_factor, SamplingType sample_type,
       NormalizationType normalize_type, AdaptiveType adaptive_rate, double rate_drop, double skip_drop, bool keep_drop,
       bool best_on_train, double random_keep, double drop_on_best);
  explicit Dart(const XmlModel &model);
  bool import_model_state(LTR_Algorithm &other) override;   // dart.cc:604-632
  void learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
             std::shared_ptr<metric::ir::Metric> training_metric, size_t partial_save,
             const std::string output_basename) override;
  std::string name() const override { return NAME_; }
  static const std::string NAME_;
  static SamplingType get_sampling_type(std::string name);
  static NormalizationType get_normalization_type(std::string name);
  static AdaptiveType get_adaptive_type(std::string name);
  static std::string get_sampling_type(SamplingType t);
  static std::string get_normalization_type(NormalizationType t);
  static std::string get_adaptive_type(AdaptiveType t);
  // every metric value Dart::learn asked for, in call order (parity tap for the tests)
  const std::vector<MetricScore> &metric_trace() const { return metric_trace_; }

 protected:
  void write_xml_info(std::ostream &os) const override;
  std::ostream &put(std::ostream &os) const override;
  // Dart::update_modelscores (dart.cc:634-687): scores += sign * weight_t * tree_t(doc) for t in trees
  void update_modelscores_trees(qr_ctx *ctx, bool add, const std::vector<int> &trees);
  std::vector<int> select_trees_to_dropout(std::vector<double> &weights, size_t trees_to_dropout);
  void normalize_trees_restore_drop(std::vector<double> &weights, const std::vector<int> &dropped_trees,
                                    double last_tree_weight);
  int get_number_of_trees_to_dropout(std::vector<double> &dropout_factor_per_iter, int dropped_before_cleaning);
  void rescore_on_device(qr_ctx *ctx, std::shared_ptr<data::Dataset> dataset);
  void check_supported() const;

  SamplingType sample_type = SamplingType::UNIFORM;
  NormalizationType normalize_type = NormalizationType::TREE;
  AdaptiveType adaptive_type = AdaptiveType::FIXED;
  double rate_drop = 0.1, skip_drop = 0.0;
  bool keep_drop = false, best_on_train = false;
  double random_keep = 0.0;
  bool drop_on_best = false;

 private:
  struct DeviceTree;                                  // flat form of one ensemble tree on this run's bins
  std::vector<std::shared_ptr<DeviceTree>> flat_;     // parallel to ensemble_model_
  std::shared_ptr<DeviceTree> make_flat(const RTNode *root) const;
  std::vector<MetricScore> metric_trace_;
};

}  // namespace forests
}  // namespace learning
}  // namespace quickrank

#endif
