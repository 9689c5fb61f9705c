// TEST INFRASTRUCTURE — not part of the product.
//
// extern "C" harness around the UNMODIFIED reference sources (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libqr_ref.so).  It
// drives the reference's own classes through their protected virtual hooks
// (include/learning/forests/mart.h:118-147 of the reference) and records what
// each hook produced, so that tests can compare the CUDA path and the C
// restatement (oracle/qr_oracle.c) against the real thing.
//
// Nothing here re-implements reference arithmetic: every number returned is
// produced by reference code.
#include <omp.h>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <sstream>
#include <iostream>

#include "data/dataset.h"
#include "data/vertical_dataset.h"
#include "metric/ir/ndcg.h"
#include "metric/ir/dcg.h"
#include "learning/ltr_algorithm.h"
#include "learning/forests/mart.h"
#include "learning/linear/line_search.h"
#include "optimization/post_learning/cleaver/cleaver.h"
#include "optimization/post_learning/cleaver/last_pruning.h"
#include "optimization/post_learning/cleaver/skip_pruning.h"
#include "optimization/post_learning/cleaver/low_weights_pruning.h"
#include "optimization/post_learning/cleaver/quality_loss_pruning.h"
#include "optimization/post_learning/cleaver/quality_loss_adv_pruning.h"
#include "optimization/post_learning/cleaver/score_loss_pruning.h"
#include "learning/forests/lambdamart.h"
#include "learning/forests/obliviousmart.h"
#include "learning/forests/obliviouslambdamart.h"
#include "io/generate_conditional_operators.h"
#include "io/generate_oblivious.h"
#include "io/generate_vpred.h"
#include "io/svml.h"
#include "learning/forests/dart.h"
#include "learning/forests/lambdamartselective.h"
#include "utils/radix.h"
#include "driver/driver.h"

// Link-time placeholder: src/driver/driver.cc needs the absent ParamsMap library and is not on
// the hot path; CLEAVER/MetaCleaver (out of scope) reference this one symbol.  Never called here.
namespace quickrank { namespace driver {
std::shared_ptr<data::Dataset> Driver::extract_partial_scores(
    std::shared_ptr<learning::LTR_Algorithm>, std::shared_ptr<data::Dataset>, bool) {
  std::cerr << "oracle: Driver::extract_partial_scores is not built" << std::endl;
  abort();
}
} }

using namespace quickrank;
using quickrank::learning::forests::Mart;
using quickrank::learning::forests::LambdaMart;
using quickrank::learning::forests::ObliviousMart;
using quickrank::learning::forests::ObliviousLambdaMart;
using quickrank::learning::forests::Dart;
using quickrank::learning::forests::LambdaMartSelective;

namespace {

// Flat pre-order (left first) tree, the exchange format shared with the
// product's C ABI (include/quickrank_b200.h: qr_flat_tree).
struct FlatTree {
  std::vector<int32_t> feature;      // -1 for leaves
  std::vector<uint32_t> threshold_idx;
  std::vector<float> threshold;
  std::vector<int32_t> left, right;
  std::vector<double> value;         // RTNode::avglabel
  std::vector<double> deviance;
  std::vector<uint64_t> count;       // RTNode::nsampleids (0 when not kept)
  double weight = 0;
};

struct Recorder {
  bool keep_gradients = false;
  std::vector<std::vector<float>> thresholds;
  std::vector<FlatTree> trees;
  std::vector<std::vector<double>> lambdas, weights, scores;
  std::vector<double> metric_train;
};

static uint32_t lookup_threshold(const std::vector<float> &thr, float v) {
  for (size_t i = 0; i < thr.size(); ++i)
    if (thr[i] == v) return (uint32_t) i;
  return UINT32_MAX;
}

static int flatten(RTNode *n, FlatTree &t, const std::vector<std::vector<float>> &thr) {
  int id = (int) t.feature.size();
  t.feature.push_back(-1);
  t.threshold_idx.push_back(UINT32_MAX);
  t.threshold.push_back(0.f);
  t.left.push_back(-1);
  t.right.push_back(-1);
  t.value.push_back(n->avglabel);
  t.deviance.push_back(n->deviance);
  t.count.push_back(n->nsampleids);
  if (!n->is_leaf()) {
    size_t f = n->get_feature_idx();
    t.feature[id] = (int32_t) f;
    t.threshold[id] = n->threshold;
    if (f < thr.size()) t.threshold_idx[id] = lookup_threshold(thr[f], n->threshold);
    int l = flatten(n->left, t, thr);
    t.left[id] = l;
    int r = flatten(n->right, t, thr);
    t.right[id] = r;
  }
  return id;
}

// Subclass that exposes and records the protected hooks.
template<class Base>
class Trace : public Base {
 public:
  using Base::Base;
  Recorder rec;

  // ---- recorded overrides (used by the real Base::learn loop) ----
  void init(std::shared_ptr<data::VerticalDataset> d) override {
    Base::init(d);
    rec.thresholds.resize(d->num_features());
    for (size_t f = 0; f < d->num_features(); ++f)
      rec.thresholds[f].assign(this->thresholds_[f],
                               this->thresholds_[f] + this->thresholds_size_[f]);
    n_ = d->num_instances();
  }
  void compute_pseudoresponses(std::shared_ptr<data::VerticalDataset> d,
                               metric::ir::Metric *m, bool *sp) override {
    Base::compute_pseudoresponses(d, m, sp);
    if (rec.keep_gradients) {
      rec.lambdas.emplace_back(this->pseudoresponses_, this->pseudoresponses_ + n_);
      if (weights_ptr())
        rec.weights.emplace_back(weights_ptr(), weights_ptr() + n_);
    }
  }
  std::unique_ptr<RegressionTree> fit_regressor_on_gradient(
      std::shared_ptr<data::VerticalDataset> d, size_t *ids) override {
    auto t = Base::fit_regressor_on_gradient(d, ids);
    FlatTree ft;
    flatten(t->get_proot(), ft, rec.thresholds);
    ft.weight = this->shrinkage_;
    rec.trees.push_back(std::move(ft));
    return t;
  }
  void update_modelscores(std::shared_ptr<data::VerticalDataset> d, Score *s,
                          RegressionTree *t) override {
    Mart::update_modelscores(d, s, t);
    if (rec.keep_gradients && s == this->scores_on_training_)
      rec.scores.emplace_back(s, s + n_);
  }
  void update_modelscores(std::shared_ptr<data::Dataset> d, Score *s,
                          RegressionTree *t) override {
    Mart::update_modelscores(d, s, t);
  }

  // ---- step-wise access for stage-level parity ----
  void x_init(std::shared_ptr<data::VerticalDataset> d) {
    this->init(d);
    ids_.resize(n_);
    for (size_t i = 0; i < n_; ++i) ids_[i] = i;
    this->ensemble_model_.set_capacity(this->ntrees_);
  }
  void x_clear(size_t nf) { this->clear(nf); }
  double *x_scores() { return this->scores_on_training_; }
  double *x_lambdas() { return this->pseudoresponses_; }
  double *x_weights() { return weights_ptr(); }
  void x_pseudo(std::shared_ptr<data::VerticalDataset> d, metric::ir::Metric *m) {
    this->compute_pseudoresponses(d, m, NULL);
  }
  // the document-sampling trainers' call (lambdamartselective.cc:194): sample_presence[doc]
  void x_pseudo_masked(std::shared_ptr<data::VerticalDataset> d, metric::ir::Metric *m, bool *presence) {
    this->compute_pseudoresponses(d, m, presence);
  }
  // mart.cc:335-345 of the reference, in that order
  void x_fit_and_update(std::shared_ptr<data::VerticalDataset> d, bool update) {
    this->hist_->update(this->pseudoresponses_, n_, ids_.data());
    std::unique_ptr<RegressionTree> t = this->fit_regressor_on_gradient(d, ids_.data());
    this->ensemble_model_.push(t->get_proot(), this->shrinkage_, 0);
    if (update) this->update_modelscores(d, this->scores_on_training_, t.get());
  }
  size_t n() const { return n_; }

 private:
  double *weights_ptr() {
    if constexpr (std::is_base_of<LambdaMart, Base>::value)
      return this->instance_weights_;
    else
      return nullptr;
  }
  size_t n_ = 0;
  std::vector<size_t> ids_;
};

class TraceNdcg : public metric::ir::Ndcg {
 public:
  explicit TraceNdcg(size_t k) : Ndcg(k) {}
  std::vector<double> *sink = nullptr;
  MetricScore evaluate_dataset(const std::shared_ptr<data::VerticalDataset> d,
                               const Score *s) const override {
    MetricScore v = Ndcg::evaluate_dataset(d, s);
    if (sink) sink->push_back(v);
    return v;
  }
  MetricScore evaluate_dataset(const std::shared_ptr<data::Dataset> d,
                               const Score *s) const override {
    return Ndcg::evaluate_dataset(d, s);
  }
};

enum Algo { A_MART = 0, A_LAMBDAMART = 1, A_OBVMART = 2, A_OBVLAMBDAMART = 3, A_DART = 4, A_SELECTIVE = 5 };

struct Session {
  int algo;
  std::shared_ptr<data::Dataset> ds;
  std::shared_ptr<data::VerticalDataset> vds;
  std::shared_ptr<TraceNdcg> metric;
  std::unique_ptr<Trace<Mart>> mart;
  std::unique_ptr<Trace<LambdaMart>> lmart;
  std::unique_ptr<Trace<ObliviousMart>> omart;
  std::unique_ptr<Trace<ObliviousLambdaMart>> olmart;
  std::unique_ptr<Trace<Dart>> dart;
  std::unique_ptr<Trace<LambdaMartSelective>> sel;
  std::string log;   // what the last quiet learn() printed
  bool inited = false;
  Recorder *rec() {
    switch (algo) {
      case A_MART: return &mart->rec;
      case A_LAMBDAMART: return &lmart->rec;
      case A_OBVMART: return &omart->rec;
      case A_OBVLAMBDAMART: return &olmart->rec;
      case A_SELECTIVE: return &sel->rec;
      default: return &dart->rec;
    }
  }
  learning::LTR_Algorithm *ltr() {
    switch (algo) {
      case A_MART: return mart.get();
      case A_LAMBDAMART: return lmart.get();
      case A_OBVMART: return omart.get();
      case A_OBVLAMBDAMART: return olmart.get();
      case A_SELECTIVE: return sel.get();
      default: return dart.get();
    }
  }
};

#define DISPATCH(s, expr)                       \
  switch ((s)->algo) {                          \
    case A_MART: { auto &a = *(s)->mart; expr; } break;          \
    case A_LAMBDAMART: { auto &a = *(s)->lmart; expr; } break;   \
    case A_OBVMART: { auto &a = *(s)->omart; expr; } break;      \
    case A_OBVLAMBDAMART: { auto &a = *(s)->olmart; expr; } break; \
    case A_SELECTIVE: { auto &a = *(s)->sel; expr; } break;      \
    default: { auto &a = *(s)->dart; expr; } break;              \
  }

struct Silence {
  std::streambuf *old;
  std::ostringstream sink;
  bool on;
  explicit Silence(bool quiet) : on(quiet) { if (on) old = std::cout.rdbuf(sink.rdbuf()); }
  ~Silence() { if (on) std::cout.rdbuf(old); }
};

}  // namespace

extern "C" {

struct qref_params {
  int32_t algo;              // Algo
  uint64_t ntrees;
  double shrinkage;
  uint64_t nthresholds;      // 0 = unlimited
  uint64_t nleaves;          // leaf-wise algos
  uint64_t treedepth;        // oblivious algos
  uint64_t minleafsupport;
  uint64_t cutoff;           // NDCG@k, 0 = no cutoff
  // DART only
  int32_t dart_sample_type, dart_normalize_type, dart_adaptive_type;
  double dart_rate_drop, dart_skip_drop;
  int32_t dart_keep_drop, dart_best_on_train;
  double dart_random_keep, dart_drop_on_best;
  // LAMBDAMART-SELECTIVE only (lambdamartselective.h:47-61); strategies by index: NO FIXED RATIO MIX / RATIO MUL POS
  int32_t sel_sampling_iterations, sel_adaptive, sel_negative, sel_pad;
  double sel_rank_factor, sel_random_factor, sel_normalization_factor;
};

static const char *kSelAdaptive[] = {"NO", "FIXED", "RATIO", "MIX"};
static const char *kSelNegative[] = {"RATIO", "MUL", "POS"};

void *qref_open(const qref_params *p, const float *rowmajor, const float *labels,
                const uint64_t *qoffsets, uint64_t N, uint64_t F, uint64_t Q) {
  auto *s = new Session();
  s->algo = p->algo;
  s->ds = std::make_shared<data::Dataset>(N, F);
  for (uint64_t q = 0; q < Q; ++q)
    for (uint64_t i = qoffsets[q]; i < qoffsets[q + 1]; ++i)
      s->ds->addInstance((QueryID) (q + 1), labels[i],
                         std::vector<Feature>(rowmajor + i * F, rowmajor + (i + 1) * F));
  s->metric = std::make_shared<TraceNdcg>(p->cutoff);
  switch (p->algo) {
    case A_MART:
      s->mart.reset(new Trace<Mart>(p->ntrees, p->shrinkage, p->nthresholds, p->nleaves,
                                    p->minleafsupport, 1.0f, 1.0f, 0, 0.0f));
      break;
    case A_LAMBDAMART:
      s->lmart.reset(new Trace<LambdaMart>(p->ntrees, p->shrinkage, p->nthresholds, p->nleaves,
                                           p->minleafsupport, 1.0f, 1.0f, 0, 0.0f));
      break;
    case A_OBVMART:
      s->omart.reset(new Trace<ObliviousMart>(p->ntrees, p->shrinkage, p->nthresholds,
                                              p->treedepth, p->minleafsupport, 1.0f, 1.0f, 0,
                                              0.0f));
      break;
    case A_OBVLAMBDAMART:
      s->olmart.reset(new Trace<ObliviousLambdaMart>(p->ntrees, p->shrinkage, p->nthresholds,
                                                     p->treedepth, p->minleafsupport, 1.0f, 1.0f,
                                                     0, 0.0f));
      break;
    case A_DART:
      s->dart.reset(new Trace<Dart>(
          p->ntrees, p->shrinkage, p->nthresholds, p->nleaves, p->minleafsupport, 1.0f, 1.0f, 0,
          0.0f, (Dart::SamplingType) p->dart_sample_type,
          (Dart::NormalizationType) p->dart_normalize_type,
          (Dart::AdaptiveType) p->dart_adaptive_type, p->dart_rate_drop, p->dart_skip_drop,
          p->dart_keep_drop != 0, p->dart_best_on_train != 0, p->dart_random_keep,
          p->dart_drop_on_best));
      break;
    case A_SELECTIVE:
      s->sel.reset(new Trace<LambdaMartSelective>(
          p->ntrees, p->shrinkage, p->nthresholds, p->nleaves, p->minleafsupport, 1.0f, 1.0f, 0, 0.0f,
          p->sel_sampling_iterations, (float) p->sel_rank_factor, (float) p->sel_random_factor,
          (float) p->sel_normalization_factor, kSelAdaptive[p->sel_adaptive & 3], kSelNegative[p->sel_negative % 3]));
      break;
    default:
      delete s;
      return nullptr;
  }
  return s;
}

void qref_close(void *h) {
  auto *s = (Session *) h;
  if (s->inited) { DISPATCH(s, a.x_clear(s->vds->num_features())); }
  delete s;
}

// Runs the reference's own learn() loop (mart.cc:208-416 / dart.cc:172-602) and
// records trees (+ gradients and scores when keep_gradients != 0).
int qref_learn(void *h, int keep_gradients, int quiet) {
  auto *s = (Session *) h;
  s->rec()->keep_gradients = keep_gradients != 0;
  s->metric->sink = &s->rec()->metric_train;
  {
    Silence sil(quiet != 0);
    s->ltr()->learn(s->ds, nullptr, s->metric, 0, "");
    if (quiet) s->log = sil.sink.str();
  }
  s->metric->sink = nullptr;
  return 0;
}

// what that learn() wrote to stdout (quiet runs); returns the length
uint64_t qref_log(void *h, char *buf, uint64_t cap) {
  auto *s = (Session *) h;
  if (buf && cap) {
    const uint64_t n = std::min<uint64_t>(cap - 1, s->log.size());
    memcpy(buf, s->log.data(), n);
    buf[n] = 0;
  }
  return s->log.size();
}

// ---- step-wise protocol (same call order as mart.cc:307-347) ----
int qref_init(void *h) {
  auto *s = (Session *) h;
  s->vds = std::make_shared<data::VerticalDataset>(s->ds);
  DISPATCH(s, a.x_init(s->vds));
  s->inited = true;
  return 0;
}
void qref_set_scores(void *h, const double *v) {
  auto *s = (Session *) h;
  DISPATCH(s, memcpy(a.x_scores(), v, a.n() * sizeof(double)));
}
void qref_get_scores(void *h, double *v) {
  auto *s = (Session *) h;
  DISPATCH(s, memcpy(v, a.x_scores(), a.n() * sizeof(double)));
}
void qref_compute_pseudoresponses(void *h) {
  auto *s = (Session *) h;
  DISPATCH(s, a.x_pseudo(s->vds, s->metric.get()));
}
void qref_compute_pseudoresponses_masked(void *h, const uint8_t *presence) {
  auto *s = (Session *) h;
  DISPATCH(s, {
    std::unique_ptr<bool[]> sp(new bool[a.n()]);
    for (size_t i = 0; i < a.n(); ++i) sp[i] = presence[i] != 0;
    a.x_pseudo_masked(s->vds, s->metric.get(), sp.get());
  });
}
void qref_get_gradients(void *h, double *lambdas, double *weights) {
  auto *s = (Session *) h;
  DISPATCH(s, {
    memcpy(lambdas, a.x_lambdas(), a.n() * sizeof(double));
    if (weights && a.x_weights()) memcpy(weights, a.x_weights(), a.n() * sizeof(double));
  });
}
void qref_set_gradients(void *h, const double *lambdas, const double *weights) {
  auto *s = (Session *) h;
  DISPATCH(s, {
    memcpy(a.x_lambdas(), lambdas, a.n() * sizeof(double));
    if (weights && a.x_weights()) memcpy(a.x_weights(), weights, a.n() * sizeof(double));
  });
}
// root-histogram refresh + fit + leaf outputs (+ score update): mart.cc:335-345
int qref_fit_tree(void *h, int update_scores) {
  auto *s = (Session *) h;
  DISPATCH(s, a.x_fit_and_update(s->vds, update_scores != 0));
  return (int) s->rec()->trees.size() - 1;
}
double qref_evaluate(void *h) {
  auto *s = (Session *) h;
  double r = 0;
  DISPATCH(s, r = s->metric->metric::ir::Ndcg::evaluate_dataset(s->vds, a.x_scores()));
  return r;
}

// ---- recorded results ----
uint64_t qref_num_trees(void *h) { return ((Session *) h)->rec()->trees.size(); }
uint64_t qref_tree_nodes(void *h, uint64_t t) { return ((Session *) h)->rec()->trees[t].feature.size(); }
void qref_tree_get(void *h, uint64_t t, int32_t *feature, uint32_t *thr_idx, float *thr,
                   int32_t *left, int32_t *right, double *value, double *deviance,
                   uint64_t *count, double *weight) {
  const FlatTree &ft = ((Session *) h)->rec()->trees[t];
  size_t n = ft.feature.size();
  if (feature) memcpy(feature, ft.feature.data(), n * sizeof(int32_t));
  if (thr_idx) memcpy(thr_idx, ft.threshold_idx.data(), n * sizeof(uint32_t));
  if (thr) memcpy(thr, ft.threshold.data(), n * sizeof(float));
  if (left) memcpy(left, ft.left.data(), n * sizeof(int32_t));
  if (right) memcpy(right, ft.right.data(), n * sizeof(int32_t));
  if (value) memcpy(value, ft.value.data(), n * sizeof(double));
  if (deviance) memcpy(deviance, ft.deviance.data(), n * sizeof(double));
  if (count) memcpy(count, ft.count.data(), n * sizeof(uint64_t));
  if (weight) *weight = ft.weight;
}
uint64_t qref_thresholds_size(void *h, uint64_t f) { return ((Session *) h)->rec()->thresholds[f].size(); }
void qref_thresholds_get(void *h, uint64_t f, float *out) {
  auto &v = ((Session *) h)->rec()->thresholds[f];
  memcpy(out, v.data(), v.size() * sizeof(float));
}
uint64_t qref_num_metric(void *h) { return ((Session *) h)->rec()->metric_train.size(); }
void qref_metric_get(void *h, double *out) {
  auto &v = ((Session *) h)->rec()->metric_train;
  memcpy(out, v.data(), v.size() * sizeof(double));
}
// iteration-indexed recordings (keep_gradients != 0); kind: 0 lambdas, 1 weights, 2 scores
uint64_t qref_num_recorded(void *h, int kind) {
  auto *r = ((Session *) h)->rec();
  return kind == 0 ? r->lambdas.size() : kind == 1 ? r->weights.size() : r->scores.size();
}
void qref_recorded_get(void *h, int kind, uint64_t it, double *out) {
  auto *r = ((Session *) h)->rec();
  auto &v = kind == 0 ? r->lambdas[it] : kind == 1 ? r->weights[it] : r->scores[it];
  memcpy(out, v.data(), v.size() * sizeof(double));
}

// ---- the reference's C code generators (src/io/generate_*.cc); kind 0: condop, 1: oblivious, 2: vpred ----
int qref_generate_code(const char *model, const char *code, int kind) {
  if (kind == 0) quickrank::io::GenOpCond().generate_conditional_operators_code(model, code);
  else if (kind == 2) quickrank::io::GenVpred().generate_vpred_input(model, code);
  else quickrank::io::GenOblivious().generate_oblivious_code(model, code);
  return 0;
}

// ---- model I/O and scoring through the reference's own code ----
// Saves the model currently held by the session (ltr_algorithm.cc:54-65).
int qref_save_model(void *h, const char *path) {
  ((Session *) h)->ltr()->save(path);
  return 0;
}
// Loads an XML model with the reference loader (ltr_algorithm.cc:67-128) and
// scores a row-major dataset with LTR_Algorithm::score_dataset (ltr_algorithm.cc:44-52).
int qref_score_with_model(const char *xml_path, const float *rowmajor, uint64_t N, uint64_t F,
                          double *scores) {
  auto model = learning::LTR_Algorithm::load_model_from_file(xml_path);
  if (!model) return 1;
  auto ds = std::make_shared<data::Dataset>(N, F);
  std::vector<Feature> row(F);
  for (uint64_t i = 0; i < N; ++i) {
    row.assign(rowmajor + i * F, rowmajor + (i + 1) * F);
    ds->addInstance(1, 0.f, row);
  }
  model->score_dataset(ds, scores);
  return 0;
}

// ---- pure functions pinned by the reference's own unit tests ----
double qref_dcg(const float *labels, const double *scores, uint64_t n, uint64_t cutoff) {
  metric::ir::Dcg m(cutoff);
  data::QueryResults qr(n, const_cast<float *>(labels), NULL);
  return m.evaluate_result_list(&qr, scores);
}
double qref_ndcg(const float *labels, const double *scores, uint64_t n, uint64_t cutoff) {
  metric::ir::Ndcg m(cutoff);
  data::QueryResults qr(n, const_cast<float *>(labels), NULL);
  return m.evaluate_result_list(&qr, scores);
}
// full Jacobian (packed upper-triangular incl. diagonal, symmatrix.h layout) for one list
void qref_ndcg_jacobian(const float *labels, const double *scores, uint64_t n, uint64_t cutoff,
                        double *out_packed) {
  metric::ir::Ndcg m(cutoff);
  auto qr = std::make_shared<data::QueryResults>(n, const_cast<float *>(labels), nullptr);
  auto ranked = std::make_shared<data::RankedResults>(qr, const_cast<double *>(scores));
  auto jac = m.jacobian(ranked);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i)
    for (size_t j = i; j < n; ++j) out_packed[k++] = jac->at(i, j);
}
// the permutation std::sort produces for these scores (queryresults.cc:47-53)
void qref_sort_indices(const double *scores, uint64_t n, uint64_t *dest) {
  std::vector<float> dummy(n, 0.f);
  data::QueryResults qr(n, dummy.data(), NULL);
  std::vector<size_t> d(n);
  qr.indexing_of_sorted_labels(scores, d.data());
  for (uint64_t i = 0; i < n; ++i) dest[i] = d[i];
}
// idx_radixsort (utils/radix.cc:35-73): stable ascending argsort of floats
void qref_radix_argsort(const float *v, uint64_t n, uint64_t *dest) {
  auto r = idx_radixsort(v, n);
  for (uint64_t i = 0; i < n; ++i) dest[i] = r[i];
}

// OpenMP team size of the reference's loops (a launcher such as torchrun exports OMP_NUM_THREADS=1: the
// benchmark's reference legs set the team size explicitly and report what the runtime then uses)
// One draw of the reference's LambdaMartSelective::sampling_query_level (lambdamartselective.cc:326-493) on given
// labels / scores, after srand(0): ids_out = the permuted sample-id list, returns the sample size.
namespace {
class SelectiveProbe : public LambdaMartSelective {
 public:
  using LambdaMartSelective::LambdaMartSelective;
  size_t draw(std::shared_ptr<data::Dataset> ds, double *scores, size_t *ids, size_t *npos, float adapt) {
    this->scores_on_training_ = scores;
    size_t n = this->sampling_query_level(ds, ids, npos, adapt);
    this->scores_on_training_ = NULL;
    return n;
  }
};
}  // namespace

uint64_t qref_selective_sample(double rank_factor, double random_factor, int adaptive, int negative, double adapt_factor,
                               const float *labels, const double *scores, const uint64_t *qoff, uint64_t N, uint64_t Q,
                               uint64_t *ids_out) {
  auto ds = std::make_shared<data::Dataset>(N, 1);
  for (uint64_t q = 0; q < Q; ++q)
    for (uint64_t i = qoff[q]; i < qoff[q + 1]; ++i) ds->addInstance((QueryID) (q + 1), labels[i], std::vector<Feature>(1, 0.0f));
  SelectiveProbe probe(1, 0.1, 0, 4, 1, 1.0f, 1.0f, 0, 0.0f, 1, (float) rank_factor, (float) random_factor, 100.0f,
                       kSelAdaptive[adaptive & 3], kSelNegative[negative % 3]);
  std::vector<size_t> ids(N), npos(Q, 0);
  for (uint64_t i = 0; i < N; ++i) ids[i] = i;
  for (uint64_t q = 0; q < Q; ++q)
    for (uint64_t i = qoff[q]; i < qoff[q + 1]; ++i) npos[q] += labels[i] > 0;
  std::vector<double> sc(scores, scores + N);
  Silence sil(true);
  srand(0);
  const size_t n = probe.draw(ds, sc.data(), ids.data(), npos.data(), (float) adapt_factor);
  for (uint64_t i = 0; i < N; ++i) ids_out[i] = ids[i];
  return n;
}

// The reference's own SVMLight reader (io::Svml::read_horizontal, src/io/svml.cc:38-161) on a file: shape, FNV-1a
// checksums of labels / query offsets / the row-major feature matrix (the quantities host/bin/svml_check prints for the
// host reader) and the wall time of the call.
static uint64_t fnv1a(const void *p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  const unsigned char *b = (const unsigned char *) p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
int qref_read_svml(const char *path, uint64_t shape[3], uint64_t sums[3], double *seconds) {
  io::Svml reader;
  const double t0 = omp_get_wtime();
  std::unique_ptr<data::Dataset> ds;
  {
    Silence sil(true);
    ds = reader.read_horizontal(path);
  }
  *seconds = omp_get_wtime() - t0;
  if (!ds) return 1;
  const size_t n = ds->num_instances(), f = ds->num_features(), q = ds->num_queries();
  shape[0] = n; shape[1] = f; shape[2] = q;
  std::vector<float> labels(n);
  std::vector<uint64_t> off(q + 1);
  for (size_t i = 0; i < n; ++i) labels[i] = ds->getLabel(i);
  for (size_t i = 0; i <= q; ++i) off[i] = ds->offset(i);
  sums[0] = fnv1a(labels.data(), n * sizeof(float));
  sums[1] = fnv1a(off.data(), off.size() * sizeof(uint64_t));
  sums[2] = fnv1a(ds->at(0, 0), n * f * sizeof(float));
  return 0;
}

void qref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int qref_max_threads(void) { return omp_get_max_threads(); }

// LineSearch::learn (src/learning/linear/line_search.cc:153-416) on a row-major matrix — the per-tree partial scores
// CLEAVER works on (driver.cc:411-446), or any feature matrix — with NDCG@cutoff as the metric and no validation
// set; returns the learned weights.  The reference's progress table goes to a discarded stream.
int qref_linesearch(const float *x, uint64_t N, uint64_t T, const float *labels, const uint64_t *qoff, uint64_t Q,
                    uint64_t cutoff, uint32_t num_points, double window_size, double reduction_factor,
                    uint32_t max_iterations, uint32_t max_failed_vali, int adaptive, uint32_t last_only,
                    const double *init_weights, double *out_weights) {
  auto ds = std::make_shared<data::Dataset>(N, T);
  std::vector<Feature> row(T);
  for (uint64_t q = 0; q < Q; ++q)
    for (uint64_t i = qoff[q]; i < qoff[q + 1]; ++i) {
      row.assign(x + i * T, x + (i + 1) * T);
      ds->addInstance((QueryID) (q + 1), labels[i], row);
    }
  auto metric = std::make_shared<metric::ir::Ndcg>(cutoff);
  learning::linear::LineSearch ls(num_points, window_size, reduction_factor, max_iterations, max_failed_vali,
                                  adaptive != 0, last_only);
  if (init_weights) {
    std::vector<double> w(init_weights, init_weights + T);
    ls.update_weights(w);
  }
  std::ostringstream sink;
  std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
  ls.learn(ds, nullptr, metric, 0, std::string());
  std::cout.rdbuf(old);
  const std::vector<double> w = ls.get_weights();
  if (w.size() != T) return 1;
  for (uint64_t f = 0; f < T; ++f) out_weights[f] = w[f];
  return 0;
}

// The same with a validation set (line_search.cc:360-383: the weights kept are those of the best validation
// iteration, `max_failed_vali` iterations without a new best stop the search).  The validation set must not hold more
// documents than the training set: the reference sizes its validation score buffer by the TRAINING set (:209-210).
int qref_linesearch_valid(const float *x, uint64_t N, uint64_t T, const float *labels, const uint64_t *qoff, uint64_t Q,
                          const float *xv, uint64_t Nv, const float *labelsv, const uint64_t *qoffv, uint64_t Qv,
                          uint64_t cutoff, uint32_t num_points, double window_size, double reduction_factor,
                          uint32_t max_iterations, uint32_t max_failed_vali, int adaptive, uint32_t last_only,
                          const double *init_weights, double *out_weights) {
  if (Nv > N) return 3;
  auto fill = [&](const float *m, uint64_t n, const float *lab, const uint64_t *off, uint64_t q) {
    auto ds = std::make_shared<data::Dataset>(n, T);
    std::vector<Feature> row(T);
    for (uint64_t k = 0; k < q; ++k)
      for (uint64_t i = off[k]; i < off[k + 1]; ++i) {
        row.assign(m + i * T, m + (i + 1) * T);
        ds->addInstance((QueryID) (k + 1), lab[i], row);
      }
    return ds;
  };
  auto ds = fill(x, N, labels, qoff, Q), dv = fill(xv, Nv, labelsv, qoffv, Qv);
  auto metric = std::make_shared<metric::ir::Ndcg>(cutoff);
  learning::linear::LineSearch ls(num_points, window_size, reduction_factor, max_iterations, max_failed_vali,
                                  adaptive != 0, last_only);
  if (init_weights) {
    std::vector<double> w(init_weights, init_weights + T);
    ls.update_weights(w);
  }
  {
    Silence sil(true);
    ls.learn(ds, dv, metric, 0, std::string());
  }
  const std::vector<double> w = ls.get_weights();
  if (w.size() != T) return 1;
  for (uint64_t f = 0; f < T; ++f) out_weights[f] = w[f];
  return 0;
}

// Cleaver::optimize (src/optimization/post_learning/cleaver/cleaver.cc:166-412) with one of the deterministic pruning
// strategies on a partial-score matrix, starting from `weights`; line search before / after pruning as the strategy
// asks (num_points == 0: no line search).  method: 0 LAST, 1 SKIP, 2 LOW_WEIGHTS, 3 QUALITY_LOSS, 4 QUALITY_LOSS_ADV,
// 5 SCORE_LOSS.
int qref_cleaver(int method, const float *x, uint64_t N, uint64_t T, const float *labels, const uint64_t *qoff,
                 uint64_t Q, uint64_t cutoff, double pruning_rate, const double *weights, uint32_t num_points,
                 double window_size, double reduction_factor, uint32_t max_iterations, double *out_weights) {
  namespace pr = optimization::post_learning::pruning;
  auto ds = std::make_shared<data::Dataset>(N, T);
  std::vector<Feature> row(T);
  for (uint64_t q = 0; q < Q; ++q)
    for (uint64_t i = qoff[q]; i < qoff[q + 1]; ++i) {
      row.assign(x + i * T, x + (i + 1) * T);
      ds->addInstance((QueryID) (q + 1), labels[i], row);
    }
  std::shared_ptr<metric::ir::Metric> metric = std::make_shared<metric::ir::Ndcg>(cutoff);
  std::shared_ptr<learning::linear::LineSearch> ls;
  if (num_points)
    ls = std::make_shared<learning::linear::LineSearch>(num_points, window_size, reduction_factor, max_iterations, 20u,
                                                        false, 0u);
  std::shared_ptr<pr::Cleaver> cl;
  switch (method) {
    case 0: cl = std::make_shared<pr::LastPruning>(pruning_rate, ls); break;
    case 1: cl = std::make_shared<pr::SkipPruning>(pruning_rate, ls); break;
    case 2: cl = std::make_shared<pr::LowWeightsPruning>(pruning_rate, ls); break;
    case 3: cl = std::make_shared<pr::QualityLossPruning>(pruning_rate, ls); break;
    case 4: cl = std::make_shared<pr::QualityLossAdvPruning>(pruning_rate, ls); break;
    case 5: cl = std::make_shared<pr::ScoreLossPruning>(pruning_rate, ls); break;
    default: return 2;
  }
  std::vector<double> w(weights, weights + T);
  cl->update_weights(w);
  cl->set_update_model(false);
  std::ostringstream sink;
  std::streambuf *old = std::cout.rdbuf(sink.rdbuf());
  cl->optimize(nullptr, ds, nullptr, metric, 0, std::string());
  std::cout.rdbuf(old);
  const std::vector<double> r = cl->get_weigths();
  if (r.size() != T) return 1;
  for (uint64_t f = 0; f < T; ++f) out_weights[f] = r[f];
  return 0;
}

}  // extern "C"
