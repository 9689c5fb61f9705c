// quickrank_b200 — the subclass INTEGRATION.md describes, compiled against the UNMODIFIED reference headers
// (/root/reference/include) to prove the drop-in boundary: stock Mart::learn (mart.cc:208-416) runs unchanged
// and every hook (include/learning/forests/mart.h:118-147) is a call into the C ABI of include/quickrank_b200.h.
// Test infrastructure (built by oracle/Makefile target `integration` into oracle/_ref/), not product code.
#pragma once

#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include "data/dataset.h"
#include "data/vertical_dataset.h"
#include "learning/forests/lambdamart.h"
#include "learning/tree/rt.h"
#include "metric/ir/ndcg.h"
#include "quickrank_b200.h"

namespace quickrank {
namespace learning {
namespace forests {

// a RegressionTree whose root is built from the flat arrays qr_fit_tree returns; ownership of the root
// passes to Ensemble::push (mart.cc:342) exactly as for the reference's own trees
class GpuRegressionTree : public RegressionTree {
 public:
  explicit GpuRegressionTree(const qr_flat_tree &t) : RegressionTree(0, NULL, NULL, 1, 0.0f) { root = build(t, 0); }

 private:
  static RTNode *build(const qr_flat_tree &t, uint32_t i) {
    if (t.feature[i] < 0) return new RTNode(t.value[i]);                                       // rtnode.h:65
    RTNode *l = build(t, (uint32_t) t.left[i]), *r = build(t, (uint32_t) t.right[i]);
    return new RTNode(t.threshold[i], (size_t) t.feature[i], (size_t) t.feature[i] + 1, l, r);   // rtnode.h:85, rt.cc:317-318
  }
};

class GpuLambdaMart : public LambdaMart {
 public:
  using LambdaMart::LambdaMart;
  virtual ~GpuLambdaMart() {}
  qr_ctx *context() const { return ctx_; }
  void set_hist_mode(uint32_t m) { hist_mode_ = m; }
  void set_cutoff(size_t k) { cutoff_ = k; }

 protected:
  void init(std::shared_ptr<data::VerticalDataset> d) override {   // mart.cc:117-176, lambdamart.cc:35-39
    const size_t n = d->num_instances();
    // host arrays the base class' learn() dereferences unconditionally (mart.cc:121-122, 335, 345)
    scores_on_training_ = new double[n]();
    pseudoresponses_ = new double[n]();
    // Mart::learn calls the NON-virtual hist_->update() right before fit_regressor_on_gradient (mart.cc:335):
    // a root histogram over zero features makes that call (almost) a no-op; the real refresh is in qr_fit_tree
    empty_.reset(new data::VerticalDataset(std::shared_ptr<data::Dataset>(new data::Dataset(0, 0))));
    hist_ = new RTRootHistogram(empty_.get(), NULL, 0, NULL, NULL);
    qr_params p;
    std::memset(&p, 0, sizeof(p));
    p.algo = QR_ALGO_LAMBDAMART;
    p.nleaves = (uint32_t) nleaves_;
    p.minleafsupport = (uint32_t) minleafsupport_;
    p.nthresholds = nthresholds_;
    p.ndcg_cutoff = cutoff_;
    p.shrinkage = shrinkage_;
    p.hist_mode = hist_mode_;
    p.device = -1;
    std::vector<uint64_t> off(d->num_queries() + 1);
    for (size_t q = 0; q <= d->num_queries(); ++q) off[q] = d->offset(q);
    std::vector<float> labels(n);
    for (size_t i = 0; i < n; ++i) labels[i] = d->getLabel(i);
    // VerticalDataset::data_ is column-major: at(0, f) is the start of feature f, at(0, 0) the whole matrix
    if (qr_ctx_create(d->at(0, 0), n, d->num_features(), labels.data(), off.data(), d->num_queries(), &p, &ctx_) != QR_OK)
      fail();
  }
  void clear(size_t) override {   // mart.cc:178-206
    qr_ctx_destroy(ctx_);
    ctx_ = NULL;
    delete hist_; hist_ = NULL;
    delete[] scores_on_training_; scores_on_training_ = NULL;
    delete[] pseudoresponses_; pseudoresponses_ = NULL;
  }
  void compute_pseudoresponses(std::shared_ptr<data::VerticalDataset>, metric::ir::Metric *, bool *sample_presence) override {
    if (sample_presence) fail("document sub-sampling is not supported on the GPU");
    if (qr_compute_pseudoresponses(ctx_) != QR_OK) fail();   // lambdamart.cc:62-152
  }
  std::unique_ptr<RegressionTree> fit_regressor_on_gradient(std::shared_ptr<data::VerticalDataset>, size_t *) override {
    const uint32_t cap = 2 * (uint32_t) nleaves_ + 1;
    std::vector<int32_t> feature(cap), left(cap), right(cap);
    std::vector<uint32_t> tidx(cap);
    std::vector<float> thr(cap);
    std::vector<double> value(cap);
    qr_flat_tree t;
    std::memset(&t, 0, sizeof(t));
    t.capacity = cap;
    t.feature = feature.data(); t.threshold_idx = tidx.data(); t.threshold = thr.data();
    t.left = left.data(); t.right = right.data(); t.value = value.data();
    if (qr_fit_tree(ctx_, &t) != QR_OK) fail();              // mart.cc:335-339, lambdamart.cc:47-60
    return std::unique_ptr<RegressionTree>(new GpuRegressionTree(t));
  }
  void update_modelscores(std::shared_ptr<data::VerticalDataset>, Score *, RegressionTree *) override {
    if (qr_update_modelscores(ctx_, shrinkage_) != QR_OK) fail();   // mart.cc:459-468
  }
  void fail(const char *m = NULL) const {   // the reference's error convention (ensemble.cc:98-101)
    std::cerr << "!!! " << (m ? m : qr_last_error()) << std::endl;
    exit(EXIT_FAILURE);
  }

  qr_ctx *ctx_ = NULL;
  uint32_t hist_mode_ = QR_HIST_FAST;
  size_t cutoff_ = 10;
  std::unique_ptr<data::VerticalDataset> empty_;
};

}  // namespace forests
}  // namespace learning

namespace metric {
namespace ir {
// Metric::evaluate_dataset is virtual (metric.h:77,93): the training-set evaluation of Mart::learn (mart.cc:347)
// is routed to the device; the scores live there (the host array handed in is the base class' zero buffer)
class GpuNdcg : public Ndcg {
 public:
  GpuNdcg(size_t k, learning::forests::GpuLambdaMart *algo) : Ndcg(k), algo_(algo) {}
  MetricScore evaluate_dataset(const std::shared_ptr<data::VerticalDataset>, const Score *) const override {
    double m = 0;
    if (qr_evaluate(algo_->context(), &m) != QR_OK) { std::cerr << "!!! " << qr_last_error() << std::endl; exit(EXIT_FAILURE); }
    return m;
  }
  using Ndcg::evaluate_dataset;

 private:
  learning::forests::GpuLambdaMart *algo_;
};
}  // namespace ir
}  // namespace metric
}  // namespace quickrank
