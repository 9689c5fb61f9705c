// quickrank_b200 host layer — LambdaMART trainers that fit every tree on a per-query document sample
// (SURVEY.md section 8f-3): LAMBDAMART-SELECTIVE (lambdamartselective.cc) and STOCHASTIC-NEGATIVE
// (stochasticnegative.cc) of the reference.
//
// Division of labour: which documents are in the sample is decided here, on the host, from the current
// scores (per-query sorts, quotas, shuffles: a few passes over N ids every `sampling-iterations` trees);
// everything that touches features or gradients runs on the GPU — the sample becomes a training context
// of its own (qr_ctx_create_sample) binned with the thresholds of the whole training set, each iteration
// pulls the current scores into it (qr_sample_pull_scores), computes the pseudo-responses of the sampled
// documents, grows the tree and fits the leaf outputs on them; the tree is then applied to ALL documents
// in the full context and NDCG is evaluated there (lambdamartselective.cc:196-215).
//
// Reference behaviour reproduced on purpose (the parity tests in tests/test_sampled_trainers.py compare
// whole training runs with the unmodified reference):
//  * once sampling is enabled a query is RANKED, for the lambdas, by scores_on_training_[d] with d the
//    document's position inside its query (lambdamart.cc:94 drops the query offset), from the first
//    iteration on; rho still uses the documents' own scores (lambdamart.cc:132-134);
//  * the per-query orderings come from libstdc++'s std::sort under the reference's comparators and the
//    random negatives from std::random_shuffle's rand() stream after srand(0): this file calls the same
//    std::sort with equivalent comparators and restates random_shuffle's loop;
//  * the quotas are computed in float, as the reference's `float * size_t` products are.
//
// Provenance: like host/src/quickrank_host.cc this is reference-facing HOST code — the training loop, its stdout table
// and the sampling rules restate hpclab/quickrank's src/learning/forests/{lambdamartselective,stochasticnegative}.cc
// (Reciprocal Public License 1.5) closely enough for the printed lines, the rand() stream and the drawn samples to be
// identical, and it is to be read under the same licence.  The structure (one shared loop with hooks, a parallel
// ordering phase and a sequential selection phase in the draw) and everything below the C ABI are this repository's.
#include "quickrank_host.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <limits>
#include <numeric>
#include <atomic>
#include <random>
#include <thread>

namespace quickrank {
namespace learning {
namespace forests {

const std::string LambdaMartSelective::NAME_ = "LAMBDAMART-SELECTIVE";
const std::string StochasticNegative::NAME_ = "STOCHASTIC-NEGATIVE";

SampledLambdaMart::~SampledLambdaMart() {
  if (sample_ctx_) qr_ctx_destroy(sample_ctx_);
}

void SampledLambdaMart::clear(size_t num_features) {
  if (sample_ctx_) qr_ctx_destroy(sample_ctx_);
  sample_ctx_ = nullptr;
  LambdaMart::clear(num_features);
}

// The sample as a device context: rows in ascending document order inside each query (the order in which
// compute_pseudoresponses compacts a query, lambdamart.cc:90-98), queries without a sampled document dropped.
void SampledLambdaMart::build_sample_context(const data::Dataset &dataset, const std::vector<size_t> &ids, size_t n) {
  const size_t N = dataset.num_instances(), F = dataset.num_features();
  std::vector<char> present(N, 0);
  for (size_t i = 0; i < n; ++i) present[ids[i]] = 1;
  std::vector<float> labels;
  labels.reserve(n);
  std::vector<uint32_t> src, key;
  src.reserve(n);
  key.reserve(n);
  std::vector<uint64_t> qoff(1, 0);
  for (size_t q = 0; q < dataset.num_queries(); ++q) {
    const size_t begin = dataset.offset(q), end = dataset.offset(q + 1);
    for (size_t d = begin; d < end; ++d) {
      if (!present[d]) continue;
      labels.push_back(dataset.getLabel(d));
      src.push_back((uint32_t) d);
      key.push_back((uint32_t) (d - begin));   // lambdamart.cc:94
    }
    if (src.size() > qoff.back()) qoff.push_back(src.size());
  }
  if (src.empty()) {
    std::cerr << "!!! The document sample is empty." << std::endl;
    exit(EXIT_FAILURE);
  }
  // no feature rows: the sample's bins are gathered on the device from the full context's; the first call creates a
  // context sized for the whole training set, later draws refill it in place
  if (sample_ctx_ == nullptr) {
    if (qr_ctx_create_sample(ctx_, nullptr, src.size(), F, labels.data(), qoff.data(), qoff.size() - 1, src.data(),
                             key.data(), &sample_ctx_) != QR_OK)
      die("Impossible to initialise the GPU context of the document sample");
  } else if (qr_sample_redraw(sample_ctx_, ctx_, src.size(), labels.data(), qoff.data(), qoff.size() - 1, src.data(),
                              key.data()) != QR_OK) {
    die("Impossible to load the new document sample");
  }
}

void SampledLambdaMart::learn(std::shared_ptr<data::Dataset> training_dataset,
                              std::shared_ptr<data::Dataset> validation_dataset,
                              std::shared_ptr<metric::ir::Metric> scorer, size_t partial_save,
                              const std::string output_basename) {
  if (scorer->name() != "NDCG") {
    std::cerr << "!!! The GPU engine optimises NDCG only (got " << scorer->name() << ")." << std::endl;
    exit(EXIT_FAILURE);
  }
  if (max_features_ != 1.0f || collapse_leaves_factor_ != 0.0f) {
    std::cerr << "!!! max_features and collapse_leaves_factor are not supported by the GPU engine." << std::endl;
    exit(EXIT_FAILURE);
  }
  if (host::sharding().world > 1) {
    std::cerr << "!!! " << name() << " trains on one GPU (the document sample is not sharded)." << std::endl;
    exit(EXIT_FAILURE);
  }
  check_supported();
  const bool sampling = sampling_enabled();
  const size_t N = training_dataset->num_instances();

  std::cout << "# Initialization";
  std::cout.flush();
  const auto init_start = std::chrono::high_resolution_clock::now();
  metric_cutoff_ = scorer->cutoff();
  std::shared_ptr<data::VerticalDataset> vertical_training(new data::VerticalDataset(training_dataset));
  best_metric_on_validation_ = best_metric_on_training_ = std::numeric_limits<double>::lowest();
  best_model_ = 0;
  ensemble_model_.set_capacity(ntrees_);
  init(vertical_training);
  if (validation_dataset &&
      qr_ctx_create_eval(ctx_, validation_dataset->data(), validation_dataset->num_instances(),
                         validation_dataset->num_features(), validation_dataset->labels(),
                         validation_dataset->offsets().data(), validation_dataset->num_queries(), &valid_ctx_) != QR_OK)
    die("Impossible to initialise the GPU validation context");
  if (ensemble_model_.is_notempty()) {   // restart from a loaded model
    best_model_ = ensemble_model_.get_size() - 1;
    std::vector<Score> s(N);
    score_dataset(training_dataset, s.data());
    if (qr_set_scores(ctx_, s.data()) != QR_OK) die("restart");
    best_metric_on_training_ = evaluate_training(scorer.get());
    if (validation_dataset) {
      std::vector<Score> v(validation_dataset->num_instances());
      score_dataset(validation_dataset, v.data());
      if (qr_set_scores(valid_ctx_, v.data()) != QR_OK) die("restart");
      if (qr_evaluate(valid_ctx_, &best_metric_on_validation_) != QR_OK) die("restart");
    }
  }

  // the sample: ids[0 .. nsample) of a permutation of the documents; until the first draw it is everything
  std::vector<size_t> ids(N), npositives;
  std::iota(ids.begin(), ids.end(), (size_t) 0);
  size_t nsample = N;
  if (sampling) {
    npositives.assign(training_dataset->num_queries(), 0);
    for (size_t q = 0; q < training_dataset->num_queries(); ++q)
      for (size_t d = training_dataset->offset(q); d < training_dataset->offset(q + 1); ++d)
        npositives[q] += training_dataset->getLabel(d) > 0;
    build_sample_context(*training_dataset, ids, nsample);
  }
  std::cout << ": " << std::setprecision(2)
            << std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - init_start).count() << " s."
            << std::endl;

  std::cout << std::fixed << std::setprecision(4);
  std::cout << "# Training:" << std::endl;
  std::cout << "# -------------------------" << std::endl;
  std::cout << "# iter. training validation" << std::endl;
  std::cout << "# -------------------------" << std::endl;
  if (ensemble_model_.is_notempty()) {
    std::cout << std::setw(7) << ensemble_model_.get_size() << std::setw(9) << best_metric_on_training_;
    if (validation_dataset) std::cout << std::setw(9) << best_metric_on_validation_;
    std::cout << " *" << std::endl;
  }
  const auto train_start = std::chrono::high_resolution_clock::now();
  before_training();

  std::vector<Score> host_scores;
  for (size_t m = ensemble_model_.get_size(); m < ntrees_; ++m) {
    if (validation_dataset && valid_iterations_ && m > best_model_ + valid_iterations_) break;
    if (sampling && resample_due(m)) {
      host_scores.resize(N);
      if (qr_get_scores(ctx_, host_scores.data()) != QR_OK) die("document sampling (scores)");
      std::iota(ids.begin(), ids.end(), (size_t) 0);
      nsample = draw_sample(*training_dataset, host_scores, npositives, ids);
      std::cout << "Reducing training size from " << N << " to " << nsample << std::endl;
      build_sample_context(*training_dataset, ids, nsample);
    }
    std::unique_ptr<RegressionTree> tree;
    if (sampling) {
      if (qr_sample_pull_scores(sample_ctx_, ctx_) != QR_OK) die("document sampling (scores of the sample)");
      if (qr_compute_pseudoresponses(sample_ctx_) != QR_OK) die("compute_pseudoresponses (document sample)");
      tree = fit_tree_on(sample_ctx_);
      ensemble_model_.push(tree->get_proot(), shrinkage_, 0);
      apply_tree_on(ctx_, tree.get());   // update_modelscores runs over all training documents
    } else {
      compute_pseudoresponses(vertical_training, scorer.get(), nullptr);
      tree = fit_regressor_on_gradient(vertical_training, nullptr);
      ensemble_model_.push(tree->get_proot(), shrinkage_, 0);
      update_modelscores(vertical_training, nullptr, tree.get());
    }
    const MetricScore metric_on_training = evaluate_training(scorer.get());
    std::cout << std::setw(7) << m + 1 << std::setw(9) << metric_on_training;
    bool is_best = false;
    if (validation_dataset) {
      update_modelscores(validation_dataset, nullptr, tree.get());
      MetricScore metric_on_validation = 0;
      if (qr_evaluate(valid_ctx_, &metric_on_validation) != QR_OK) die("evaluate_dataset (validation)");
      std::cout << std::setw(9) << metric_on_validation;
      if (metric_on_validation > best_metric_on_validation_) {
        best_metric_on_training_ = metric_on_training;
        best_metric_on_validation_ = metric_on_validation;
        is_best = true;
      }
    } else if (metric_on_training > best_metric_on_training_) {
      best_metric_on_training_ = metric_on_training;
      is_best = true;
    }
    if (is_best) {
      best_model_ = ensemble_model_.get_size() - 1;
      std::cout << " *";
    }
    std::cout << std::endl;
    // (the reference looks at best_model_, so a tie with an older best counts as "no improvement")
    after_iteration(m, best_model_ == ensemble_model_.get_size() - 1);
    if (partial_save != 0 && !output_basename.empty() && (m + 1) % partial_save == 0) save(output_basename, (int) (m + 1));
  }
  if (validation_dataset)
    while (ensemble_model_.is_notempty() && ensemble_model_.get_size() > best_model_ + 1) ensemble_model_.pop();
  const double train_time =
      std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - train_start).count();
  std::cout << std::endl;
  std::cout << *scorer << " on training data = " << best_metric_on_training_ << std::endl;
  if (validation_dataset) std::cout << *scorer << " on validation data = " << best_metric_on_validation_ << std::endl;
  clear(vertical_training->num_features());
  std::cout << std::endl;
  std::cout << "#\t Training Time: " << std::setprecision(2) << train_time << " s." << std::endl;
}

// ------------------------------------------------------------------------------------------------
// LAMBDAMART-SELECTIVE
// ------------------------------------------------------------------------------------------------

void LambdaMartSelective::check_supported() const {
  if (subsample_ != 1.0f) {
    std::cerr << "!!! " << name() << ": subsample is seeded from the wall clock in the reference and is not supported."
              << std::endl;
    exit(EXIT_FAILURE);
  }
  if (sampling_enabled() && sampling_iterations <= 0) {
    // lambdamartselective.cc:170-171 evaluates m % sampling_iterations: the reference dies on a division by zero
    std::cerr << "!!! " << name() << ": --sampling-iterations must be positive when a sampling factor is set." << std::endl;
    exit(EXIT_FAILURE);
  }
  static const char *adaptive[] = {"NO", "FIXED", "RATIO", "MIX"};
  static const char *negative[] = {"RATIO", "MUL", "POS"};
  if (std::find(std::begin(adaptive), std::end(adaptive), adaptive_strategy) == std::end(adaptive) ||
      std::find(std::begin(negative), std::end(negative), negative_strategy) == std::end(negative)) {
    std::cerr << "!!! " << name() << ": unknown adaptive strategy (NO, FIXED, RATIO, MIX) or negative strategy (RATIO, MUL, POS)."
              << std::endl;
    exit(EXIT_FAILURE);
  }
}

void LambdaMartSelective::before_training() {
  srand(0);   // lambdamartselective.cc:160
  improvements_.assign((size_t) std::max(0, (int) normalization_factor), true);
  adapt_factor_ = 1;
}

// lambdamartselective.cc:246-256: the share of the last `normalization_factor` iterations that set a new best
void LambdaMartSelective::after_iteration(size_t m, bool is_best) {
  if (adaptive_strategy == "NO" || !(normalization_factor > 0) || improvements_.empty()) return;
  improvements_[m % improvements_.size()] = is_best;
  const float hits = (float) std::accumulate(improvements_.begin(), improvements_.end(), 0.0);
  adapt_factor_ = hits / improvements_.size();
}

std::ostream &LambdaMartSelective::put(std::ostream &os) const {
  Mart::put(os);
  os << "# sampling iterations = " << sampling_iterations << std::endl;
  os << "# rank sampling factor = " << rank_sampling_factor << std::endl;
  os << "# random sampling factor = " << random_sampling_factor << std::endl;
  os << "# normalization factor = " << normalization_factor << std::endl;
  os << "# adaptive strategy = " << adaptive_strategy << std::endl;
  os << "# negative strategy = " << negative_strategy << std::endl;
  return os;
}

namespace {

// std::random_shuffle(first, last) of libstdc++ (bits/stl_algo.h): element i trades places with one of
// 0..i drawn with rand() % (i + 1)
size_t g_rand_calls = 0;   // rand() calls made by the draws of this process (host/selective_check.cc reports it)
template <typename T>
void rand_shuffle(std::vector<T> &v) {
  g_rand_calls += v.size() > 1 ? v.size() - 1 : 0;
  for (size_t i = 1; i < v.size(); ++i) {
    const size_t j = (size_t) (std::rand() % (long) (i + 1));
    if (i != j) std::swap(v[i], v[j]);
  }
}

// how many of a query's negatives a factor selects (the reference multiplies a float by a size_t: float product)
inline size_t quota(float factor, size_t count) { return (size_t) std::round(factor * count); }

}  // namespace

size_t LambdaMartSelective::rand_calls() { return g_rand_calls; }

// One draw (lambdamartselective.cc:326-493).  Per query: the positives, the `n_top` negatives the model scores
// highest and `n_random` of the remaining ones; the selected ids of all queries end up contiguous at the front of
// `ids`, the rest behind them.
size_t LambdaMartSelective::sampling_query_level(const data::Dataset &dataset, const std::vector<Score> &scores,
                                                 const std::vector<size_t> &npositives, std::vector<size_t> &ids,
                                                 float adapt_factor) {
  if (!sampling_iterations) return dataset.num_instances();
  const Score *score = scores.data();
  const float lo = std::min(rank_sampling_factor, random_sampling_factor);
  const float hi = std::max(rank_sampling_factor, random_sampling_factor);
  const float total = rank_sampling_factor + random_sampling_factor;
  const float slack = 1 - adapt_factor;
  float rank_factor = rank_sampling_factor, random_factor = random_sampling_factor;
  if (adaptive_strategy == "FIXED") {
    // (fmin / fmax / the sum below are the C double functions in the reference: the blend is a double expression)
    rank_factor = random_factor = (float) ((double) lo + slack * ((double) hi - (double) lo));
  } else if (adaptive_strategy == "RATIO") {
    rank_factor = total * adapt_factor;
    random_factor = total - rank_factor;
  } else if (adaptive_strategy == "MIX") {
    const double blend = (double) lo + slack * ((double) hi - (double) lo);
    rank_factor = (float) (blend * adapt_factor);
    random_factor = (float) (blend - rank_factor);
  }
  std::cout << "Rank Factor: " << rank_factor << " - Random Factor: " << random_factor
            << " - Adapt Factor: " << adapt_factor << std::setprecision(4) << std::endl;

  const bool by_ratio = negative_strategy == "RATIO", by_mul = negative_strategy == "MUL";
  const size_t Q = dataset.num_queries();
  // Phase 1, queries in parallel: each query's quotas and its final order.  A query's sorts read and write its own
  // segment of `ids` only, and phase 2 of an earlier query never writes behind that query's end, so doing all the
  // sorts first leaves every segment exactly as the reference's interleaved loop finds it.
  struct Quota { size_t n_top, n_random; };
  std::vector<Quota> quotas(Q);
  std::atomic<size_t> bad_query(Q);
  auto order_queries = [&](size_t q0, size_t q1) {
    for (size_t q = q0; q < q1; ++q) {
      const size_t begin = dataset.offset(q), end = dataset.offset(q + 1), len = end - begin;
      const size_t npos = npositives[q], nneg = len - npos;
      size_t n_top = 0, n_random = 0;
      if (by_ratio) {
        n_top = quota(rank_factor, nneg);
        n_random = quota(random_factor, nneg);
      } else if (by_mul) {
        n_top = std::min(quota(rank_factor, npos), nneg);
        n_random = std::min(quota(random_factor, npos), nneg);
      } else if (npos > 0) {   // POS: quotas relative to the negatives ranked above the last positive
        std::sort(ids.begin() + begin, ids.begin() + end, [score](size_t a, size_t b) { return score[a] > score[b]; });
        size_t last_positive = 0;
        for (size_t i = 0; i < len; ++i)
          if (dataset.getLabel(ids[begin + i]) > 0) last_positive = i;
        const size_t above = last_positive - npos + 1;
        n_top = std::min(quota(rank_factor, above), nneg);
        n_random = std::min(quota(random_factor, above), nneg - n_top);
      }
      if (n_top > nneg) {   // (the reference's unsigned `n_neg_query - n_top_neg` wraps around here and it dies in a vector constructor)
        size_t seen = bad_query.load();
        while (q < seen && !bad_query.compare_exchange_weak(seen, q)) {}
        n_top = nneg;
      }
      if (n_top + n_random > nneg) n_random = nneg - n_top;
      quotas[q] = Quota{n_top, n_random};
      // positives first, then negatives, each group by decreasing score (the reference's comparator, verbatim in
      // meaning: a positive precedes a zero label or a lower score; a non-positive precedes only a zero label of
      // lower score)
      std::sort(ids.begin() + begin, ids.begin() + end, [score, &dataset](size_t a, size_t b) {
        const bool higher = score[a] > score[b];
        const bool b_zero = dataset.getLabel(b) == 0;
        return dataset.getLabel(a) > 0 ? (b_zero || higher) : (b_zero && higher);
      });
    }
  };
  const size_t nthreads = std::max<size_t>(1, std::min<size_t>({(size_t) std::thread::hardware_concurrency(), (size_t) 32,
                                                               dataset.num_instances() / 20000 + 1}));
  if (nthreads == 1) {
    order_queries(0, Q);
  } else {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nthreads; ++t) pool.emplace_back(order_queries, Q * t / nthreads, Q * (t + 1) / nthreads);
    for (auto &t : pool) t.join();
  }
  if (bad_query.load() < Q) {
    std::cerr << "!!! " << name() << ": the rank sampling factor selects more negatives than query " << bad_query.load()
              << " has." << std::endl;
    exit(EXIT_FAILURE);
  }
  // Phase 2, in query order: the selected documents move to the front, the random negatives follow rand()'s stream
  size_t front = 0, picked_top = 0, picked_random = 0, positives = 0;
  for (size_t q = 0; q < Q; ++q) {
    const size_t begin = dataset.offset(q), len = dataset.offset(q + 1) - begin;
    const size_t npos = npositives[q], n_top = quotas[q].n_top, n_random = quotas[q].n_random;
    picked_top += n_top;
    picked_random += n_random;
    positives += npos;
    const size_t head = npos + n_top;
    if (front > 0)
      for (size_t j = 0; j < head; ++j) std::swap(ids[front + j], ids[begin + j]);
    if (n_random > 0) {
      std::vector<int> rest(len - head);
      std::iota(rest.begin(), rest.end(), (int) head);
      rand_shuffle(rest);
      for (size_t j = 0; j < n_random; ++j) std::swap(ids[front + head + j], ids[begin + rest[j]]);
    }
    front += head + n_random;
  }
  std::cout << std::setprecision(0) << "N. Positives: " << positives << " - Neg sel rank: " << picked_top
            << " - Neg sel random: " << picked_random << std::setprecision(4) << std::endl;
  return front;
}

// ------------------------------------------------------------------------------------------------
// STOCHASTIC-NEGATIVE
// ------------------------------------------------------------------------------------------------

void StochasticNegative::check_supported() const {
  if (!(subsample_ > 0.0f)) {
    std::cerr << "!!! " << name() << ": subsample must be positive." << std::endl;
    exit(EXIT_FAILURE);
  }
}

// stochasticnegative.cc:285-333: per query, the documents by decreasing label (std::sort), the non-positive ones
// shuffled, the positives and the first `share` negatives kept
size_t StochasticNegative::draw_sample(const data::Dataset &dataset, const std::vector<Score> &,
                                       const std::vector<size_t> &npositives, std::vector<size_t> &ids) {
  if (subsample_ == 1.0f) return dataset.num_instances();
  std::default_random_engine rng((std::default_random_engine::result_type) (seed_ + draws_++));
  size_t front = 0;
  for (size_t q = 0; q < dataset.num_queries(); ++q) {
    const size_t begin = dataset.offset(q), end = dataset.offset(q + 1);
    const size_t npos = npositives[q], nneg = end - begin - npos;
    const size_t share = subsample_ > 1.0f ? std::min((size_t) subsample_, nneg) : (size_t) std::floor(subsample_ * nneg);
    std::sort(ids.begin() + begin, ids.begin() + end,
              [&dataset](size_t a, size_t b) { return dataset.getLabel(a) > dataset.getLabel(b); });
    std::shuffle(ids.begin() + begin + npos, ids.begin() + end, rng);
    if (front > 0)
      for (size_t j = 0; j < npos + share; ++j) std::swap(ids[front + j], ids[begin + j]);
    front += npos + share;
  }
  return front;
}

}  // namespace forests
}  // namespace learning
}  // namespace quickrank
