// quickrank_b200 host layer — implementation.  See host/include/quickrank_host.h.
//
// Provenance: this file is the reference-facing HOST layer (north_star keeps QuickRank's CLI, Dataset API, XML model
// format and training-loop control flow; the GPU engine sits below it behind include/quickrank_b200.h).  Its control
// flow — Mart::learn / Dart::learn, DART's tree selection and weight normalisation, the XML and stdout formats, the
// code generators — is a condensed restatement of hpclab/quickrank's src/learning/forests/{mart,dart}.cc,
// src/io/generate_*.cc and src/data/*.cc (Reciprocal Public License 1.5), kept statement-compatible on purpose: the
// std::rand() stream, the printed table and the saved model must equal the reference's for the parity tests to mean
// anything.  It is derived work of that code and is to be read under the same licence; no arithmetic on the hot path
// lives here (every pass over documents is a call into the C ABI), and it is not meant to grow.
#include "quickrank_host.h"
#include "qr_fast_float.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <map>
#include <numeric>
#include <sstream>
#include <thread>

#include <sys/stat.h>
#include <unistd.h>

namespace quickrank {

// ------------------------------------------------------------------------------------------------
// data
// ------------------------------------------------------------------------------------------------
namespace data {

void QueryResults::indexing_of_sorted_labels(const Score *scores, size_t *dest) const {
  for (size_t i = 0; i < num_results_; ++i) dest[i] = i;
  // libstdc++ std::sort with a strict "greater" comparator: the same call the reference makes
  // (queryresults.cc:47-53), hence the same permutation of tied scores
  std::sort(dest, dest + num_results_, [scores](int i, int j) { return scores[i] > scores[j]; });
}

void QueryResults::sorted_labels(const Score *scores, Label *dest, size_t cutoff) const {
  std::vector<size_t> idx(num_results_);
  indexing_of_sorted_labels(scores, idx.data());
  for (size_t i = 0; i < num_results_ && i < cutoff; ++i) dest[i] = labels_[idx[i]];
}

Dataset::Dataset(size_t n_instances, size_t n_features) : num_features_(n_features), max_instances_(n_instances) {
  void *p = nullptr, *q = nullptr;
  if (posix_memalign(&p, 64, std::max<size_t>(64, max_instances_ * num_features_ * sizeof(Feature))) != 0 ||
      posix_memalign(&q, 64, std::max<size_t>(64, max_instances_ * sizeof(Label))) != 0) {
    std::cerr << "!!! Impossible to allocate memory for dataset storage." << std::endl;
    exit(EXIT_FAILURE);
  }
  data_ = (Feature *) p;
  labels_ = (Label *) q;
  std::memset(data_, 0, max_instances_ * num_features_ * sizeof(Feature));
  offsets_.push_back(0);
}

Dataset::~Dataset() {
  free(data_);
  free(labels_);
}

void Dataset::addInstance(QueryID q_id, Label i_label, const std::vector<Feature> &i_features) {
  if (i_features.size() > num_features_ || num_instances_ == max_instances_) {
    std::cerr << "!!! Impossible to add a new instance to the dataset." << std::endl;
    exit(EXIT_FAILURE);
  }
  labels_[num_instances_] = i_label;
  std::copy(i_features.begin(), i_features.end(), data_ + num_instances_ * num_features_);
  if (num_instances_ == 0 || last_instance_id_ != q_id) {
    num_queries_++;
    offsets_.push_back(0);
    last_instance_id_ = q_id;
  }
  num_instances_++;
  offsets_.back() = num_instances_;
}

void Dataset::set_structure(const Label *labels, const std::vector<uint64_t> &offsets) {
  if (offsets.empty() || offsets.front() != 0 || offsets.back() != max_instances_) {
    std::cerr << "!!! Impossible to set the dataset structure." << std::endl;
    exit(EXIT_FAILURE);
  }
  std::copy(labels, labels + max_instances_, labels_);
  offsets_ = offsets;
  num_instances_ = max_instances_;
  num_queries_ = offsets.size() - 1;
}

std::unique_ptr<QueryResults> Dataset::getQueryResults(size_t i) const {
  const size_t n = offsets_[i + 1] - offsets_[i];
  return std::unique_ptr<QueryResults>(
      new QueryResults(n, labels_ + offsets_[i], data_ + offsets_[i] * num_features_));
}

VerticalDataset::VerticalDataset(std::shared_ptr<Dataset> h_dataset) : src_(h_dataset) {}

Feature *VerticalDataset::at(size_t document_id, size_t feature_id) {
  if (col_.empty()) {   // host transpose only on demand; training transposes on the device
    const size_t n = src_->num_instances(), f = src_->num_features();
    col_.resize(n * f);
    for (size_t i = 0; i < n; ++i)
      for (size_t j = 0; j < f; ++j) col_[j * n + i] = src_->data()[i * f + j];
  }
  return col_.data() + feature_id * src_->num_instances() + document_id;
}

std::unique_ptr<QueryResults> VerticalDataset::getQueryResults(size_t i) {
  const size_t d0 = src_->offset(i), n = src_->offset(i + 1) - d0;
  return std::unique_ptr<QueryResults>(new QueryResults(n, const_cast<Label *>(src_->labels()) + d0, at(d0, 0)));
}

}  // namespace data

// ------------------------------------------------------------------------------------------------
// metric (host evaluation of arbitrary score vectors: used outside the training loop)
// ------------------------------------------------------------------------------------------------
namespace metric {
namespace ir {

MetricScore Metric::evaluate_dataset(const std::shared_ptr<data::Dataset> dataset, const Score *scores) const {
  if (dataset->num_queries() == 0) return 0.0;
  MetricScore avg = 0.0;
  for (size_t q = 0; q < dataset->num_queries(); q++) {
    std::unique_ptr<data::QueryResults> r = dataset->getQueryResults(q);
    avg += evaluate_result_list(r.get(), scores);
    scores += r->num_results();
  }
  return avg / (MetricScore) dataset->num_queries();
}

MetricScore Dcg::compute_dcg(const Label *labels, size_t len) const {
  const size_t size = std::min(cutoff(), len);
  double dcg = 0.0;
  for (size_t i = 0; i < size; ++i) dcg += (std::pow(2.0, (double) labels[i]) - 1.0) / std::log2((double) ((float) i + 2.0f));
  return dcg;
}

MetricScore Dcg::evaluate_result_list(const data::QueryResults *rl, const Score *scores) const {
  const size_t size = std::min(cutoff(), rl->num_results());
  if (size == 0) return 0.0;
  std::vector<Label> sorted_l(size);
  rl->sorted_labels(scores, sorted_l.data(), cutoff());
  return compute_dcg(sorted_l.data(), size);
}

std::ostream &Dcg::put(std::ostream &os) const {
  if (cutoff() != Metric::NO_CUTOFF) return os << name() << "@" << cutoff();
  return os << name();
}

MetricScore Ndcg::compute_idcg(const data::QueryResults *rl) const {
  std::vector<Label> copy(rl->labels(), rl->labels() + rl->num_results());
  std::sort(copy.begin(), copy.end(), std::greater<int>());   // ndcg.cc:40-41 (int comparison)
  return compute_dcg(copy.data(), copy.size());
}

MetricScore Ndcg::evaluate_result_list(const data::QueryResults *rl, const Score *scores) const {
  if (rl->num_results() == 0) return 0.0;
  const MetricScore idcg = compute_idcg(rl);
  return idcg > 0 ? Dcg::evaluate_result_list(rl, scores) / idcg : 0.0;
}

std::ostream &Ndcg::put(std::ostream &os) const {
  if (cutoff() != Metric::NO_CUTOFF) return os << name() << "@" << cutoff();
  return os << name();
}

}  // namespace ir
}  // namespace metric

// ------------------------------------------------------------------------------------------------
// SVMLight I/O
// ------------------------------------------------------------------------------------------------
namespace io {

// Svml::read_horizontal (svml.cc:38-161).  Same grammar and error exits (2: malformed label / qid,
// 4: malformed feature), but the file is read once into memory and its lines are parsed by all host
// threads (QR_SVML_THREADS, default: hardware concurrency): the reference's getline + sscanf loop takes
// minutes on a config-2-sized text file, longer than training on the GPU does.
namespace {
struct SvmlChunk {
  std::vector<QueryID> qids;
  std::vector<Label> labels;
  std::vector<size_t> row_start;                          // per row, index into fids/vals; one extra entry at the end
  std::vector<uint32_t> fids;
  std::vector<Feature> vals;
  size_t maxfid = 0;
  int error = 0;
};

// parses the lines of [p, end); `end` is a line boundary (or the end of the buffer, which is NUL-terminated)
void parse_svml_range(char *p, char *end, SvmlChunk &out) {
  while (p < end) {
    char *eol = (char *) memchr(p, '\n', (size_t) (end - p));
    if (!eol) eol = end;
    char *q = p;
    p = eol < end ? eol + 1 : end;
    *eol = '\0';
    while (*q == ' ' || *q == '\t') ++q;
    if (*q == '#' || *q == '\0' || *q == '\r') continue;
    char *hash = strchr(q, '#');
    if (hash) *hash = '\0';
    char *e = nullptr;
    const Label rel = (Label) strtod(q, &e);
    if (e == q) { out.error = 2; return; }
    q = e;
    while (*q == ' ' || *q == '\t') ++q;
    if (strncmp(q, "qid:", 4) != 0) { out.error = 2; return; }
    const QueryID qid = (QueryID) strtoull(q + 4, &e, 10);
    q = e;
    out.row_start.push_back(out.fids.size());
    for (;;) {
      while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
      if (*q == '\0') break;
      size_t fid = 0;
      for (e = q; *e >= '0' && *e <= '9' && e - q < 18; ++e) fid = fid * 10 + (size_t) (*e - '0');
      if (e == q || *e != ':') { out.error = 4; return; }
      q = e + 1;
      const Feature v = host::parse_float(q, &e);   // = strtof(q, &e), bit for bit (host/include/qr_fast_float.h)
      if (e == q || fid == 0) { out.error = 4; return; }
      q = e;
      out.fids.push_back((uint32_t) fid);
      out.vals.push_back(v);
      out.maxfid = std::max(out.maxfid, fid);
    }
    out.qids.push_back(qid);
    out.labels.push_back(rel);
  }
  out.row_start.push_back(out.fids.size());
}
}  // namespace

// Binary cache of a parsed SVMLight file (QR_SVML_CACHE=1): header, labels, query offsets, row-major features.
namespace {
struct QrbHeader {
  char magic[8];            // "QRB1\0\0\0\0"
  uint64_t src_size;        // size and modification time of the text the cache was made from
  int64_t src_mtime_s, src_mtime_ns;
  uint64_t nrows, nfeatures, nqueries;
};

bool svml_cache_enabled() {
  const char *e = getenv("QR_SVML_CACHE");
  return e != nullptr && atoi(e) != 0;
}

bool stat_source(const std::string &filename, QrbHeader *h) {
  struct stat st;
  if (stat(filename.c_str(), &st) != 0) return false;
  h->src_size = (uint64_t) st.st_size;
  h->src_mtime_s = (int64_t) st.st_mtim.tv_sec;
  h->src_mtime_ns = (int64_t) st.st_mtim.tv_nsec;
  return true;
}

std::unique_ptr<data::Dataset> load_svml_cache(const std::string &filename) {
  QrbHeader want, got;
  std::memset(&want, 0, sizeof(want));
  if (!stat_source(filename, &want)) return nullptr;
  FILE *f = fopen((filename + ".qrb").c_str(), "rb");
  if (!f) return nullptr;
  std::unique_ptr<data::Dataset> ds;
  if (fread(&got, sizeof(got), 1, f) == 1 && std::memcmp(got.magic, "QRB1\0\0\0", 8) == 0 &&
      got.src_size == want.src_size && got.src_mtime_s == want.src_mtime_s && got.src_mtime_ns == want.src_mtime_ns &&
      got.nqueries <= got.nrows) {
    std::vector<Label> labels(got.nrows);
    std::vector<uint64_t> offsets(got.nqueries + 1);
    ds.reset(new data::Dataset(got.nrows, got.nfeatures));
    const size_t cells = (size_t) got.nrows * got.nfeatures;
    const bool ok = fread(labels.data(), sizeof(Label), labels.size(), f) == labels.size() &&
                    fread(offsets.data(), sizeof(uint64_t), offsets.size(), f) == offsets.size() &&
                    offsets.front() == 0 && offsets.back() == got.nrows &&
                    std::is_sorted(offsets.begin(), offsets.end()) &&
                    (cells == 0 || fread(ds->at(0, 0), sizeof(Feature), cells, f) == cells);
    if (ok) ds->set_structure(labels.data(), offsets);
    else ds.reset();
  }
  fclose(f);
  return ds;
}

void write_svml_cache(const std::string &filename, const data::Dataset &ds) {
  QrbHeader h;
  std::memset(&h, 0, sizeof(h));
  std::memcpy(h.magic, "QRB1", 4);
  if (!stat_source(filename, &h)) return;
  h.nrows = ds.num_instances(); h.nfeatures = ds.num_features(); h.nqueries = ds.num_queries();
  const std::string tmp = filename + ".qrb.tmp" + std::to_string((long) getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return;   // read-only directory: no cache, no error
  const size_t cells = (size_t) h.nrows * h.nfeatures;
  const bool ok = fwrite(&h, sizeof(h), 1, f) == 1 &&
                  fwrite(ds.labels(), sizeof(Label), h.nrows, f) == h.nrows &&
                  fwrite(ds.offsets().data(), sizeof(uint64_t), ds.offsets().size(), f) == ds.offsets().size() &&
                  (cells == 0 || fwrite(ds.data(), sizeof(Feature), cells, f) == cells);
  if (fclose(f) != 0 || !ok || rename(tmp.c_str(), (filename + ".qrb").c_str()) != 0) remove(tmp.c_str());
}
}  // namespace

std::unique_ptr<data::Dataset> Svml::read_horizontal(const std::string &filename) {
  if (svml_cache_enabled()) {
    std::unique_ptr<data::Dataset> cached = load_svml_cache(filename);
    if (cached) return cached;
  }
  FILE *f = fopen(filename.c_str(), "rb");
  if (!f) {
    std::cerr << "!!! Error while opening file " << filename << "." << std::endl;
    exit(EXIT_FAILURE);
  }
  fseek(f, 0, SEEK_END);
  const long fsize = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t) std::max<long>(fsize, 0) + 1);
  if (fsize > 0 && fread(buf.data(), 1, (size_t) fsize, f) != (size_t) fsize) {
    std::cerr << "!!! Error while reading file " << filename << "." << std::endl;
    exit(EXIT_FAILURE);
  }
  fclose(f);
  buf[(size_t) std::max<long>(fsize, 0)] = '\0';
  char *base = buf.data(), *end = base + std::max<long>(fsize, 0);
  unsigned nthreads = std::thread::hardware_concurrency();
  // the ranks of an external launcher each parse their copy on the same host: share the cores between them
  // (quicklearn's own `--gpus N` parses once, before it forks the ranks)
  if (const char *lw = getenv("LOCAL_WORLD_SIZE")) nthreads = std::max(1u, nthreads / (unsigned) std::max(1, atoi(lw)));
  if (const char *env = getenv("QR_SVML_THREADS")) nthreads = (unsigned) atoi(env);
  nthreads = std::max(1u, std::min(nthreads, (unsigned) (fsize / (1 << 20)) + 1u));   // >= 1 MB per thread
  // chunk boundaries on line starts
  std::vector<char *> cut(nthreads + 1);
  cut[0] = base;
  cut[nthreads] = end;
  for (unsigned t = 1; t < nthreads; ++t) {
    char *p = base + (size_t) fsize * t / nthreads;
    char *nl = (char *) memchr(p, '\n', (size_t) (end - p));
    cut[t] = nl ? nl + 1 : end;
  }
  for (unsigned t = 1; t <= nthreads; ++t) cut[t] = std::max(cut[t], cut[t - 1]);
  std::vector<SvmlChunk> chunks(nthreads);
  {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(parse_svml_range, cut[t], cut[t + 1], std::ref(chunks[t]));
    parse_svml_range(cut[0], cut[1], chunks[0]);
    for (auto &th : pool) th.join();
  }
  size_t nrows = 0, maxfid = 0;
  for (auto &c : chunks) {
    if (c.error) exit(c.error);   // the reference's exit codes for malformed lines (svml.cc:91,112)
    nrows += c.qids.size();
    maxfid = std::max(maxfid, c.maxfid);
  }
  std::unique_ptr<data::Dataset> ds(new data::Dataset(nrows, maxfid));
  // labels and query boundaries in file order (cheap), then every thread scatters its own rows into the
  // zero-initialised row-major matrix
  const std::vector<Feature> none;
  std::vector<size_t> first_row(nthreads + 1, 0);
  for (unsigned t = 0; t < nthreads; ++t) {
    for (size_t i = 0; i < chunks[t].qids.size(); ++i) ds->addInstance(chunks[t].qids[i], chunks[t].labels[i], none);
    first_row[t + 1] = first_row[t] + chunks[t].qids.size();
  }
  auto fill = [&](unsigned t) {
    const SvmlChunk &c = chunks[t];
    for (size_t i = 0; i < c.qids.size(); ++i) {
      Feature *row = ds->at(first_row[t] + i, 0);
      for (size_t k = c.row_start[i]; k < c.row_start[i + 1]; ++k) row[c.fids[k] - 1] = c.vals[k];
    }
  };
  {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(fill, t);
    fill(0);
    for (auto &th : pool) th.join();
  }
  if (svml_cache_enabled()) write_svml_cache(filename, *ds);
  return ds;
}

void Svml::write(std::shared_ptr<data::Dataset> dataset, const std::string &filename) {
  std::ofstream out(filename, std::ofstream::out | std::ofstream::trunc);
  out << std::setprecision(std::numeric_limits<float>::max_digits10);
  for (size_t q = 0; q < dataset->num_queries(); q++) {
    auto r = dataset->getQueryResults(q);
    const Feature *x = r->features();
    for (size_t i = 0; i < r->num_results(); ++i) {
      out << r->labels()[i] << " qid:" << q + 1;
      for (size_t f = 0; f < dataset->num_features(); ++f) out << " " << f + 1 << ":" << x[i * dataset->num_features() + f];
      out << std::endl;
    }
  }
}

}  // namespace io
}  // namespace quickrank

// ------------------------------------------------------------------------------------------------
// trees
// ------------------------------------------------------------------------------------------------
static RTNode *from_flat_rec(const qr_flat_tree &t, int32_t i) {
  if (t.feature[i] < 0) {
    RTNode *n = new RTNode(t.value[i]);
    if (t.count) n->nsampleids = (size_t) t.count[i];
    if (t.deviance) n->deviance = t.deviance[i];
    return n;
  }
  RTNode *l = from_flat_rec(t, t.left[i]);
  RTNode *r = from_flat_rec(t, t.right[i]);
  // featureid = featureidx + 1 (rt.cc:350-352)
  RTNode *n = new RTNode(t.threshold[i], (size_t) t.feature[i], (size_t) t.feature[i] + 1, l, r);
  n->avglabel = t.value ? t.value[i] : 0.0;
  if (t.count) n->nsampleids = (size_t) t.count[i];
  if (t.deviance) n->deviance = t.deviance[i];
  return n;
}

RTNode *RegressionTree::from_flat(const qr_flat_tree &t) { return t.nnodes ? from_flat_rec(t, 0) : nullptr; }

void RegressionTree::to_flat(const RTNode *n, std::vector<int32_t> &feature, std::vector<float> &threshold,
                             std::vector<int32_t> &left, std::vector<int32_t> &right, std::vector<double> &value) {
  const int32_t id = (int32_t) feature.size();
  feature.push_back(n->is_leaf() ? -1 : (int32_t) n->get_feature_idx());
  threshold.push_back(n->threshold);
  left.push_back(-1);
  right.push_back(-1);
  value.push_back(n->avglabel);
  if (!n->is_leaf()) {
    left[id] = (int32_t) feature.size();
    to_flat(n->left, feature, threshold, left, right, value);
    right[id] = (int32_t) feature.size();
    to_flat(n->right, feature, threshold, left, right, value);
  }
}

Ensemble::~Ensemble() {
  for (auto &t : trees_) delete t.root;
}

Ensemble &Ensemble::operator=(Ensemble &&other) {
  if (this != &other) {
    for (auto &t : trees_) delete t.root;
    trees_ = std::move(other.trees_);
    other.trees_.clear();
  }
  return *this;
}

void Ensemble::push(RTNode *root, double weight, float maxlabel) { trees_.push_back({root, weight, maxlabel}); }

void Ensemble::pop() {
  delete trees_.back().root;
  trees_.pop_back();
}

quickrank::Score Ensemble::score_instance(const quickrank::Feature *d, size_t offset) const {
  double sum = 0.0;
  for (auto &t : trees_) sum += t.root->score_instance(d, offset) * t.weight;
  return sum;
}

std::vector<double> Ensemble::get_weights() const {
  std::vector<double> w;
  for (auto &t : trees_) w.push_back(t.weight);
  return w;
}

bool Ensemble::update_ensemble_weights(std::vector<double> &weights) {
  if (weights.size() != trees_.size()) {
    std::cerr << "# ## ERROR!! Ensemble size does not match size of the weight vector in updating the weights"
              << std::endl;
    exit(EXIT_FAILURE);
  }
  for (size_t i = 0; i < trees_.size(); ++i) trees_[i].weight = weights[i];
  return true;
}

bool Ensemble::filter_out_zero_weighted_trees() {   // ensemble.cc:149-169
  size_t kept = 0;
  for (size_t i = 0; i < trees_.size(); ++i) {
    if (trees_[i].weight == 0) delete trees_[i].root;
    else trees_[kept++] = trees_[i];
  }
  trees_.resize(kept);
  return true;
}

bool Ensemble::update_ensemble_weights(std::vector<double> &weights, bool remove) {   // ensemble.cc:171-188
  update_ensemble_weights(weights);
  if (remove) return filter_out_zero_weighted_trees();
  return true;
}

// --- XML (tab indentation, no declaration; schema of mart.cc:470-491, ensemble.cc:133-147,
//     rtnode.cc:48-77) ---
static std::string fmt_g(double v, int digits) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.*g", digits, v);
  return buf;
}

static void tabs(std::ostream &os, int n) {
  for (int i = 0; i < n; ++i) os << '\t';
}

static void write_node(std::ostream &os, const RTNode *n, int indent, const char *pos) {
  tabs(os, indent);
  os << "<split";
  if (pos) os << " pos=\"" << pos << "\"";
  os << ">\n";
  if (n->is_leaf()) {
    tabs(os, indent + 1);
    os << "<output>" << fmt_g(n->avglabel, std::numeric_limits<double>::max_digits10) << "</output>\n";
  } else {
    tabs(os, indent + 1);
    os << "<feature>" << n->get_feature_id() << "</feature>\n";
    tabs(os, indent + 1);
    os << "<threshold>" << fmt_g((double) n->threshold, std::numeric_limits<float>::max_digits10) << "</threshold>\n";
    write_node(os, n->left, indent + 1, "left");
    write_node(os, n->right, indent + 1, "right");
  }
  tabs(os, indent);
  os << "</split>\n";
}

void Ensemble::write_xml(std::ostream &os, int indent) const {
  tabs(os, indent);
  os << "<ensemble>\n";
  for (size_t i = 0; i < trees_.size(); ++i) {
    tabs(os, indent + 1);
    os << "<tree id=\"" << i + 1 << "\" weight=\"" << fmt_g(trees_[i].weight, 17) << "\">\n";
    write_node(os, trees_[i].root, indent + 2, nullptr);
    tabs(os, indent + 1);
    os << "</tree>\n";
  }
  tabs(os, indent);
  os << "</ensemble>\n";
}

namespace quickrank {
namespace learning {
namespace forests {

// Minimal reader for the model schema: elements, attributes, text.  No entities beyond the five
// predefined ones are ever written by either implementation.
struct XmlNode {
  std::string name, text;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const XmlNode *child(const std::string &n) const {
    for (auto &k : kids) if (k->name == n) return k.get();
    return nullptr;
  }
  std::string child_text(const std::string &n, const std::string &def = "") const {
    const XmlNode *c = child(n);
    return c ? c->text : def;
  }
};

struct XmlModel {
  std::unique_ptr<XmlNode> root;   // <ranker>
};

static bool parse_xml(const std::string &s, XmlModel *out) {
  std::vector<XmlNode *> stack;
  std::unique_ptr<XmlNode> top;
  size_t i = 0;
  const size_t n = s.size();
  while (i < n) {
    if (s[i] != '<') {
      const size_t e = s.find('<', i);
      const std::string t = s.substr(i, (e == std::string::npos ? n : e) - i);
      if (!stack.empty()) {
        const size_t a = t.find_first_not_of(" \t\r\n");
        if (a != std::string::npos) stack.back()->text += t.substr(a, t.find_last_not_of(" \t\r\n") - a + 1);
      }
      i = e == std::string::npos ? n : e;
      continue;
    }
    if (s.compare(i, 2, "<?") == 0) { i = s.find("?>", i); if (i == std::string::npos) return false; i += 2; continue; }
    if (s.compare(i, 4, "<!--") == 0) { i = s.find("-->", i); if (i == std::string::npos) return false; i += 3; continue; }
    const size_t e = s.find('>', i);
    if (e == std::string::npos) return false;
    if (s[i + 1] == '/') {
      if (stack.empty()) return false;
      stack.pop_back();
      i = e + 1;
      continue;
    }
    std::string tag = s.substr(i + 1, e - i - 1);
    const bool selfclose = !tag.empty() && tag.back() == '/';
    if (selfclose) tag.pop_back();
    std::unique_ptr<XmlNode> node(new XmlNode());
    size_t p = 0;
    while (p < tag.size() && !isspace((unsigned char) tag[p])) ++p;
    node->name = tag.substr(0, p);
    while (p < tag.size()) {
      while (p < tag.size() && isspace((unsigned char) tag[p])) ++p;
      const size_t eq = tag.find('=', p);
      if (eq == std::string::npos) break;
      const std::string key = tag.substr(p, eq - p);
      const char q = tag[eq + 1];
      const size_t qe = tag.find(q, eq + 2);
      if (qe == std::string::npos) return false;
      node->attr[key] = tag.substr(eq + 2, qe - eq - 2);
      p = qe + 1;
    }
    XmlNode *raw = node.get();
    if (stack.empty()) top = std::move(node);
    else stack.back()->kids.push_back(std::move(node));
    if (!selfclose) stack.push_back(raw);
    i = e + 1;
  }
  if (!top || !stack.empty()) return false;
  out->root = std::move(top);
  return true;
}

// RTNode::parse_xml (rtnode.cc:79-117): featureidx = feature - 1
static RTNode *parse_split(const XmlNode &sp) {
  const XmlNode *out = sp.child("output");
  if (out) return new RTNode(strtod(out->text.c_str(), nullptr));
  const unsigned fid = (unsigned) strtoul(sp.child_text("feature", "0").c_str(), nullptr, 10);
  const float thr = strtof(sp.child_text("threshold", "0").c_str(), nullptr);
  RTNode *l = nullptr, *r = nullptr;
  for (auto &k : sp.kids) {
    if (k->name != "split") continue;
    auto it = k->attr.find("pos");
    if (it != k->attr.end() && it->second == "left") l = parse_split(*k);
    else r = parse_split(*k);
  }
  if (!l || !r) { delete l; delete r; return nullptr; }
  return new RTNode(thr, (size_t) fid - 1, (size_t) fid, l, r);
}

const std::string Mart::NAME_ = "MART";
const std::string LambdaMart::NAME_ = "LAMBDAMART";
const std::string ObliviousMart::NAME_ = "OBVMART";
const std::string ObliviousLambdaMart::NAME_ = "OBVLAMBDAMART";

void Mart::die(const char *what) const {
  std::cerr << "!!! " << what << ": " << qr_last_error() << std::endl;
  exit(EXIT_FAILURE);
}

// Mart(const pugi::xml_document&) (mart.cc:37-89)
Mart::Mart(const XmlModel &model) {
  const XmlNode *info = model.root->child("info");
  const XmlNode *ens = model.root->child("ensemble");
  auto geti = [&](const char *k) { return (size_t) strtoull(info ? info->child_text(k, "0").c_str() : "0", nullptr, 10); };
  ntrees_ = geti("trees");
  nleaves_ = geti("leaves");
  minleafsupport_ = geti("leafsupport");
  nthresholds_ = geti("discretization");
  valid_iterations_ = geti("estop");
  shrinkage_ = info ? strtod(info->child_text("shrinkage", "0").c_str(), nullptr) : 0.0;
  subsample_ = info && info->child("subsample") ? strtof(info->child_text("subsample").c_str(), nullptr) : 1.0f;
  max_features_ = info && info->child("max_features") ? strtof(info->child_text("max_features").c_str(), nullptr) : 1.0f;
  collapse_leaves_factor_ = 0;
  ensemble_model_.set_capacity(ntrees_);
  if (ens)
    for (auto &t : ens->kids) {
      if (t->name != "tree") continue;
      const XmlNode *sp = t->child("split");
      RTNode *root = sp ? parse_split(*sp) : nullptr;
      if (!root) {
        std::cerr << "!!! Unable to parse tree from XML model." << std::endl;
        exit(EXIT_FAILURE);
      }
      auto w = t->attr.find("weight");
      ensemble_model_.push(root, w != t->attr.end() ? strtod(w->second.c_str(), nullptr) : 1.0, -1);
    }
}

ObliviousMart::ObliviousMart(const XmlModel &model) : Mart(model) {
  const XmlNode *info = model.root->child("info");
  treedepth_ = info ? (size_t) strtoull(info->child_text("depth", "0").c_str(), nullptr, 10) : 0;
}

Mart::~Mart() {
  if (ctx_) qr_ctx_destroy(ctx_);
  if (valid_ctx_) qr_ctx_destroy(valid_ctx_);
}

std::ostream &Mart::put(std::ostream &os) const {
  os << "# Ranker: " << name() << std::endl
     << "# max no. of trees = " << ntrees_ << std::endl
     << "# no. of tree leaves = " << nleaves_ << std::endl
     << "# shrinkage = " << shrinkage_ << std::endl
     << "# min leaf support = " << minleafsupport_ << std::endl;
  if (nthresholds_) os << "# no. of thresholds = " << nthresholds_ << std::endl;
  else os << "# no. of thresholds = unlimited" << std::endl;
  if (valid_iterations_) os << "# no. of no gain rounds before early stop = " << valid_iterations_ << std::endl;
  os << "# engine = quickrank_b200 (CUDA sm_100a), histogram mode = "
     << (hist_mode_ == QR_HIST_REFERENCE ? "reference order" : "fixed point") << std::endl;
  return os;
}

std::ostream &ObliviousMart::put(std::ostream &os) const {
  os << "# Ranker: " << name() << std::endl
     << "# max no. of trees = " << ntrees_ << std::endl
     << "# max tree depth = " << treedepth_ << std::endl
     << "# shrinkage = " << shrinkage_ << std::endl
     << "# min leaf support = " << minleafsupport_ << std::endl;
  if (nthresholds_) os << "# no. of thresholds = " << nthresholds_ << std::endl;
  else os << "# no. of thresholds = unlimited" << std::endl;
  if (valid_iterations_) os << "# no. of no gain rounds before early stop = " << valid_iterations_ << std::endl;
  return os;
}

// ---- the hooks: every body is a call into the CUDA library ------------------------------------

void Mart::init(std::shared_ptr<data::VerticalDataset> training_dataset) {
  shard_d0_ = 0;
  qr_params p;
  std::memset(&p, 0, sizeof(p));
  p.algo = algo_id();
  p.nleaves = (uint32_t) nleaves_;
  p.treedepth = (uint32_t) tree_depth();
  p.minleafsupport = (uint32_t) minleafsupport_;
  p.nthresholds = nthresholds_;
  p.ndcg_cutoff = metric_cutoff_ == metric::ir::Metric::NO_CUTOFF ? 0 : metric_cutoff_;
  p.shrinkage = shrinkage_;
  p.hist_mode = hist_mode_;
  p.device = device_;
  auto h = training_dataset->horizontal();
  const host::Sharding &sh = host::sharding();
  if (sh.world > 1) {
    // this process's shard: a contiguous range of whole queries; rows are contiguous in the row-major matrix
    if (h->num_queries() < (size_t) sh.world) {
      std::cerr << "!!! Cannot shard " << h->num_queries() << " queries over " << sh.world << " GPUs." << std::endl;
      exit(EXIT_FAILURE);
    }
    const auto shards = host::query_shards(h->offsets().data(), h->num_queries(), sh.world);
    const size_t q0 = shards[sh.rank].first, q1 = shards[sh.rank].second;
    const uint64_t d0 = h->offsets()[q0], d1 = h->offsets()[q1];
    shard_d0_ = (size_t) d0;
    std::vector<uint64_t> off(q1 - q0 + 1);
    for (size_t q = q0; q <= q1; ++q) off[q - q0] = h->offsets()[q] - d0;
    unsigned char id[QR_COMM_ID_BYTES];
    std::memset(id, 0, sizeof(id));
    if (sh.rank == 0 && qr_comm_unique_id(id) != QR_OK) die("Impossible to create the communicator id");
    if (!host::exchange_bytes(id, sizeof(id), sh)) die("Impossible to exchange the communicator id");
    if (p.device < 0) p.device = sh.local_rank;
    if (qr_ctx_create_sharded(h->data() + d0 * h->num_features(), 1, (size_t) (d1 - d0), h->num_features(),
                              h->labels() + d0, off.data(), q1 - q0, &p, id, sh.rank, sh.world, &ctx_) != QR_OK)
      die("Impossible to initialise the sharded GPU training context");
    return;
  }
  if (qr_ctx_create_rowmajor(h->data(), h->num_instances(), h->num_features(), h->labels(), h->offsets().data(),
                             h->num_queries(), &p, &ctx_) != QR_OK)
    die("Impossible to initialise the GPU training context");
}

void Mart::clear(size_t) {
  if (ctx_) qr_ctx_destroy(ctx_);
  if (valid_ctx_) qr_ctx_destroy(valid_ctx_);
  ctx_ = valid_ctx_ = nullptr;
}

void Mart::compute_pseudoresponses(std::shared_ptr<data::VerticalDataset>, metric::ir::Metric *, bool *sample_presence) {
  if (sample_presence) {
    std::cerr << "!!! Document sub-sampling is not supported by the GPU engine." << std::endl;
    exit(EXIT_FAILURE);
  }
  if (qr_compute_pseudoresponses(ctx_) != QR_OK) die("compute_pseudoresponses");
}

std::unique_ptr<RegressionTree> Mart::fit_tree_on(qr_ctx *ctx) {
  const uint32_t cap = 2 * (uint32_t) std::max<size_t>(nleaves_, 1) + 1;
  std::vector<int32_t> feature(cap), left(cap), right(cap);
  std::vector<uint32_t> tidx(cap);
  std::vector<float> thr(cap);
  std::vector<double> value(cap), dev(cap);
  std::vector<uint64_t> count(cap);
  qr_flat_tree t;
  t.capacity = cap; t.nnodes = t.nleaves = 0;
  t.feature = feature.data(); t.threshold_idx = tidx.data(); t.threshold = thr.data();
  t.left = left.data(); t.right = right.data(); t.value = value.data(); t.deviance = dev.data(); t.count = count.data();
  if (qr_fit_tree(ctx, &t) != QR_OK) die("fit_regressor_on_gradient");
  return std::unique_ptr<RegressionTree>(new RegressionTree(RegressionTree::from_flat(t)));
}

std::unique_ptr<RegressionTree> Mart::fit_regressor_on_gradient(std::shared_ptr<data::VerticalDataset>, size_t *) {
  return fit_tree_on(ctx_);
}

void Mart::update_modelscores(std::shared_ptr<data::VerticalDataset>, Score *, RegressionTree *) {
  if (qr_update_modelscores(ctx_, shrinkage_) != QR_OK) die("update_modelscores");
}

// validation set: the new tree is applied on the device to the validation context
void Mart::update_modelscores(std::shared_ptr<data::Dataset>, Score *, RegressionTree *tree) {
  apply_tree_on(valid_ctx_, tree);
}

void Mart::apply_tree_on(qr_ctx *target, RegressionTree *tree) {
  std::vector<int32_t> feature, left, right;
  std::vector<float> thr;
  std::vector<double> value;
  RegressionTree::to_flat(tree->get_proot(), feature, thr, left, right, value);
  std::vector<uint32_t> tidx(feature.size(), 0);
  for (size_t i = 0; i < feature.size(); ++i) {
    if (feature[i] < 0) continue;
    const float *tv = nullptr;
    size_t tn = 0;
    if (qr_get_thresholds(ctx_, (size_t) feature[i], &tv, &tn) != QR_OK) die("update_modelscores");
    const float *it = std::find(tv, tv + tn, thr[i]);
    tidx[i] = (uint32_t) (it - tv);
  }
  qr_flat_tree t;
  t.capacity = t.nnodes = (uint32_t) feature.size();
  t.nleaves = 0;
  t.feature = feature.data(); t.threshold_idx = tidx.data(); t.threshold = thr.data();
  t.left = left.data(); t.right = right.data(); t.value = value.data(); t.deviance = nullptr; t.count = nullptr;
  if (qr_apply_tree(target, &t, shrinkage_) != QR_OK) die("update_modelscores (tree applied to a dataset)");
}

MetricScore Mart::evaluate_training(metric::ir::Metric *) {
  double m = 0;
  if (qr_evaluate(ctx_, &m) != QR_OK) die("evaluate_dataset");
  return m;
}

bool Mart::import_model_state(LTR_Algorithm &other) {
  Mart *o = dynamic_cast<Mart *>(&other);
  if (!o) return false;
  if (std::abs(shrinkage_ - o->shrinkage_) > 0.000001 || nthresholds_ != o->nthresholds_ || nleaves_ != o->nleaves_ ||
      minleafsupport_ != o->minleafsupport_ || valid_iterations_ != o->valid_iterations_)
    return false;
  ensemble_model_ = std::move(o->ensemble_model_);
  return true;
}

bool ObliviousMart::import_model_state(LTR_Algorithm &other) {
  ObliviousMart *o = dynamic_cast<ObliviousMart *>(&other);
  if (!o || treedepth_ != o->treedepth_) return false;
  return Mart::import_model_state(other);
}

// Mart::learn (mart.cc:208-416): same loop, same stdout table
void Mart::learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
                 std::shared_ptr<metric::ir::Metric> scorer, size_t partial_save, const std::string output_basename) {
  if (scorer->name() != "NDCG") {
    std::cerr << "!!! The GPU engine optimises NDCG only (got " << scorer->name() << ")." << std::endl;
    exit(EXIT_FAILURE);
  }
  if (subsample_ != 1.0f || max_features_ != 1.0f || collapse_leaves_factor_ != 0.0f) {
    std::cerr << "!!! subsample, max_features and collapse_leaves_factor are not supported by the GPU engine."
              << std::endl;
    exit(EXIT_FAILURE);
  }
  std::cout << "# Initialization";
  std::cout.flush();
  auto t0 = std::chrono::high_resolution_clock::now();
  metric_cutoff_ = scorer->cutoff();
  std::shared_ptr<data::VerticalDataset> vertical_training(new data::VerticalDataset(training_dataset));
  best_metric_on_validation_ = std::numeric_limits<double>::lowest();
  best_metric_on_training_ = std::numeric_limits<double>::lowest();
  best_model_ = 0;
  ensemble_model_.set_capacity(ntrees_);
  init(vertical_training);
  if (validation_dataset) {
    if (qr_ctx_create_eval(ctx_, validation_dataset->data(), validation_dataset->num_instances(),
                           validation_dataset->num_features(), validation_dataset->labels(),
                           validation_dataset->offsets().data(), validation_dataset->num_queries(), &valid_ctx_) != QR_OK)
      die("Impossible to initialise the GPU validation context");
  }
  if (ensemble_model_.is_notempty()) {   // restart from a loaded model (mart.cc:237-253)
    best_model_ = ensemble_model_.get_size() - 1;
    std::vector<Score> s(training_dataset->num_instances());
    score_dataset(training_dataset, s.data());
    // (sharded training: the context holds this rank's documents only)
    if (qr_set_scores(ctx_, s.data() + shard_d0_) != QR_OK) die("restart");
    best_metric_on_training_ = evaluate_training(scorer.get());
    if (validation_dataset) {
      std::vector<Score> v(validation_dataset->num_instances());
      score_dataset(validation_dataset, v.data());
      if (qr_set_scores(valid_ctx_, v.data()) != QR_OK) die("restart");
      if (qr_evaluate(valid_ctx_, &best_metric_on_validation_) != QR_OK) die("restart");
    }
  }
  double init_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  std::cout << ": " << std::setprecision(2) << init_time << " s." << std::endl;

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
  auto t1 = std::chrono::high_resolution_clock::now();
  for (size_t m = ensemble_model_.get_size(); m < ntrees_; ++m) {
    if (validation_dataset && (valid_iterations_ && m > best_model_ + valid_iterations_)) break;
    compute_pseudoresponses(vertical_training, scorer.get(), nullptr);
    // (the root-histogram refresh of mart.cc:335 happens inside fit_regressor_on_gradient)
    std::unique_ptr<RegressionTree> tree = fit_regressor_on_gradient(vertical_training, nullptr);
    ensemble_model_.push(tree->get_proot(), shrinkage_, 0);
    update_modelscores(vertical_training, nullptr, tree.get());
    const MetricScore metric_on_training = evaluate_training(scorer.get());
    std::cout << std::setw(7) << m + 1 << std::setw(9) << metric_on_training;
    if (validation_dataset) {
      update_modelscores(validation_dataset, nullptr, tree.get());
      MetricScore metric_on_validation = 0;
      if (qr_evaluate(valid_ctx_, &metric_on_validation) != QR_OK) die("evaluate_dataset (validation)");
      std::cout << std::setw(9) << metric_on_validation;
      if (metric_on_validation > best_metric_on_validation_) {
        best_metric_on_training_ = metric_on_training;
        best_metric_on_validation_ = metric_on_validation;
        best_model_ = ensemble_model_.get_size() - 1;
        std::cout << " *";
      }
    } else if (metric_on_training > best_metric_on_training_) {
      best_metric_on_training_ = metric_on_training;
      best_model_ = ensemble_model_.get_size() - 1;
      std::cout << " *";
    }
    std::cout << std::endl;
    if (partial_save != 0 && !output_basename.empty() && (m + 1) % partial_save == 0) save(output_basename, (int) (m + 1));
  }
  if (validation_dataset)   // roll back to the best model on validation (mart.cc:390-395)
    while (ensemble_model_.is_notempty() && ensemble_model_.get_size() > best_model_ + 1) ensemble_model_.pop();
  double train_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count();
  std::cout << std::endl;
  std::cout << *scorer << " on training data = " << best_metric_on_training_ << std::endl;
  if (validation_dataset) std::cout << *scorer << " on validation data = " << best_metric_on_validation_ << std::endl;
  clear(vertical_training->num_features());
  std::cout << std::endl;
  std::cout << "#\t Training Time: " << std::setprecision(2) << train_time << " s." << std::endl;
}

// LTR_Algorithm::score_dataset (ltr_algorithm.cc:44-52) on the GPU
void Mart::score_dataset(std::shared_ptr<data::Dataset> dataset, Score *scores) const {
  const size_t nt = ensemble_model_.get_size();
  std::vector<std::vector<int32_t>> feature(nt), left(nt), right(nt);
  std::vector<std::vector<float>> thr(nt);
  std::vector<std::vector<double>> value(nt);
  std::vector<qr_flat_tree> flat(nt);
  std::vector<double> weights(nt);
  for (size_t i = 0; i < nt; ++i) {
    RegressionTree::to_flat(ensemble_model_.getTree((int) i), feature[i], thr[i], left[i], right[i], value[i]);
    qr_flat_tree &t = flat[i];
    t.capacity = t.nnodes = (uint32_t) feature[i].size();
    t.nleaves = 0;
    t.feature = feature[i].data(); t.threshold_idx = nullptr; t.threshold = thr[i].data();
    t.left = left[i].data(); t.right = right[i].data(); t.value = value[i].data(); t.deviance = nullptr; t.count = nullptr;
    weights[i] = ensemble_model_.getWeight((int) i);
  }
  qr_scorer *sc = nullptr;
  if (qr_scorer_create(flat.data(), weights.data(), nt, dataset->num_features(), device_, &sc) != QR_OK) die("score_dataset");
  if (qr_score_dataset(sc, dataset->data(), dataset->num_instances(), dataset->num_features(), scores) != QR_OK)
    die("score_dataset");
  qr_scorer_destroy(sc);
}

void Mart::write_xml_info(std::ostream &os) const {
  os << "\t\t<type>" << name() << "</type>\n"
     << "\t\t<trees>" << ntrees_ << "</trees>\n"
     << "\t\t<leaves>" << nleaves_ << "</leaves>\n"
     << "\t\t<shrinkage>" << fmt_g(shrinkage_, 17) << "</shrinkage>\n"
     << "\t\t<leafsupport>" << minleafsupport_ << "</leafsupport>\n"
     << "\t\t<discretization>" << nthresholds_ << "</discretization>\n"
     << "\t\t<estop>" << valid_iterations_ << "</estop>\n"
     << "\t\t<subsample>" << fmt_g(subsample_, 9) << "</subsample>\n"
     << "\t\t<max_features>" << fmt_g(max_features_, 9) << "</max_features>\n"
     << "\t\t<collapse_leaves_factor>" << fmt_g(collapse_leaves_factor_, 9) << "</collapse_leaves_factor>\n";
}

// obliviousmart.cc:74-81 (the reference writes the discretisation into <estop> as well)
void ObliviousMart::write_xml_info(std::ostream &os) const {
  os << "\t\t<type>" << name() << "</type>\n"
     << "\t\t<trees>" << ntrees_ << "</trees>\n"
     << "\t\t<leaves>" << nleaves_ << "</leaves>\n"
     << "\t\t<depth>" << treedepth_ << "</depth>\n"
     << "\t\t<shrinkage>" << fmt_g(shrinkage_, 17) << "</shrinkage>\n"
     << "\t\t<leafsupport>" << minleafsupport_ << "</leafsupport>\n"
     << "\t\t<discretization>" << nthresholds_ << "</discretization>\n"
     << "\t\t<estop>" << nthresholds_ << "</estop>\n";
}

void Mart::write_xml_model(std::ostream &os) const {
  os << "<ranker>\n\t<info>\n";
  write_xml_info(os);
  os << "\t</info>\n";
  ensemble_model_.write_xml(os, 1);
  os << "</ranker>\n";
}


// ------------------------------------------------------------------------------------------
// DART (dart.cc).  Host logic only; documents are touched on the GPU.
// ------------------------------------------------------------------------------------------
const std::string Dart::NAME_ = "DART";

static const std::vector<std::string> kDartSampling = {"UNIFORM", "WEIGHTED", "WEIGHTED_INV", "TOP_FIFTY", "CONTR",
                                                       "CONTR_INV", "WCONTR", "WCONTR_INV", "TOP_WCONTR", "LESS_WCONTR"};
static const std::vector<std::string> kDartNormalization = {"TREE", "NONE", "WEIGHTED", "FOREST", "TREE_ADAPTIVE",
                                                            "LINESEARCH", "TREE_BOOST3", "CONTR", "WCONTR", "LMART_ADAPTIVE"};
static const std::vector<std::string> kDartAdaptive = {"FIXED", "PLUS1_DIV2", "PLUSHALF_DIV2", "PLUSONETHIRD_DIV2",
                                                       "PLUSHALF_RESET", "PLUSHALF_RESET_LB1_UB5",
                                                       "PLUSHALF_RESET_LB1_UB10", "PLUSHALF_RESET_LB1_UBRD"};
// the reference maps names to enum values through the index in these tables (dart.h:130-170);
// its sampling table skips the COUNT* enumerators, so only the names whose index equals the
// enumerator are usable there too — the ones this build supports all are
static int dart_index(const std::vector<std::string> &names, std::string name, const char *what) {
  std::transform(name.begin(), name.end(), name.begin(), ::toupper);
  for (size_t i = 0; i < names.size(); ++i)
    if (names[i] == name) return (int) i;
  std::cerr << "!!! Unknown DART " << what << " " << name << std::endl;
  exit(EXIT_FAILURE);
}
Dart::SamplingType Dart::get_sampling_type(std::string name) {
  const int i = dart_index(kDartSampling, name, "sample type");
  if (kDartSampling[i] == "UNIFORM") return SamplingType::UNIFORM;
  if (kDartSampling[i] == "TOP_FIFTY") return SamplingType::TOP_FIFTY;
  if (kDartSampling[i] == "WEIGHTED") return SamplingType::WEIGHTED;
  return SamplingType::CONTR;   // any unsupported one: rejected by check_supported()
}
Dart::NormalizationType Dart::get_normalization_type(std::string name) {
  return (NormalizationType) dart_index(kDartNormalization, name, "normalization type");
}
Dart::AdaptiveType Dart::get_adaptive_type(std::string name) {
  return (AdaptiveType) dart_index(kDartAdaptive, name, "adaptive type");
}
std::string Dart::get_sampling_type(SamplingType t) {
  switch (t) {
    case SamplingType::UNIFORM: return "UNIFORM";
    case SamplingType::TOP_FIFTY: return "TOP_FIFTY";
    case SamplingType::WEIGHTED: return "WEIGHTED";
    default: return "CONTR";
  }
}
std::string Dart::get_normalization_type(NormalizationType t) { return kDartNormalization[(size_t) t]; }
std::string Dart::get_adaptive_type(AdaptiveType t) { return kDartAdaptive[(size_t) t]; }

Dart::Dart(size_t ntrees, double shrinkage, size_t nthresholds, size_t ntreeleaves, size_t minleafsupport, float subsample,
           float max_features, size_t valid_iterations, float collapse_leaves_factor, SamplingType st,
           NormalizationType nt, AdaptiveType at, double rd, double sd, bool kd, bool bot, double rk, double dob)
    : LambdaMart(ntrees, shrinkage, nthresholds, ntreeleaves, minleafsupport, subsample, max_features, valid_iterations,
                 collapse_leaves_factor),
      sample_type(st), normalize_type(nt), adaptive_type(at), rate_drop(rd), skip_drop(sd), keep_drop(kd),
      best_on_train(bot), random_keep(rk), drop_on_best(dob != 0.0) {
  check_supported();
}

// Dart(const pugi::xml_document&) (dart.cc:54-98)
Dart::Dart(const XmlModel &model) : LambdaMart(model) {
  const XmlNode *info = model.root->child("info");
  if (info) {
    sample_type = get_sampling_type(info->child_text("sample_type", "UNIFORM"));
    normalize_type = get_normalization_type(info->child_text("normalize_type", "TREE"));
    adaptive_type = get_adaptive_type(info->child_text("adaptive_type", "FIXED"));
    rate_drop = strtod(info->child_text("rate_drop", "0.1").c_str(), nullptr);
    skip_drop = strtod(info->child_text("skip_drop", "0").c_str(), nullptr);
    best_on_train = info->child_text("best_on_train", "false") == "true";
    random_keep = strtod(info->child_text("random_keep", "0").c_str(), nullptr);
    drop_on_best = info->child_text("drop_on_best", "false") == "true";
  }
}

bool Dart::import_model_state(LTR_Algorithm &other) {
  Dart *o = dynamic_cast<Dart *>(&other);
  if (!o) return false;
  if (std::abs(shrinkage_ - o->shrinkage_) > 0.000001 || nthresholds_ != o->nthresholds_ || nleaves_ != o->nleaves_ ||
      minleafsupport_ != o->minleafsupport_ || valid_iterations_ != o->valid_iterations_ ||
      sample_type != o->sample_type || normalize_type != o->normalize_type || rate_drop != o->rate_drop ||
      skip_drop != o->skip_drop)
    return false;
  ensemble_model_ = std::move(o->ensemble_model_);
  return true;
}

void Dart::check_supported() const {
  const bool st_ok = sample_type == SamplingType::UNIFORM || sample_type == SamplingType::TOP_FIFTY;
  const bool nt_ok = normalize_type == NormalizationType::TREE || normalize_type == NormalizationType::NONE ||
                     normalize_type == NormalizationType::WEIGHTED || normalize_type == NormalizationType::FOREST ||
                     normalize_type == NormalizationType::TREE_BOOST3;
  if (!st_ok || !nt_ok || adaptive_type != AdaptiveType::FIXED) {
    std::cerr << "!!! This DART variant (sample type " << get_sampling_type(sample_type) << ", normalization "
              << get_normalization_type(normalize_type) << ", adaptive type " << get_adaptive_type(adaptive_type)
              << ") is not supported by the GPU engine: UNIFORM|TOP_FIFTY x TREE|NONE|WEIGHTED|FOREST|TREE_BOOST3 x FIXED."
              << std::endl;
    exit(EXIT_FAILURE);
  }
}

void Dart::write_xml_info(std::ostream &os) const {   // dart.cc:108-133
  os << "\t\t<type>" << name() << "</type>\n"
     << "\t\t<trees>" << ntrees_ << "</trees>\n"
     << "\t\t<leaves>" << nleaves_ << "</leaves>\n"
     << "\t\t<shrinkage>" << fmt_g(shrinkage_, 17) << "</shrinkage>\n"
     << "\t\t<leafsupport>" << minleafsupport_ << "</leafsupport>\n"
     << "\t\t<discretization>" << nthresholds_ << "</discretization>\n"
     << "\t\t<estop>" << valid_iterations_ << "</estop>\n"
     << "\t\t<sample_type>" << get_sampling_type(sample_type) << "</sample_type>\n"
     << "\t\t<normalize_type>" << get_normalization_type(normalize_type) << "</normalize_type>\n"
     << "\t\t<adaptive_type>" << get_adaptive_type(adaptive_type) << "</adaptive_type>\n"
     << "\t\t<rate_drop>" << fmt_g(rate_drop, 17) << "</rate_drop>\n"
     << "\t\t<skip_drop>" << fmt_g(skip_drop, 17) << "</skip_drop>\n"
     << "\t\t<best_on_train>" << (best_on_train ? "true" : "false") << "</best_on_train>\n"
     << "\t\t<random_keep>" << fmt_g(random_keep, 17) << "</random_keep>\n"
     << "\t\t<drop_on_best>" << (drop_on_best ? "true" : "false") << "</drop_on_best>\n";
}

std::ostream &Dart::put(std::ostream &os) const {   // dart.cc:142-170
  os << "# Ranker: " << name() << std::endl
     << "# max no. of trees = " << ntrees_ << std::endl
     << "# no. of tree leaves = " << nleaves_ << std::endl
     << "# shrinkage = " << shrinkage_ << std::endl
     << "# min leaf support = " << minleafsupport_ << std::endl;
  if (nthresholds_) os << "# no. of thresholds = " << nthresholds_ << std::endl;
  else os << "# no. of thresholds = unlimited" << std::endl;
  if (valid_iterations_) os << "# no. of no gain rounds before early stop = " << valid_iterations_ << std::endl;
  os << "# sample type = " << get_sampling_type(sample_type) << std::endl
     << "# normalization type = " << get_normalization_type(normalize_type) << std::endl
     << "# adaptive type = " << get_adaptive_type(adaptive_type) << std::endl
     << "# rate drop = " << rate_drop << std::endl
     << "# skip drop = " << skip_drop << std::endl
     << "# keep drop = " << keep_drop << std::endl
     << "# best on train = " << best_on_train << std::endl
     << "# keep dropout at random = " << random_keep << std::endl
     << "# keep dropout based on best = " << drop_on_best << std::endl;
  return os;
}

struct Dart::DeviceTree {
  std::vector<int32_t> feature, left, right;
  std::vector<uint32_t> tidx;
  std::vector<float> thr;
  std::vector<double> value;
  qr_flat_tree flat() {
    qr_flat_tree t;
    t.capacity = t.nnodes = (uint32_t) feature.size();
    t.nleaves = 0;
    t.feature = feature.data(); t.threshold_idx = tidx.data(); t.threshold = thr.data();
    t.left = left.data(); t.right = right.data(); t.value = value.data(); t.deviance = nullptr; t.count = nullptr;
    return t;
  }
};

// a tree of the ensemble expressed on this run's bins (threshold value -> threshold index)
std::shared_ptr<Dart::DeviceTree> Dart::make_flat(const RTNode *root) const {
  auto d = std::make_shared<DeviceTree>();
  RegressionTree::to_flat(root, d->feature, d->thr, d->left, d->right, d->value);
  d->tidx.assign(d->feature.size(), 0);
  for (size_t i = 0; i < d->feature.size(); ++i) {
    if (d->feature[i] < 0) continue;
    const float *tv = nullptr;
    size_t tn = 0;
    if (qr_get_thresholds(ctx_, (size_t) d->feature[i], &tv, &tn) != QR_OK) die("DART: thresholds");
    // smallest index whose threshold is >= the node's: x <= thr  <=>  bin(x) <= that index
    const float *it = std::lower_bound(tv, tv + tn, d->thr[i]);
    d->tidx[i] = (uint32_t) (it - tv);
  }
  return d;
}

void Dart::update_modelscores_trees(qr_ctx *ctx, bool add, const std::vector<int> &trees) {
  const double sign = add ? 1.0 : -1.0;
  std::vector<qr_flat_tree> ft;
  std::vector<double> w;
  for (int t : trees) {
    ft.push_back(flat_[(size_t) t]->flat());
    w.push_back(sign * ensemble_model_.getWeight(t));
  }
  // one pass over the documents for the whole set; per document the trees are applied in this order
  if (qr_apply_trees(ctx, ft.data(), w.data(), ft.size()) != QR_OK) die("DART: update_modelscores");
}

// full rescoring of a dataset with the current ensemble (score_dataset in dart.cc:552-558)
void Dart::rescore_on_device(qr_ctx *ctx, std::shared_ptr<data::Dataset> dataset) {
  std::vector<Score> zero(dataset->num_instances(), 0.0);
  if (qr_set_scores(ctx, zero.data()) != QR_OK) die("DART: rescore");
  std::vector<int> all;
  // Ensemble::score_instance sums weight * tree in index order (ensemble.cc:111-118); so does this loop
  for (size_t t = 0; t < ensemble_model_.get_size(); ++t) all.push_back((int) t);
  update_modelscores_trees(ctx, true, all);
}

// dart.cc:708-736 (UNIFORM / TOP_FIFTY)
std::vector<int> Dart::select_trees_to_dropout(std::vector<double> &weights, size_t trees_to_dropout) {
  if (trees_to_dropout == 0) return std::vector<int>();
  std::vector<int> dropped;
  size_t size = weights.size();
  if (sample_type == SamplingType::TOP_FIFTY) size = (size_t) round(size / 2);
  std::vector<int> idx(size);
  std::iota(idx.begin(), idx.end(), 0);
  // libstdc++'s std::random_shuffle(first, last, rng): for i in 1..n-1 swap(i, rng(i + 1)), with the
  // reference's generator (dart.cc:725-729); written out because std::random_shuffle is gone in C++17
  for (size_t i = 1; i < idx.size(); ++i) {
    const size_t j = (size_t) (int) (std::rand() / (1.0 + RAND_MAX) * (int) (i + 1));
    if (i != j) std::swap(idx[i], idx[j]);
  }
  for (size_t i = 0; dropped.size() < trees_to_dropout && i < idx.size(); ++i)
    if (weights[(size_t) idx[i]] > 0) dropped.push_back(idx[i]);
  return dropped;
}

// dart.cc:856-905
void Dart::normalize_trees_restore_drop(std::vector<double> &weights, const std::vector<int> &dropped_trees,
                                        double /*last_tree_weight*/) {
  const size_t k = dropped_trees.size();
  if (normalize_type == NormalizationType::TREE || normalize_type == NormalizationType::TREE_BOOST3) {
    const double alpha = normalize_type == NormalizationType::TREE_BOOST3 ? 3 : 1;
    weights.push_back((shrinkage_ * alpha) / ((shrinkage_ * alpha) + k));
    const double norm = (double) k / (k + (shrinkage_ * alpha));
    for (int idx : dropped_trees) weights[(size_t) idx] *= norm;
  } else if (normalize_type == NormalizationType::NONE) {
    weights.push_back(shrinkage_);
  } else if (normalize_type == NormalizationType::WEIGHTED) {
    double sum = 0;
    for (int t : dropped_trees) sum += weights[(size_t) t];
    const double sumWithLast = sum + shrinkage_;
    const double norm = sum / sumWithLast;
    weights.push_back(shrinkage_ / sumWithLast);
    for (int t : dropped_trees) weights[(size_t) t] *= norm;
  } else if (normalize_type == NormalizationType::FOREST) {
    weights.push_back(shrinkage_ / (1 + shrinkage_));
    const double norm = 1 / (1 + shrinkage_);
    for (int idx : dropped_trees) weights[(size_t) idx] *= norm;
  }
}

// dart.cc:1095-1181 (FIXED)
int Dart::get_number_of_trees_to_dropout(std::vector<double> &dropout_factor_per_iter, int dropped_before_cleaning) {
  const double prob_skip_dropout = (double) rand() / (double) (RAND_MAX);
  const int model_size = (int) ensemble_model_.get_size() - dropped_before_cleaning;
  double trees_to_dropout = 0;
  if (prob_skip_dropout > skip_drop && model_size > 0) {
    if (rate_drop >= 1) {
      if ((rate_drop * 2) <= model_size) trees_to_dropout = rate_drop;
    } else {
      trees_to_dropout = rate_drop * model_size;
    }
  }
  trees_to_dropout = trees_to_dropout > model_size / 2 ? model_size / 2 : trees_to_dropout;   // integer division, as there
  dropout_factor_per_iter.push_back(trees_to_dropout);
  return (int) round(trees_to_dropout);
}

// Dart::learn (dart.cc:172-602)
void Dart::learn(std::shared_ptr<data::Dataset> training_dataset, std::shared_ptr<data::Dataset> validation_dataset,
                 std::shared_ptr<metric::ir::Metric> scorer, size_t partial_save, const std::string output_basename) {
  if (scorer->name() != "NDCG") {
    std::cerr << "!!! The GPU engine optimises NDCG only (got " << scorer->name() << ")." << std::endl;
    exit(EXIT_FAILURE);
  }
  if (subsample_ != 1.0f || max_features_ != 1.0f || collapse_leaves_factor_ != 0.0f) {
    std::cerr << "!!! subsample, max_features and collapse_leaves_factor are not supported by the GPU engine."
              << std::endl;
    exit(EXIT_FAILURE);
  }
  check_supported();
  std::cout << "# Initialization";
  std::cout.flush();
  std::srand(0);   // dart.cc:181
  auto t0 = std::chrono::high_resolution_clock::now();
  metric_cutoff_ = scorer->cutoff();
  metric_trace_.clear();
  std::shared_ptr<data::VerticalDataset> vertical_training(new data::VerticalDataset(training_dataset));
  best_metric_on_validation_ = std::numeric_limits<double>::lowest();
  best_metric_on_training_ = std::numeric_limits<double>::lowest();
  best_model_ = 0;
  size_t best_iter_ = 0;
  std::vector<double> best_weights;
  ensemble_model_.set_capacity(ntrees_ + valid_iterations_);
  init(vertical_training);
  if (validation_dataset &&
      qr_ctx_create_eval(ctx_, validation_dataset->data(), validation_dataset->num_instances(),
                         validation_dataset->num_features(), validation_dataset->labels(),
                         validation_dataset->offsets().data(), validation_dataset->num_queries(), &valid_ctx_) != QR_OK)
    die("Impossible to initialise the GPU validation context");
  auto eval = [&](qr_ctx *ctx) {
    double m = 0;
    if (qr_evaluate(ctx, &m) != QR_OK) die("evaluate_dataset");
    metric_trace_.push_back(m);
    return m;
  };
  flat_.clear();
  if (ensemble_model_.is_notempty()) {   // restart from a loaded model (dart.cc:206-226)
    for (size_t t = 0; t < ensemble_model_.get_size(); ++t) flat_.push_back(make_flat(ensemble_model_.getTree((int) t)));
    best_model_ = ensemble_model_.get_size() - 1;
    best_iter_ = best_model_;
    best_weights = ensemble_model_.get_weights();
    rescore_on_device(ctx_, training_dataset);
    best_metric_on_training_ = eval(ctx_);
    if (validation_dataset) {
      rescore_on_device(valid_ctx_, validation_dataset);
      best_metric_on_validation_ = eval(valid_ctx_);
    }
  }
  double init_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  std::cout << ": " << std::setprecision(2) << init_time << " s." << std::endl;
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
  auto t1 = std::chrono::high_resolution_clock::now();
  MetricScore metric_on_training = std::numeric_limits<double>::lowest();
  MetricScore metric_on_validation = std::numeric_limits<double>::lowest();
  size_t dropped_before_cleaning = 0;
  size_t m = (size_t) -1;
  size_t last_iteration_global_scoring = 0;
  std::vector<double> dropout_factor_per_iter;
  while ((ensemble_model_.get_size() - dropped_before_cleaning) < ntrees_) {
    ++m;
    if (validation_dataset && (valid_iterations_ && m > best_iter_ + valid_iterations_)) break;
    std::vector<double> orig_weights = ensemble_model_.get_weights();
    const int trees_to_dropout = get_number_of_trees_to_dropout(dropout_factor_per_iter, (int) dropped_before_cleaning);
    const double prob_random_keep = (double) rand() / (double) (RAND_MAX);
    const bool random_keep_iter = trees_to_dropout > 0 && prob_random_keep <= random_keep;
    double metric_on_training_dropout = 0, metric_on_validation_dropout = 0;
    std::vector<int> dropped_trees;
    bool dropout_better_than_full = false;
    std::vector<double> dropped_weights(orig_weights);
    if (trees_to_dropout > 0) {
      dropped_trees = select_trees_to_dropout(orig_weights, (size_t) trees_to_dropout);
      // subtract the dropped trees from the scores (train and validation)
      update_modelscores_trees(ctx_, false, dropped_trees);
      metric_on_training_dropout = eval(ctx_);
      if (validation_dataset) {
        update_modelscores_trees(valid_ctx_, false, dropped_trees);
        metric_on_validation_dropout = eval(valid_ctx_);
        if (metric_on_validation_dropout > metric_on_validation) dropout_better_than_full = true;
      } else if (metric_on_training_dropout > metric_on_training) {
        dropout_better_than_full = true;
      }
      for (int idx : dropped_trees) dropped_weights[(size_t) idx] = 0;
      ensemble_model_.update_ensemble_weights(dropped_weights, false);
    }
    compute_pseudoresponses(vertical_training, scorer.get(), nullptr);
    std::unique_ptr<RegressionTree> tree = fit_regressor_on_gradient(vertical_training, nullptr);
    // (update_contribution_scores, dart.cc:378, only feeds the CONTR normalisations, which are rejected)
    // get_weight_last_tree (dart.cc:944-975) for the supported normalisations
    const double tree_weight = normalize_type == NormalizationType::TREE_BOOST3
                                   ? (shrinkage_ * 3) / ((shrinkage_ * 3) + dropped_trees.size())
                                   : shrinkage_;
    ensemble_model_.push(tree->get_proot(), tree_weight, 0);
    flat_.push_back(make_flat(tree->get_proot()));
    const int lastTreeIndex = (int) ensemble_model_.get_size() - 1;
    std::vector<int> lastTree = {lastTreeIndex};
    update_modelscores_trees(ctx_, true, lastTree);
    const double metric_on_training_fit = eval(ctx_);
    double metric_on_validation_fit = std::numeric_limits<double>::lowest();
    if (validation_dataset) {
      update_modelscores_trees(valid_ctx_, true, lastTree);
      metric_on_validation_fit = eval(valid_ctx_);
    }
    bool fit_after_dropout_improvement = false;
    if (trees_to_dropout > 0) {
      double reference_metric_training = metric_on_training, reference_metric_validation = metric_on_validation;
      if (drop_on_best) {
        reference_metric_training = best_metric_on_training_;
        reference_metric_validation = best_metric_on_validation_;
      }
      if (validation_dataset) {
        if (metric_on_validation_fit > reference_metric_validation) fit_after_dropout_improvement = true;
      } else if (metric_on_training_fit > reference_metric_training) {
        fit_after_dropout_improvement = true;
      }
    }
    if (keep_drop && (fit_after_dropout_improvement || random_keep_iter)) {
      dropped_before_cleaning += (size_t) trees_to_dropout;
      metric_on_training = metric_on_training_fit;
      metric_on_validation = metric_on_validation_fit;
    } else {
      // back to the scores before the new tree, normalise, then add the dropped trees and the new one
      update_modelscores_trees(ctx_, false, lastTree);
      if (validation_dataset) update_modelscores_trees(valid_ctx_, false, lastTree);
      if (trees_to_dropout > 0) {
        normalize_trees_restore_drop(orig_weights, dropped_trees, tree_weight);
        ensemble_model_.update_ensemble_weights(orig_weights, false);
      }
      dropped_trees.push_back((int) ensemble_model_.get_size() - 1);
      update_modelscores_trees(ctx_, true, dropped_trees);
      metric_on_training = eval(ctx_);
      if (validation_dataset) {
        update_modelscores_trees(valid_ctx_, true, dropped_trees);
        metric_on_validation = eval(valid_ctx_);
      }
    }
    std::cout << std::setw(7) << m + 1 << std::setw(9) << metric_on_training;
    bool best_improved = false;
    if (validation_dataset && !best_on_train) {
      std::cout << std::setw(9) << metric_on_validation;
      if (metric_on_validation > best_metric_on_validation_) best_improved = true;
    } else if (metric_on_training > best_metric_on_training_) {
      best_improved = true;
    }
    bool best_vali_improved = false;
    if (validation_dataset && best_on_train && metric_on_validation > best_metric_on_validation_) {
      best_vali_improved = true;
      best_metric_on_validation_ = metric_on_validation;
    }
    if (best_improved) {
      best_metric_on_training_ = metric_on_training;
      if (!best_on_train) best_metric_on_validation_ = metric_on_validation;
      best_iter_ = m;
      std::cout << " *";
      // trees whose weight is 0 (kept dropouts) leave the ensemble
      {
        std::vector<std::shared_ptr<DeviceTree>> kept;
        for (size_t t = 0; t < ensemble_model_.get_size(); ++t)
          if (ensemble_model_.getWeight((int) t) != 0) kept.push_back(flat_[t]);
        flat_.swap(kept);
      }
      ensemble_model_.filter_out_zero_weighted_trees();
      best_weights = ensemble_model_.get_weights();
      best_model_ = ensemble_model_.get_size();
      dropped_before_cleaning = 0;
    }
    std::string improved = best_vali_improved ? " *" : "  ";
    std::cout << "\t[ " << metric_on_training_dropout << " - " << metric_on_training_fit << " - " << metric_on_training
              << " | " << metric_on_validation_dropout << (dropout_better_than_full ? " *" : "  ") << " - "
              << metric_on_validation_fit << (fit_after_dropout_improvement ? " *" : "  ") << " - " << metric_on_validation
              << improved << "]";
    std::cout << " \t" << trees_to_dropout << " Dropped Trees - Ensemble size: "
              << ensemble_model_.get_size() - dropped_before_cleaning;
    if (keep_drop && fit_after_dropout_improvement) std::cout << " - Keep Dropout";
    else if (random_keep_iter) std::cout << " - Keep Dropout (RANDOM)";
    else if (trees_to_dropout > 0) std::cout << " - Dropout";
    if (best_improved) {
      std::cout << " - CLEANED";
      if ((m - last_iteration_global_scoring) > 10) {
        rescore_on_device(ctx_, training_dataset);
        if (validation_dataset) rescore_on_device(valid_ctx_, validation_dataset);
        std::cout << " (update)";
        last_iteration_global_scoring = m;
      }
    }
    std::cout << std::endl;
    if (partial_save != 0 && !output_basename.empty() &&
        (ensemble_model_.get_size() - dropped_before_cleaning) % partial_save == 0)
      save(output_basename, (int) (ensemble_model_.get_size() - dropped_before_cleaning));
  }
  if (validation_dataset) {   // roll back to the best model on validation (dart.cc:575-581)
    while (ensemble_model_.is_notempty() && ensemble_model_.get_size() > best_model_) { ensemble_model_.pop(); flat_.pop_back(); }
    ensemble_model_.update_ensemble_weights(best_weights, true);
  }
  double train_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t1).count();
  std::cout << std::endl;
  std::cout << *scorer << " on training data = " << best_metric_on_training_ << std::endl;
  if (validation_dataset) std::cout << *scorer << " on validation data = " << best_metric_on_validation_ << std::endl;
  flat_.clear();
  clear(vertical_training->num_features());
  std::cout << std::endl;
  std::cout << "#\t Training Time: " << std::setprecision(2) << train_time << " s." << std::endl;
}

}  // namespace forests

void LTR_Algorithm::score_dataset(std::shared_ptr<data::Dataset> dataset, Score *scores) const {
  const Feature *d = dataset->data();
  for (size_t i = 0; i < dataset->num_instances(); i++) scores[i] = score_document(d + i * dataset->num_features());
}

void LTR_Algorithm::save(std::string output_basename, int iteration) const {
  if (output_basename.empty()) return;
  std::string filename(output_basename);
  if (iteration != -1) filename += ".T" + std::to_string(iteration) + ".xml";
  std::ofstream f(filename);
  if (!f) {
    std::cerr << "!!! Impossible to write model file " << filename << "." << std::endl;
    exit(EXIT_FAILURE);
  }
  write_xml_model(f);
}

// ltr_algorithm.cc:67-128: dispatch on <info><type>
std::shared_ptr<LTR_Algorithm> LTR_Algorithm::load_model_from_file(std::string model_filename) {
  if (model_filename.empty()) {
    std::cerr << "!!! Model filename is empty." << std::endl;
    exit(EXIT_FAILURE);
  }
  std::ifstream f(model_filename);
  std::stringstream ss;
  ss << f.rdbuf();
  forests::XmlModel model;
  if (!f || !forests::parse_xml(ss.str(), &model) || model.root->name != "ranker") {
    std::cerr << "!!! Model " + model_filename + " is not parsed correctly." << std::endl;
    exit(EXIT_FAILURE);
  }
  const forests::XmlNode *info = model.root->child("info");
  const std::string type = info ? info->child_text("type") : "";
  if (type == forests::Mart::NAME_) return std::shared_ptr<LTR_Algorithm>(new forests::Mart(model));
  if (type == forests::LambdaMart::NAME_) return std::shared_ptr<LTR_Algorithm>(new forests::LambdaMart(model));
  if (type == forests::LambdaMartSelective::NAME_)   // ltr_algorithm.cc:100-102
    return std::shared_ptr<LTR_Algorithm>(new forests::LambdaMartSelective(model));
  if (type == forests::ObliviousMart::NAME_) return std::shared_ptr<LTR_Algorithm>(new forests::ObliviousMart(model));
  if (type == forests::ObliviousLambdaMart::NAME_)
    return std::shared_ptr<LTR_Algorithm>(new forests::ObliviousLambdaMart(model));
  if (type == forests::Dart::NAME_) return std::shared_ptr<LTR_Algorithm>(new forests::Dart(model));
  return nullptr;
}

}  // namespace learning

// ------------------------------------------------------------------------------------------
// C code generators (src/io/generate_conditional_operators.cc, src/io/generate_oblivious.cc)
// ------------------------------------------------------------------------------------------
namespace io {

using learning::forests::XmlModel;
using learning::forests::XmlNode;

static std::unique_ptr<XmlNode> load_ranker(const std::string &model_filename) {
  if (model_filename.empty()) {
    std::cerr << "!!! Model filename is empty." << std::endl;
    exit(EXIT_FAILURE);
  }
  std::ifstream f(model_filename);
  std::stringstream ss;
  ss << f.rdbuf();
  XmlModel model;
  if (!f || !learning::forests::parse_xml(ss.str(), &model) || !model.root || model.root->name != "ranker") {
    std::cerr << "!!! Model " + model_filename + " is not parsed correctly." << std::endl;
    exit(EXIT_FAILURE);
  }
  return std::move(model.root);
}

static const XmlNode *child_at(const XmlNode &n, const char *pos) {
  for (auto &k : n.kids)
    if (k->name == "split") {
      auto it = k->attr.find("pos");
      if (it != k->attr.end() && it->second == pos) return k.get();
    }
  return nullptr;
}
static bool is_leaf(const XmlNode &n) { return n.child("output") != nullptr; }

// generate_conditional_operators.cc:28-76
static void node_to_condop(const XmlNode &n, std::ostream &os) {
  if (is_leaf(n)) { os << n.child_text("output"); return; }
  const unsigned feature_id = (unsigned) strtoul(n.child_text("feature", "0").c_str(), nullptr, 10);
  std::string threshold = n.child_text("threshold");
  if (threshold.find(".") == std::string::npos) threshold += ".0";   // integer-looking values (:50-53)
  const XmlNode *left = child_at(n, "left"), *right = child_at(n, "right");
  os << "( v[" << feature_id - 1 << "] <= " << threshold << "f" << " ? ";
  if (left) node_to_condop(*left, os);
  os << " : ";
  if (right) node_to_condop(*right, os);
  os << " )";
}

void GenOpCond::generate_conditional_operators_code(const std::string model_filename, const std::string code_filename) {
  std::unique_ptr<XmlNode> ranker = load_ranker(model_filename);
  std::stringstream source_code;
  source_code.setf(std::ios::floatfield, std::ios::fixed);
  source_code << "double ranker(float* v) {" << std::endl;
  source_code << "\treturn 0.0 ";
  const XmlNode *ensemble = ranker->child("ensemble");
  if (ensemble)
    for (auto &tree : ensemble->kids) {
      if (tree->name != "tree") continue;
      auto w = tree->attr.find("weight");
      const float tree_weight = w != tree->attr.end() ? (float) strtod(w->second.c_str(), nullptr) : 0.f;
      const XmlNode *content = tree->child("split");
      if (content) {
        source_code << std::endl << "\t\t + " << std::setprecision(3) << tree_weight << "f * ";
        node_to_condop(*content, source_code);
      }
    }
  source_code << ";" << std::endl << "}" << std::endl;
  std::ofstream output(code_filename, std::ofstream::out);
  output << source_code.str();
}

// generate_vpred.cc:43-58: depth of the deepest leaf below a <split> (a node holding <output> counts 1)
static uint32_t vpred_depth(const XmlNode &split) {
  uint32_t ld = 0, rd = 0;
  for (auto &k : split.kids) {
    if (k->name == "output") return 1;
    if (k->name == "split") {
      auto it = k->attr.find("pos");
      if (it != k->attr.end() && it->second == "left") ld = 1 + vpred_depth(*k);
      else rd = 1 + vpred_depth(*k);
    }
  }
  return std::max(ld, rd);
}

// generate_vpred.cc:92-172
void GenVpred::generate_vpred_input(const std::string &ensemble_file, const std::string &output_file) {
  if (ensemble_file.empty()) {
    std::cerr << "!!! Model filename is empty." << std::endl;
    exit(EXIT_FAILURE);
  }
  std::unique_ptr<XmlNode> ranker = load_ranker(ensemble_file);
  std::ofstream output(output_file, std::ofstream::out);
  const XmlNode *info = ranker->child("info");
  const double learning_rate = info ? strtod(info->child_text("shrinkage", "0").c_str(), nullptr) : 0.0;
  std::vector<const XmlNode *> trees;
  if (const XmlNode *ensemble = ranker->child("ensemble"))
    for (auto &t : ensemble->kids)
      if (t->name == "tree") trees.push_back(t.get());
  output << trees.size() << std::endl;
  struct Item { const XmlNode *node; uint32_t id, pid; bool left; std::string parent_feature; };
  for (const XmlNode *tree : trees) {
    const XmlNode *root = tree->child("split");
    if (!root) { output << "end" << std::endl; continue; }
    const uint32_t depth = vpred_depth(*root) - 1;
    output << depth << std::endl;
    const uint32_t tree_size = (uint32_t) std::pow(2, depth) - 1;
    uint32_t local_id = 0;
    std::deque<Item> queue;
    for (queue.push_back(Item{root, local_id++, (uint32_t) -1, false, ""}); !queue.empty(); queue.pop_front()) {
      const Item it = queue.front();
      if (is_leaf(*it.node)) {
        const double out = learning_rate * std::stod(it.node->child_text("output"));
        if (it.id >= tree_size)
          output << "leaf" << " " << it.id << " " << it.pid << " " << it.left << " " << out << std::endl;
        else   // a leaf above the last level: listed as a node with its parent's feature (:137-141)
          output << "node" << " " << it.id << " " << it.pid << " " << (std::stoi(it.parent_feature) - 1) << " "
                 << it.left << " " << out << std::endl;
      } else {
        const std::string feature = it.node->child_text("feature"), theta = it.node->child_text("threshold");
        if (it.id == 0)
          output << "root" << " " << it.id << " " << (std::stoi(feature) - 1) << " " << theta << std::endl;
        else
          output << "node" << " " << it.id << " " << it.pid << " " << (std::stoi(feature) - 1) << " " << it.left
                 << " " << theta << std::endl;
        for (auto &k : it.node->kids)
          if (k->name == "split") {
            auto a = k->attr.find("pos");
            queue.push_back(Item{k.get(), local_id++, it.id, a != k->attr.end() && a->second == "left", feature});
          }
      }
    }
    output << "end" << std::endl;
  }
}

// generate_oblivious.cc:31-135: the three walks (all leaves left to right; features / thresholds down the
// left spine — the trees are symmetric)
static void obv_leaves(const XmlNode &n, std::vector<std::string> &leaves) {
  if (is_leaf(n)) { leaves.push_back(n.child_text("output")); return; }
  if (const XmlNode *l = child_at(n, "left")) obv_leaves(*l, leaves);
  if (const XmlNode *r = child_at(n, "right")) obv_leaves(*r, leaves);
}
static void obv_spine(const XmlNode &n, std::vector<unsigned> &features, std::vector<std::string> &thresholds) {
  if (is_leaf(n)) return;
  features.push_back((unsigned) strtoul(n.child_text("feature", "0").c_str(), nullptr, 10) - 1);
  thresholds.push_back(n.child_text("threshold"));
  if (const XmlNode *l = child_at(n, "left")) obv_spine(*l, features, thresholds);
}

void GenOblivious::generate_oblivious_code(const std::string model_filename, const std::string code_filename) {
  std::unique_ptr<XmlNode> ranker = load_ranker(model_filename);
  std::stringstream source_code;
  source_code.setf(std::ios::floatfield, std::ios::fixed);
  const XmlNode *info = ranker->child("info");
  const unsigned depth = info ? (unsigned) strtoul(info->child_text("depth", "0").c_str(), nullptr, 10) : 0;
  const unsigned max_leaves = 1u << depth;
  const XmlNode *ensemble = ranker->child("ensemble");
  std::vector<float> tree_weights;
  std::vector<int> tree_depths;
  std::vector<std::vector<std::string>> tree_outputs, thresholds;
  std::vector<std::vector<unsigned>> feature_ids;
  if (ensemble)
    for (auto &tree : ensemble->kids) {
      if (tree->name != "tree") continue;
      auto w = tree->attr.find("weight");
      tree_weights.push_back(w != tree->attr.end() ? (float) strtod(w->second.c_str(), nullptr) : 0.f);
      const XmlNode *root = tree->child("split");
      tree_outputs.emplace_back(); feature_ids.emplace_back(); thresholds.emplace_back();
      if (root) {
        obv_leaves(*root, tree_outputs.back());
        obv_spine(*root, feature_ids.back(), thresholds.back());
      }
      // splits along the left spine; the reference's walk (:170-181) never returns less than 1
      tree_depths.push_back(std::max<int>(1, (int) feature_ids.back().size()));
    }
  const int actual_model_size = (int) tree_depths.size();
  if (actual_model_size == 0) {
    std::cerr << "!!! The model has no trees." << std::endl;
    exit(EXIT_FAILURE);
  }
  // trees in order of depth (:210-216; same std::sort call on the same data, hence the same order)
  std::vector<size_t> tree_mapping(tree_depths.size());
  std::iota(tree_mapping.begin(), tree_mapping.end(), 0);
  std::sort(tree_mapping.begin(), tree_mapping.end(), [&tree_depths](int a, int b) { return tree_depths[a] < tree_depths[b]; });
  std::vector<size_t> depths_population;   // number of trees of each depth (:219-233)
  const int max_depth = tree_depths[tree_mapping.back()];
  int curr_depth = 1;
  size_t start_position = 0;
  for (size_t i = 0; i < tree_mapping.size(); i++) {
    while (tree_depths[tree_mapping[i]] > curr_depth) {
      depths_population.push_back(i - start_position);
      curr_depth++;
      start_position = i;
    }
    if (curr_depth == max_depth) break;
  }
  depths_population.push_back(tree_mapping.size() - start_position);

  source_code << "#define N " << actual_model_size << " // no. of trees" << std::endl;
  source_code << "#define M " << depth << " // max tree depth" << std::endl;
  source_code << "#define F " << max_leaves << " // max number of leaves" << std::endl << std::endl;
  source_code << std::setprecision(std::numeric_limits<float>::max_digits10);
  source_code << "const float tree_weights[N] = { ";
  for (size_t i = 0; i < tree_weights.size(); i++) {
    if (i != 0) source_code << ", ";
    source_code << tree_weights[tree_mapping[i]];
  }
  source_code << " };" << std::endl << std::endl;
  auto emit = [&](const char *decl, auto &rows) {
    source_code << decl << std::endl << '\t';
    for (size_t i = 0; i < rows.size(); i++) {
      if (i != 0) source_code << "," << std::endl << '\t';
      source_code << "\t{ ";
      for (size_t j = 0; j < rows[tree_mapping[i]].size(); j++) {
        if (j != 0) source_code << ", ";
        source_code << rows[tree_mapping[i]][j];
      }
      source_code << " }";
    }
    source_code << std::endl << "};" << std::endl << std::endl;
  };
  emit("const double leaf_outputs[N][F] = { ", tree_outputs);
  emit("const unsigned int features_ids[N][M] = { ", feature_ids);
  emit("const float thresholds[N][M] = { ", thresholds);
  source_code << "#define SHL(n,p) ((n)<<(p))" << std::endl << std::endl;
  source_code << "unsigned int leaf_id(float *v, unsigned int const *fids, float const *thresh, const unsigned int m) {"
              << std::endl << "  unsigned int leafidx=0;" << std::endl
              << "  for (unsigned int i=0; i<m; ++i)" << std::endl
              << "    leafidx |= SHL( v[fids[i]]>thresh[i], m-1-i);" << std::endl
              << "  return leafidx;" << std::endl << "}" << std::endl << std::endl;
  source_code << "double ranker(float *v) {" << std::endl << "  double score = 0.0;" << std::endl << "  int i = 0;" << std::endl;
  for (int d = 0; d < max_depth; d++) {
    source_code << "  for (int j = 0; j < " << depths_population[d] << "; ++j) {" << std::endl;
    source_code << "    score += tree_weights[i] * leaf_outputs[i][leaf_id(v, features_ids[i], thresholds[i], " << d + 1
                << ")];" << std::endl;
    source_code << "    i++;" << std::endl;
    source_code << "  }" << std::endl;
  }
  source_code << "  return score;" << std::endl << "}" << std::endl;
  std::ofstream output(code_filename, std::ofstream::out);
  output << source_code.str();
}

}  // namespace io
}  // namespace quickrank

// ---- multi-GPU plumbing of the host layer -----------------------------------------------------
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

namespace quickrank {
namespace host {

static Sharding g_sharding;
void set_sharding(const Sharding &s) { g_sharding = s; }
const Sharding &sharding() { return g_sharding; }

std::vector<std::pair<size_t, size_t>> query_shards(const uint64_t *offsets, size_t nq, int world) {
  std::vector<size_t> bounds{0};
  const uint64_t n = offsets[nq];
  for (int r = 1; r < world; ++r) {
    const double target = (double) (n * (uint64_t) r) / (double) world;
    // first index whose offset is >= target
    size_t q = (size_t) (std::lower_bound(offsets, offsets + nq + 1, target,
                                          [](uint64_t o, double t) { return (double) o < t; }) - offsets);
    if (q > 0 && std::fabs((double) offsets[q - 1] - target) <= std::fabs((double) offsets[std::min(q, nq)] - target)) --q;
    q = std::max(q, bounds.back() + 1);
    q = std::min(q, nq - (size_t) (world - r));
    bounds.push_back(q);
  }
  bounds.push_back(nq);
  std::vector<std::pair<size_t, size_t>> out;
  for (int r = 0; r < world; ++r) out.emplace_back(bounds[r], bounds[r + 1]);
  return out;
}

static bool send_all(int fd, const unsigned char *p, size_t n) {
  while (n) {
    const ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k; n -= (size_t) k;
  }
  return true;
}
static bool recv_all(int fd, unsigned char *p, size_t n) {
  while (n) {
    const ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) return false;
    p += k; n -= (size_t) k;
  }
  return true;
}

namespace {
const char kRendezvousMagic[9] = "QRB2RDV1";
struct RendezvousHello { char magic[8]; int32_t rank, world; };
}  // namespace

bool exchange_bytes(unsigned char *id, size_t nbytes, const Sharding &s, int timeout_s) {
  if (s.world <= 1) return true;
  sockaddr_in a;
  std::memset(&a, 0, sizeof(a));
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t) s.port);
  if (s.rank == 0) {
    const int ls = ::socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) { perror("socket"); return false; }
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    // listen on the rendezvous address itself (127.0.0.1 when quicklearn forked the ranks), not on every interface
    if (inet_pton(AF_INET, s.addr.c_str(), &a.sin_addr) != 1) a.sin_addr.s_addr = htonl(INADDR_ANY);
    if (::bind(ls, (sockaddr *) &a, sizeof(a)) != 0) {
      a.sin_addr.s_addr = htonl(INADDR_ANY);   // (an address of another interface of this host: fall back)
      if (::bind(ls, (sockaddr *) &a, sizeof(a)) != 0) {
        std::cerr << "!!! rank 0 cannot bind port " << s.port << ": " << strerror(errno) << std::endl;
        ::close(ls);
        return false;
      }
    }
    if (::listen(ls, s.world + 8) != 0) {
      std::cerr << "!!! rank 0 cannot listen on port " << s.port << ": " << strerror(errno) << std::endl;
      ::close(ls);
      return false;
    }
    timeval tv{1, 0};
    setsockopt(ls, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    // a peer introduces itself (magic, rank, world); anything else that connects — a port scanner, another job that
    // guessed the same port — is dropped without using up one of the world-1 hand-offs
    std::vector<char> served((size_t) s.world, 0);
    int remaining = s.world - 1;
    const auto t0 = std::chrono::steady_clock::now();
    bool ok = true;
    while (remaining > 0) {
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
        std::cerr << "!!! rank 0: " << remaining << " peer(s) did not connect within " << timeout_s << " s" << std::endl;
        ok = false;
        break;
      }
      const int fd = ::accept(ls, nullptr, nullptr);
      if (fd < 0) continue;   // (accept timed out after 1 s: look at the clock again)
      timeval ptv{5, 0};
      setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &ptv, sizeof(ptv));
      RendezvousHello h{};
      const bool hello = recv_all(fd, (unsigned char *) &h, sizeof(h)) && std::memcmp(h.magic, kRendezvousMagic, 8) == 0 &&
                         h.world == s.world && h.rank >= 1 && h.rank < s.world && !served[(size_t) h.rank];
      if (hello && send_all(fd, id, nbytes)) {
        served[(size_t) h.rank] = 1;
        --remaining;
      }
      ::close(fd);
    }
    ::close(ls);
    return ok;
  }
  if (inet_pton(AF_INET, s.addr.c_str(), &a.sin_addr) != 1) {   // not dotted IPv4: a host name
    addrinfo hints, *res = nullptr;
    std::memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(s.addr.c_str(), nullptr, &hints, &res) != 0 || res == nullptr) {
      std::cerr << "!!! cannot resolve the rendezvous address " << s.addr << std::endl;
      return false;
    }
    a.sin_addr = ((sockaddr_in *) res->ai_addr)->sin_addr;
    freeaddrinfo(res);
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) { perror("socket"); return false; }
    if (::connect(fd, (sockaddr *) &a, sizeof(a)) == 0) {
      timeval tv{timeout_s, 0};
      setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      RendezvousHello h{};
      std::memcpy(h.magic, kRendezvousMagic, 8);
      h.rank = s.rank;
      h.world = s.world;
      const bool ok = send_all(fd, (const unsigned char *) &h, sizeof(h)) && recv_all(fd, id, nbytes);
      ::close(fd);
      if (!ok) std::cerr << "!!! rank " << s.rank << ": the communicator id did not arrive" << std::endl;
      return ok;
    }
    ::close(fd);
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
      std::cerr << "!!! rank " << s.rank << " cannot reach rank 0 at " << s.addr << ":" << s.port << std::endl;
      return false;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
}

}  // namespace host
}  // namespace quickrank
