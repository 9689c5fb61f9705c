// Runs the reference's own Mart::learn (mart.cc:208) on a GpuLambdaMart (gpulambdamart.h) and saves the model with
// the reference's own XML writer: usage  gpu_lambdamart_demo <train.svml> <model.xml> <trees> <leaves> [reference]
#include <cstring>
#include <iostream>

#include "gpulambdamart.h"
#include "io/svml.h"

using namespace quickrank;

int main(int argc, char **argv) {
  if (argc < 5) { std::cerr << "usage: " << argv[0] << " train.svml model.xml trees leaves [reference]" << std::endl; return 2; }
  io::Svml reader;
  std::shared_ptr<data::Dataset> train = reader.read_horizontal(argv[1]);
  const size_t ntrees = (size_t) atol(argv[3]), nleaves = (size_t) atol(argv[4]);
  // same constructor as LambdaMart (mart.h:52-66): trees, shrinkage, thresholds, leaves, minls, subsample,
  // max_features, esr, collapse_leaves_factor
  auto algo = std::make_shared<learning::forests::GpuLambdaMart>(ntrees, 0.1, 0, nleaves, 1, 1.0f, 1.0f, 100, 0.0f);
  if (argc > 5 && !std::strcmp(argv[5], "reference")) algo->set_hist_mode(QR_HIST_REFERENCE);
  std::shared_ptr<metric::ir::Metric> scorer(new metric::ir::GpuNdcg(10, algo.get()));
  algo->learn(train, nullptr, scorer, 0, std::string());
  algo->save(argv[2]);
  return 0;
}
