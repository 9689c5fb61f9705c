// Test helper (no GPU needed): the reference's dataset unit test (catch-unit-tests/data/test-hdata.cc:33-105)
// restated as a program.  Reads an SVMLight file with the host reader and prints, one "key value" per line,
// everything that test asserts on: shape, the first labels / features of queries 0 and 1 through
// Dataset::getQueryResults, DCG@3 and NDCG@3 of query 0 for the two score vectors of the test, and the same
// cells through VerticalDataset (column-major layout).
#include <cstdio>
#include <iostream>
#include <vector>

#include "quickrank_host.h"

using namespace quickrank;

int main(int argc, char **argv) {
  if (argc < 2) { std::cerr << "usage: hdata_check file" << std::endl; return 1; }
  io::Svml reader;
  std::shared_ptr<data::Dataset> ds = reader.read_horizontal(argv[1]);
  const size_t F = ds->num_features(), N = ds->num_instances();
  printf("num_features %zu\nnum_instances %zu\nnum_queries %zu\n", F, N, ds->num_queries());
  for (size_t q = 0; q < 2 && q < ds->num_queries(); ++q) {
    auto qr = ds->getQueryResults(q);
    printf("q%zu.num_results %zu\n", q, qr->num_results());
    for (size_t i = 0; i < 3 && i < qr->num_results(); ++i) {
      printf("q%zu.label%zu %.9g\n", q, i, qr->labels()[i]);
      printf("q%zu.feature_%zu_%zu %.9g\n", q, i, i, qr->features()[i * F + i]);   // features()[i * num_features + i]
    }
  }
  auto qr = ds->getQueryResults(0);
  std::vector<Score> s1(qr->num_results(), 0.0), s2(qr->num_results(), 0.0);
  s1[0] = 3; s1[1] = 2; s1[2] = 1;
  s2[0] = 1; s2[1] = 2; s2[2] = 3;
  metric::ir::Dcg dcg(3);
  metric::ir::Ndcg ndcg(3);
  printf("dcg3.a %.17g\ndcg3.b %.17g\n", dcg.evaluate_result_list(qr.get(), s1.data()), dcg.evaluate_result_list(qr.get(), s2.data()));
  printf("ndcg3.a %.17g\nndcg3.b %.17g\n", ndcg.evaluate_result_list(qr.get(), s1.data()), ndcg.evaluate_result_list(qr.get(), s2.data()));
  data::VerticalDataset vd(ds);
  printf("v.num_features %zu\nv.num_instances %zu\nv.num_queries %zu\n", vd.num_features(), vd.num_instances(), vd.num_queries());
  auto vq = vd.getQueryResults(0);
  printf("v.q0.num_results %zu\n", vq->num_results());
  for (size_t i = 0; i < 3 && i < vq->num_results(); ++i)
    printf("v.q0.feature_%zu_%zu %.9g\n", i, i, vq->features()[i * N + i]);       // features()[i * num_instances + i]
  auto vq1 = vd.getQueryResults(1);
  printf("v.q1.feature_2_2 %.9g\n", vq1->features()[2 * N + 2]);
  printf("v.at_5_1 %.9g\n", *vd.at(5, 1));
  return 0;
}
