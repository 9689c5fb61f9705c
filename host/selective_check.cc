// Development / test tool: one draw of LambdaMartSelective::sampling_query_level on the host, no GPU involved.
//   selective_check <rank_factor> <random_factor> <adaptive> <negative> <adapt_factor> [rand_calls_to_skip] < input
// input (text): Q N, then Q+1 query offsets, then N lines "label score".  srand(0) before the draw.
// rand_calls_to_skip: rand() calls earlier draws of the same training run have consumed since srand(0).
// output: the size of the sample, then the N ids of the permuted list; stderr: the draw's log, its duration and the
// number of rand() calls it made.  tests/test_sampled_trainers.py compares it
// with the unmodified reference's function on the same input.
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <vector>

#include "quickrank_host.h"

using namespace quickrank;

int main(int argc, char **argv) {
  if (argc < 6) { std::cerr << "usage: selective_check rank_factor random_factor adaptive negative adapt_factor" << std::endl; return 2; }
  size_t Q = 0, N = 0;
  std::cin >> Q >> N;
  std::vector<uint64_t> off(Q + 1);
  for (auto &o : off) std::cin >> o;
  std::vector<Label> labels(N);
  std::vector<Score> scores(N);
  for (size_t i = 0; i < N; ++i) std::cin >> labels[i] >> scores[i];
  if (!std::cin || off[Q] != N) { std::cerr << "bad input" << std::endl; return 2; }
  data::Dataset ds(N, 1);
  ds.set_structure(labels.data(), off);
  learning::forests::LambdaMartSelective algo(1, 0.1, 0, 4, 1, 1.0f, 1.0f, 0, 0.0f, 1, strtof(argv[1], nullptr),
                                              strtof(argv[2], nullptr), 100.0f, argv[3], argv[4]);
  std::vector<size_t> npos(Q, 0), ids(N);
  for (size_t q = 0; q < Q; ++q)
    for (size_t d = off[q]; d < off[q + 1]; ++d) npos[q] += labels[d] > 0;
  for (size_t i = 0; i < N; ++i) ids[i] = i;
  std::ostringstream log;
  std::streambuf *keep = std::cout.rdbuf(log.rdbuf());
  srand(0);
  for (long k = argc > 6 ? atol(argv[6]) : 0; k > 0; --k) (void) rand();
  const auto t0 = std::chrono::steady_clock::now();
  const size_t n = algo.sampling_query_level(ds, scores, npos, ids, strtof(argv[5], nullptr));
  const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::cout.rdbuf(keep);
  std::cerr << "draw over " << N << " documents: " << ms << " ms" << std::endl;
  std::cerr << "rand calls: " << learning::forests::LambdaMartSelective::rand_calls() << std::endl;
  std::cout << n << "\n";
  for (size_t i = 0; i < N; ++i) std::cout << ids[i] << "\n";
  std::cerr << log.str();
  return 0;
}
