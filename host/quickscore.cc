// quickscore (B200): the scoring benchmark of the reference (src/quickscore.cc:64-134).  The reference
// links a generated `double ranker(float*)` and calls it per document in a serial loop; here the
// ensemble of an XML model is uploaded once and every round scores the whole dataset in one call.
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "quickrank_host.h"

using namespace quickrank;

int main(int argc, char **argv) {
  std::map<std::string, std::string> opt;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string a = argv[i];
    if (a.rfind("--", 0) == 0) opt[a.substr(2)] = argv[i + 1];
    else if (a == "-d") opt["dataset"] = argv[i + 1];
    else if (a == "-r") opt["rounds"] = argv[i + 1];
    else if (a == "-s") opt["scores"] = argv[i + 1];
    else if (a == "-m") opt["model"] = argv[i + 1];
  }
  if (!opt.count("dataset") || !opt.count("model")) {
    std::cout << "usage: quickscore -d <svml dataset> -m <xml model> [-r rounds (10)] [-s scores file]" << std::endl;
    return EXIT_FAILURE;
  }
  const size_t rounds = opt.count("rounds") ? (size_t) strtoull(opt["rounds"].c_str(), nullptr, 10) : 10;
  std::cout << "# ## ================================== ## #" << std::endl
            << "# ## quickscore on NVIDIA B200 (quickrank_b200)" << std::endl
            << "# ## ================================== ## #" << std::endl;
  io::Svml reader;
  std::shared_ptr<data::Dataset> ds = reader.read_horizontal(opt["dataset"]);
  std::cout << "#\t Dataset size: " << ds->num_instances() << " x " << ds->num_features()
            << " (instances x features)" << std::endl << "#\t Num queries: " << ds->num_queries() << std::endl;
  auto model = learning::LTR_Algorithm::load_model_from_file(opt["model"]);
  if (!model) { std::cerr << "!!! Model type not supported for loading" << std::endl; return EXIT_FAILURE; }
  std::vector<Score> scores(ds->num_instances());
  auto t0 = std::chrono::high_resolution_clock::now();
  for (size_t r = 0; r < rounds; ++r) model->score_dataset(ds, scores.data());
  const double total = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
  std::cout << "       Total scoring time: " << total << " s." << std::endl
            << "Avg. Dataset scoring time: " << total / rounds << " s." << std::endl
            << "Avg.    Doc. scoring time: " << total / rounds / ds->num_instances() << " s." << std::endl;
  if (opt.count("scores")) {
    std::ofstream os(opt["scores"]);
    os << std::setprecision(15);
    for (auto v : scores) os << v << std::endl;
  }
  return EXIT_SUCCESS;
}
