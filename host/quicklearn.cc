// quicklearn (B200): the command-line front end of the reference (src/quicklearn.cc, src/driver/driver.cc)
// for the algorithms whose hot path runs on the GPU.  Same option names and defaults
// (quicklearn.cc:97-140): LAMBDAMART, 1000 trees, shrinkage 0.1, --num-thresholds 0 (= one per distinct
// value), min leaf support 1, end-after-rounds 100, 10 leaves, depth 3, NDCG@10, partial save 100.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <string>

#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

#include <sstream>
#include <vector>

#include "quickrank_host.h"

using namespace quickrank;

// Multi-GPU training: `--gpus N` forks N-1 more processes (one per GPU, before any CUDA call); when an
// external launcher has set RANK / WORLD_SIZE (= N) / LOCAL_RANK (torchrun, mpirun wrappers) its processes
// are taken as they come instead.  Rank 0 serves the NCCL communicator id on MASTER_ADDR : QR_COMM_PORT.  Every rank
// runs the same training loop; only rank 0 prints and writes files.
static std::vector<pid_t> g_children;
// rank 0 is going away without having waited for the others (an error path): stop them
static void reap_children() {
  for (pid_t k : g_children) {
    int st = 0;
    if (waitpid(k, &st, WNOHANG) == 0) { kill(k, SIGTERM); waitpid(k, &st, 0); }
  }
  g_children.clear();
}
static int setup_sharding(int gpus) {
  host::Sharding sh;
  const char *er = getenv("RANK"), *ew = getenv("WORLD_SIZE");
  if (gpus <= 1) return 0;   // sharding is opt-in: an inherited RANK / WORLD_SIZE alone changes nothing
  if (er && ew && atoi(ew) > 1) {
    if (atoi(ew) != gpus) {
      std::cerr << "!!! --gpus " << gpus << " but the launcher's WORLD_SIZE is " << atoi(ew) << "." << std::endl;
      exit(EXIT_FAILURE);
    }
    sh.rank = atoi(er);
    sh.world = atoi(ew);
    sh.local_rank = getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : sh.rank;
    if (getenv("MASTER_ADDR")) sh.addr = getenv("MASTER_ADDR");
    sh.port = getenv("QR_COMM_PORT") ? atoi(getenv("QR_COMM_PORT")) : (getenv("MASTER_PORT") ? atoi(getenv("MASTER_PORT")) + 17 : 29517);
  } else {
    sh.world = gpus;
    sh.port = getenv("QR_COMM_PORT") ? atoi(getenv("QR_COMM_PORT")) : 20000 + (int) (getpid() % 20000);
    std::cout.flush();
    for (int r = 1; r < gpus; ++r) {
      const pid_t pid = fork();
      if (pid < 0) { perror("fork"); exit(EXIT_FAILURE); }
      if (pid == 0) { sh.rank = r; g_children.clear(); break; }
      g_children.push_back(pid);
    }
    if (sh.rank == 0) atexit(reap_children);   // exit() paths (die(), early returns) must not leave the peers waiting
    sh.local_rank = sh.rank;
  }
  if (sh.rank < 0 || sh.rank >= sh.world) { std::cerr << "!!! Bad RANK / WORLD_SIZE" << std::endl; exit(EXIT_FAILURE); }
  host::set_sharding(sh);
  return sh.rank;
}
// rank 0 of a self-launched run: the exit status covers the other ranks
static int finish(int rc) {
  if (rc != EXIT_SUCCESS) { reap_children(); return rc; }
  for (pid_t k : g_children) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { std::cerr << "!!! A training process failed." << std::endl; rc = EXIT_FAILURE; }
  }
  g_children.clear();
  return rc;
}

static void usage() {
  std::cout << "quicklearn (quickrank_b200): LambdaMART / MART / oblivious variants on NVIDIA B200\n"
               "  --algo <MART|LAMBDAMART|OBVMART|OBVLAMBDAMART|DART|LAMBDAMART-SELECTIVE|STOCHASTIC-NEGATIVE>   (default LAMBDAMART)\n"
               "  --train <svml file> [--valid <svml file>] [--test <svml file>]\n"
               "  --model-out <xml> | --model-in <xml>  [--restart-train]\n"
               "  --num-trees N (1000)  --shrinkage X (0.1)  --num-thresholds N (0 = unlimited)\n"
               "  --min-leaf-support N (1)  --end-after-rounds N (100)  --num-leaves N (10)  --tree-depth N (3)\n"
               "  --train-metric NDCG  --train-cutoff K (10)  --test-metric NDCG  --test-cutoff K (10)\n"
               "  --partial N (100)  --scores <file>\n"
               "  LAMBDAMART-SELECTIVE: --sampling-iterations N  --rank-sampling-factor X (1.0)  --random-sampling-factor X (0)\n"
               "                        --normalization-factor X (100)  --adaptive-strategy <NO|FIXED|RATIO|MIX>  --negative-strategy <RATIO|MUL|POS>\n"
               "  STOCHASTIC-NEGATIVE:  --subsample X (share, or count if > 1, of each query's negatives)  --seed N (0)\n"
               "  --hist-mode <fast|reference> (fast)  --device N\n"
               "  --gpus N   train on N GPUs (documents sharded by query, one process per GPU: quicklearn forks them,\n"
               "             or takes the N processes of a launcher that sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR)\n";
}

int main(int argc, char **argv) {
  std::map<std::string, std::string> opt;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "-h" || a == "--help") { usage(); return EXIT_SUCCESS; }
    if (a.rfind("--", 0) != 0) { std::cerr << "!!! Unexpected argument " << a << std::endl; return EXIT_FAILURE; }
    a = a.substr(2);
    if (a == "restart-train" || a == "keep-drop" || a == "best-on-train" || a == "drop-on-best") { opt[a] = "1"; continue; }
    if (i + 1 >= argc) { std::cerr << "!!! Option --" << a << " needs a value" << std::endl; return EXIT_FAILURE; }
    opt[a] = argv[++i];
  }
  auto get = [&](const char *k, const char *def) { auto it = opt.find(k); return it == opt.end() ? std::string(def) : it->second; };
  auto geti = [&](const char *k, size_t def) { auto it = opt.find(k); return it == opt.end() ? def : (size_t) strtoull(it->second.c_str(), nullptr, 10); };

  // code generation (driver.cc:199-223): --model-file m.xml --code-file ranker.cc [--generator condop|oblivious|vpred]
  if (opt.count("model-file") && opt.count("code-file")) {
    const std::string gen = get("generator", "condop");
    std::cout << "# Generating code (" << gen << ") from " << opt["model-file"] << " into " << opt["code-file"] << std::endl;
    if (gen == "condop") io::GenOpCond().generate_conditional_operators_code(opt["model-file"], opt["code-file"]);
    else if (gen == "oblivious") io::GenOblivious().generate_oblivious_code(opt["model-file"], opt["code-file"]);
    else if (gen == "vpred") io::GenVpred().generate_vpred_input(opt["model-file"], opt["code-file"]);
    else { std::cerr << "!!! Generator " << gen << " is not supported (condop, oblivious, vpred)." << std::endl; return EXIT_FAILURE; }
    return EXIT_SUCCESS;
  }
  const std::string algo = get("algo", "LAMBDAMART");
  int rank = 0;
  auto load = [](const std::string &file, std::ostream &log) {
    io::Svml reader;
    auto t0 = std::chrono::high_resolution_clock::now();
    std::shared_ptr<data::Dataset> ds = reader.read_horizontal(file);
    double s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    log << "#\t Reading time: " << std::setprecision(2) << s << " s." << std::endl
        << "#\t Dataset size: " << ds->num_instances() << " x " << ds->num_features()
        << " (instances x features)" << std::endl
        << "#\t Num queries: " << ds->num_queries() << std::endl;
    return ds;
  };
  // `--gpus N` without an external launcher forks the ranks: the datasets are parsed ONCE, here, before the fork (no
  // CUDA call has been made yet), and the ranks share the parsed pages copy-on-write instead of each parsing the text
  // with every core and holding a private copy.  What the reader reports is printed where the reference prints it.
  std::shared_ptr<data::Dataset> preloaded_train, preloaded_valid;
  std::ostringstream preload_train_log, preload_valid_log;
  if (opt.count("train") && (!opt.count("model-in") || opt.count("restart-train"))) {
    const int gpus = (int) geti("gpus", 1);
    const bool launcher = getenv("RANK") && getenv("WORLD_SIZE") && atoi(getenv("WORLD_SIZE")) > 1;
    if (gpus > 1 && !launcher) {
      preloaded_train = load(opt["train"], preload_train_log);
      if (opt.count("valid")) preloaded_valid = load(opt["valid"], preload_valid_log);
    }
    rank = setup_sharding(gpus);
    if (rank != 0) {   // same loop, no report
      // (never destroyed: std::cout is flushed once more when the process exits)
      std::cout.rdbuf((new std::ostringstream())->rdbuf());
      opt.erase("model-out");
      opt.erase("scores");
    }
  }
  const size_t ntrees = geti("num-trees", 1000), nthr = geti("num-thresholds", 0), minls = geti("min-leaf-support", 1);
  const size_t esr = geti("end-after-rounds", 100), nleaves = geti("num-leaves", 10), depth = geti("tree-depth", 3);
  const double shrinkage = strtod(get("shrinkage", "0.1").c_str(), nullptr);
  const size_t partial = geti("partial", 100);

  // ltr_algorithm_factory.cc:45-259: --model-in alone = the loaded model, no training; with --restart-train the
  // algorithm is built from the command line and takes over the loaded trees if the two are compatible
  std::shared_ptr<learning::LTR_Algorithm> ranker, loaded;
  learning::forests::Mart *mart = nullptr;
  const bool from_scratch = !opt.count("model-in") || opt.count("restart-train");
  if (opt.count("model-in")) {
    std::cout << "# Loading model from file " << opt["model-in"] << std::endl;
    loaded = learning::LTR_Algorithm::load_model_from_file(opt["model-in"]);
    if (!loaded) std::cerr << " !! Unable to load model from file." << std::endl;
  }
  if (!from_scratch) {
    if (!loaded) return finish(EXIT_FAILURE);
    ranker = loaded;
  } else if (algo == "MART") {
    ranker.reset(mart = new learning::forests::Mart(ntrees, shrinkage, nthr, nleaves, minls, 1.0f, 1.0f, esr, 0.0f));
  } else if (algo == "LAMBDAMART") {
    ranker.reset(mart = new learning::forests::LambdaMart(ntrees, shrinkage, nthr, nleaves, minls, 1.0f, 1.0f, esr, 0.0f));
  } else if (algo == "LAMBDAMART-SELECTIVE") {   // ltr_algorithm_factory.cc:80-98, defaults of quicklearn.cc:114-119
    ranker.reset(mart = new learning::forests::LambdaMartSelective(
                     ntrees, shrinkage, nthr, nleaves, minls, 1.0f, 1.0f, esr, 0.0f, atoi(get("sampling-iterations", "0").c_str()),
                     strtof(get("rank-sampling-factor", "1.0").c_str(), nullptr),
                     strtof(get("random-sampling-factor", "0.0").c_str(), nullptr),
                     strtof(get("normalization-factor", "100").c_str(), nullptr), get("adaptive-strategy", "NO"),
                     get("negative-strategy", "RATIO")));
  } else if (algo == "STOCHASTIC-NEGATIVE") {    // ltr_algorithm_factory.cc:99-111
    auto *sn = new learning::forests::StochasticNegative(ntrees, shrinkage, nthr, nleaves, minls,
                                                         strtof(get("subsample", "1.0").c_str(), nullptr), 1.0f, esr, 0.0f);
    sn->set_seed(strtoull(get("seed", "0").c_str(), nullptr, 10));
    ranker.reset(mart = sn);
  } else if (algo == "OBVMART") {
    ranker.reset(mart = new learning::forests::ObliviousMart(ntrees, shrinkage, nthr, depth, minls, 1.0f, 1.0f, esr, 0.0f));
  } else if (algo == "OBVLAMBDAMART") {
    ranker.reset(mart = new learning::forests::ObliviousLambdaMart(ntrees, shrinkage, nthr, depth, minls, 1.0f, 1.0f, esr, 0.0f));
  } else if (algo == "DART") {
    using learning::forests::Dart;
    ranker.reset(mart = new Dart(ntrees, shrinkage, nthr, nleaves, minls, 1.0f, 1.0f, esr, 0.0f,
                                 Dart::get_sampling_type(get("sample-type", "UNIFORM")),
                                 Dart::get_normalization_type(get("normalize-type", "TREE")),
                                 Dart::get_adaptive_type(get("adaptive-type", "FIXED")),
                                 strtod(get("rate-drop", "0.1").c_str(), nullptr),
                                 strtod(get("skip-drop", "0").c_str(), nullptr), opt.count("keep-drop") != 0,
                                 opt.count("best-on-train") != 0, strtod(get("random-keep", "0").c_str(), nullptr),
                                 opt.count("drop-on-best") ? 1.0 : 0.0));
  } else {
    std::cerr << "!!! Algorithm " << algo << " is not accelerated by this build (see DESIGN.md, out of scope)." << std::endl;
    return finish(EXIT_FAILURE);
  }
  if (loaded && from_scratch && !ranker->import_model_state(*loaded)) {
    std::cerr << " !! Models not compatible for restart!" << std::endl;
    return finish(EXIT_FAILURE);
  }
  if (!mart) mart = dynamic_cast<learning::forests::Mart *>(ranker.get());
  if (mart) {
    mart->set_hist_mode(get("hist-mode", "fast") == "reference" ? QR_HIST_REFERENCE : QR_HIST_FAST);
    if (opt.count("device")) mart->set_device(atoi(opt["device"].c_str()));
  }
  std::cout << "#" << std::endl << *ranker << "#" << std::endl;
  if (host::sharding().world > 1)
    std::cout << "# training on " << host::sharding().world << " GPUs (documents sharded by query)" << std::endl << "#" << std::endl;

  if (get("train-metric", "NDCG") != "NDCG" || get("test-metric", "NDCG") != "NDCG") {
    std::cerr << "!!! Only NDCG is supported by the GPU engine." << std::endl;
    return finish(EXIT_FAILURE);
  }
  if (opt.count("train") && from_scratch) {   // driver.cc:136-138: a model loaded without --restart-train is not trained
    std::shared_ptr<metric::ir::Metric> train_metric(new metric::ir::Ndcg(geti("train-cutoff", 10)));
    std::cout << "# Reading training dataset: " << opt["train"] << std::endl;
    std::shared_ptr<data::Dataset> train = preloaded_train;
    if (train) std::cout << preload_train_log.str();
    else train = load(opt["train"], std::cout);
    std::shared_ptr<data::Dataset> valid;
    if (opt.count("valid")) {
      std::cout << "# Reading validation dataset: " << opt["valid"] << std::endl;
      valid = preloaded_valid;
      if (valid) std::cout << preload_valid_log.str();
      else valid = load(opt["valid"], std::cout);
    }
    std::cout << "#" << std::endl << "# training scorer: " << *train_metric << std::endl;
    ranker->learn(train, valid, train_metric, rank == 0 ? partial : 0, get("model-out", ""));
    if (opt.count("model-out")) {
      std::cout << "# Writing model to file: " << opt["model-out"] << std::endl;
      ranker->save(opt["model-out"]);
    }
  }
  if (opt.count("test") && rank == 0) {   // (the other ranks of a multi-GPU run have nothing to report)
    std::shared_ptr<metric::ir::Metric> test_metric(new metric::ir::Ndcg(geti("test-cutoff", 10)));
    std::cout << "# Reading test dataset: " << opt["test"] << std::endl;
    auto test = load(opt["test"], std::cout);
    std::vector<Score> scores(test->num_instances());
    ranker->score_dataset(test, scores.data());
    const MetricScore m = test_metric->evaluate_dataset(test, scores.data());
    std::cout << std::endl << *test_metric << " on test data = " << std::setprecision(4) << m << std::endl << std::endl;
    if (opt.count("scores")) {
      std::ofstream os(opt["scores"]);
      os << std::setprecision(15);
      for (size_t i = 0; i < test->num_instances(); ++i) os << scores[i] << std::endl;
      std::cout << "# Scores written to file: " << opt["scores"] << std::endl;
    }
  }
  return finish(EXIT_SUCCESS);
}
