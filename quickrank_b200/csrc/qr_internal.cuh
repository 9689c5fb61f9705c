// quickrank_b200 — internal declarations shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/quickrank_b200.h"
#include "qr_task.cuh"

namespace qr {

void set_error(const char *fmt, ...);

#define QR_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      qr::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                    cudaGetErrorString(_e));                                            \
      return QR_ECUDA;                                                                  \
    }                                                                                   \
  } while (0)

#define QR_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != QR_OK) return _rc; \
  } while (0)

constexpr int kPanelBytes = 16;   // one 128-bit load = one document's bins for one panel
constexpr int kNumPhases = 6;
enum Phase { PH_PSEUDO = 0, PH_HIST = 1, PH_SCAN = 2, PH_PARTITION = 3, PH_LEAF = 4, PH_RANK = 5 };

// Best split of one node, written by the scan kernels and read back by the host
// (replaces the tail of RegressionTree::split, rt.cc:297-318 of the reference).
struct SplitResult {
  double score;       // best lsum^2/lcnt + rsum^2/rcnt, -1 if none
  double sum;         // node sum of pseudo-responses (feature 0, last bin)
  double squares;     // squares_sum_
  double deviance;    // squares - sum^2/n
  uint64_t n;         // node size (feature 0, last bin); global when several ranks train together
  uint64_t lcount;    // left size at the best split
  uint32_t feature;
  uint32_t threshold_idx;
  uint32_t valid;
  uint32_t pad;
};

struct HostNode {
  uint32_t lo = 0, n = 0;   // REFERENCE mode: segment [lo, lo+n) of the id buffer `buf`
  int buf = 0;              // REFERENCE mode: 0 / 1: id buffer; 2: identity list (root)
  int parent = -1;          // FAST mode: the node's index is the id its documents carry in node_of_doc
  int hist = -1;            // histogram slot, -1 once released
  int left = -1, right = -1;
  bool expanded = false;    // children (and their split scans) have been computed
  bool pushed = false;      // the heap replay split this node (it is an internal node of the tree)
  SplitResult res{};
  double value = 0.0;       // leaf output
  bool is_leaf() const { return left < 0; }
};

struct Comm;      // NCCL plumbing (qr_comm.cu)
struct ChildOut;
struct DevCand;

}  // namespace qr

struct qr_ctx {
  qr_params p{};
  int device = 0;
  size_t N = 0, F = 0, Q = 0;
  size_t cutoff = 0;             // SIZE_MAX when "no cutoff"
  bool lambda = false, oblivious = false, exact = false;
  bool eval_only = false;        // validation / test set binned with another context's thresholds
  cudaStream_t stream = nullptr;

  // binning (host copies for XML / split thresholds)
  std::vector<std::vector<float>> thr;
  std::vector<uint32_t> thr_off;  // prefix over features, size F+1
  uint32_t ncells = 0;
  uint32_t max_thr = 0;
  int bin_bytes = 1;              // 1: u8 bins, 2: u16 bins
  int fpp = 16;                   // features per 16-byte panel
  uint32_t npanels = 0;
  uint32_t max_panel_cells = 0;

  // device state
  uint4 *d_panels = nullptr;      // [npanels][N]
  uint4 *d_rows = nullptr;        // [N][npanels] document-major copy for gathered histogram launches (FAST mode)
  uint32_t *d_thr_off = nullptr;  // [F+1]
  float *d_labels = nullptr;      // [N]
  double *d_gain = nullptr;       // [N] pow(2, label)
  uint32_t *d_qoff = nullptr;     // [Q+1]
  double *d_idcg = nullptr;       // [Q]
  double *d_invlg = nullptr;      // [maxlen] 1/log2(i+2)
  double *d_lg = nullptr;         // [maxlen] log2((float)i+2)
  double *d_scores = nullptr, *d_lambda = nullptr, *d_weight = nullptr;  // [N]
  long long *d_lamq = nullptr;    // [N] fixed-point pseudo-responses (FAST mode)
  // FAST mode (qr_fast_kernels.cuh): a node is the set of documents carrying its id
  uint16_t *d_node = nullptr;     // [N padded to 4] node_of_doc
  uint32_t *d_cids = nullptr;     // [compact_cap] compact lists of the round's built children: document ids ...
  long long *d_clamq = nullptr;   // [compact_cap] ... and their fixed-point pseudo-responses
  size_t compact_cap = 0;
  uint32_t *d_counts = nullptr;   // [2][max_tasks] entries appended to each task's list (double-buffered by round)
  uint32_t count_parity = 0;
  uint32_t max_nodes = 0;         // node ids are 16 bits
  // one GPU: the split scan hands per-feature candidates to the host, which reduces over features
  bool fused_scan = false;
  bool pub_ok = false;              // scan_pub_kernel usable (fixed-point mode, F <= 65535): one GPU, and sharded over peer memory
  qr::ChildOut *h_out = nullptr, *d_out_mapped = nullptr;           // [max_tasks][2] mapped pinned: the round's results
  qr::DevCand *d_cand = nullptr;                                    // [max_tasks][2][F] per-feature winners
  double2 *d_noderec = nullptr;                                     // [max_tasks][2]
  ulonglong2 *d_sq_acc = nullptr;                                   // [max_tasks]
  bool maxabs_valid = false;      // d_maxabs holds max |lambda| of the current pseudo-responses (written by their kernel)
  unsigned long long *d_maxabs = nullptr;  // bits of max |lambda|
  int *d_qexp = nullptr;          // fixed-point exponent chosen for this tree
  uint32_t *d_rankpos = nullptr;  // [N] position (within its query) of the doc at each rank
  // document sample of another context (qr_ctx_create_sample): where each document's score and its ranking key
  // live in the sampled context's score array, and the gathered ranking keys (nullptr: rank by d_scores)
  uint32_t *d_src_doc = nullptr, *d_key_doc = nullptr;   // [N]
  double *d_rankkey = nullptr;                           // [N]
  size_t sample_of_N = 0;                                // documents of the sampled context (0: not a sample)
  // buffers that scale with the documents / queries / longest query are allocated for these counts (>= N, Q, maxlen):
  // a sample context is sized for the whole sampled set, so that qr_sample_redraw can refill it in place
  size_t cap_N = 0, cap_Q = 0;
  uint32_t cap_maxlen = 0;
  double *d_qndcg = nullptr;      // [Q]
  double *d_metric = nullptr;     // [1]
  double *d_vec_qndcg = nullptr, *d_vec_metric = nullptr;   // evaluate_vectors: [vec_cap][Q] per-query values, [vec_cap] means
  uint32_t vec_cap = 0;
  uint32_t *d_ids[2] = {nullptr, nullptr};  // [N] node document lists (ping-pong)
  uint32_t *d_leaf_of_doc = nullptr;        // [N]
  uint32_t *d_blockcnt = nullptr;           // partition scratch (REFERENCE mode)
  double *d_partials = nullptr;             // REFERENCE: squares per task [max_tasks]
  // REFERENCE mode, large nodes (qr_exact_kernels.cuh): per feature the documents sorted by (bin, document),
  // where each histogram cell's documents start, this round's node marks, chunk records of the ordered squares sums
  uint32_t *d_perm = nullptr;               // [F][N]
  unsigned long long *d_cell_pos = nullptr; // [ncells + 1]
  uint32_t *d_mark = nullptr;               // [N]
  uint8_t *d_fskip = nullptr;               // [F] 1: single-bin feature, not accumulated (nullptr: none)
  void *d_sq_chunks = nullptr;              // [N / kSqChunk + max_tasks] SqChunk
  uint32_t walk_min = 0;                    // built children of at least this many documents take the walk
  uint32_t mark_tag = 0;
  unsigned long long *d_sq_replayed = nullptr;   // chunks of the ordered squares sums replayed addition by addition
  ulonglong2 *d_sq128 = nullptr;            // FAST: exact squares per histogram slice [max_slices]
  uint32_t max_slices = 0;
  uint32_t *d_task_done = nullptr;          // [max_tasks] finalize completion counters
  double *d_sq_built = nullptr;             // [max_tasks] squares sum of each task's built child (split scan scratch)
  uint32_t *d_root_cnt = nullptr;           // [ncells] per-bin document counts of the whole dataset
  unsigned long long *d_hist_sum = nullptr; // [nslots][ncells] int64 (FAST) or double (REFERENCE)
  uint32_t *d_hist_cnt = nullptr;           // [nslots][ncells]
  int nslots = 0;
  std::vector<int> free_slots;
  uint32_t max_tasks = 0;                   // node expansions per round
  qr::NodeTask *d_tasks = nullptr, *h_tasks = nullptr;   // [max_tasks] (host copy pinned)
  qr::TaskPack pack{};                      // task records of a small round, passed as kernel parameters
  uint32_t *d_lcount = nullptr;             // [max_tasks] left counts (REFERENCE mode partition)
  double *d_fbest_score = nullptr;          // [max_tasks][2][F] per-feature winners of the split scan
  uint32_t *d_fbest_t = nullptr;            // [max_tasks][2][F]
  uint32_t *d_fbest_lc = nullptr;           // [max_tasks][2][F] left count at each feature's best split
  ulonglong2 *d_totals = nullptr;           // [max_tasks][2] (node size, node sum bits)
  qr::SplitResult *d_res = nullptr;         // [max_tasks][2] (oblivious level arg-max)
  qr::SplitResult *h_res = nullptr;         // pinned + mapped: scan_kernel writes it directly
  qr::SplitResult *d_res_mapped = nullptr;  // device view of h_res
  uint32_t *h_flags = nullptr, *d_flags_mapped = nullptr;   // [max_tasks] per-task "result published" round ids
  uint32_t round_id = 0;
  uint32_t *h_err = nullptr, *d_err_mapped = nullptr;       // [1] set by a kernel whose wait for a peer rank timed out
  qr::LeafSeg *d_segs = nullptr, *h_segs = nullptr;      // [maxleaves] (REFERENCE mode)
  double2 *d_leaf_partials = nullptr;       // [blocks][leaves] per-block leaf sums (FAST mode)
  size_t leaf_part_cap = 0, leaf_smem_set = 0;
  unsigned char *h_leafmeta = nullptr, *d_leafmeta = nullptr;   // node -> leaf table (u16) | leaf sizes (u64 at leafn_off)
  size_t leafn_off = 0;
  double2 *d_leafsum = nullptr;             // [maxleaves] (sum lambda, sum weight)
  double *d_leafval = nullptr;              // [maxleaves]
  double *h_leafval = nullptr;              // pinned
  double *d_obv_scores = nullptr;           // [ncells] oblivious level sums
  int *d_obv_slots = nullptr;               // [max_tasks]
  uint64_t *d_obv_lcounts = nullptr, *h_obv_lcounts = nullptr;  // [max_tasks]
  uint32_t maxlen = 0;                      // longest query

  bool ranking_valid = false;  // d_rankpos/d_qndcg match d_scores
  bool has_tree = false;
  std::vector<qr::HostNode> nodes;  // last fitted tree
  std::vector<int> leaves;          // node ids in DFS order
  double rho = 0, sigma = 0, beta = 0;
  uint32_t nsplits = 0, nrounds = 0;

  uint64_t launches = 0;
  bool profiling = false;
  double phase_ms[qr::kNumPhases] = {0};
  uint64_t phase_launches[qr::kNumPhases] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // around each histogram-kernel launch while profiling
  unsigned long long *d_kspan = nullptr;          // [2] first block start / last block end of the profiled launch (ns)
  double histk_ms = 0;
  uint64_t histk_launches = 0;
  double histk_docs = 0;                          // documents accumulated by those launches

  void *d_apply = nullptr, *h_apply = nullptr;   // staging of qr_apply_trees (device / pinned host)
  size_t apply_cap = 0;

  qr::Comm *comm = nullptr;
  size_t N_global = 0, Q_global = 0;
  size_t N_local_max = 0;                   // largest shard (documents) over the ranks
  // sharded training over peer memory (qr_comm.cu): per-round state of the histogram exchange
  int stage_slot0 = 0;                      // first of the 2 x max_tasks staging slots (after the nslots pool slots)
  uint32_t xround = 0;                      // exchange rounds so far (its parity selects the staging set)
  bool round_fused = false;                 // this round's all-reduce happens inside scan_kernel
  uint32_t round_parity = 0, round_sq_off = 0;   // staging set / offset into d_sq128 of this round
  bool peer_fused = true;                   // QR_PEER_FUSED=0: always the stand-alone exchange kernel
  bool sliced = false;                      // leaf-wise growth over peer memory: every rank adds up and scans F / world
                                            // features, the ranks exchange their winners (QR_PEER_SLICED=0: every rank scans all)
  uint32_t oneshot_max = 12;                // fuse while (world - 1) * tasks * histogram bytes <= this many MB (QR_PEER_ONESHOT_MAX)
};

namespace qr {
// NDCG@k of several score vectors over a REFERENCE-mode context's documents in one ranking launch (qr_train.cu)
int evaluate_vectors(qr_ctx *c, const double *scores_dev, uint32_t nvec, double *metrics_host);
}  // namespace qr
