// quickrank_b200 — host side of tree growth (included by qr_train.cu inside namespace qr).
//
// The reference grows a tree one node at a time (RegressionTree::fit, rt.cc:49-163): pop the
// frontier node with the largest deviance from a max-heap, split it, push its children.  Here the
// SAME sequence of pops and pushes is replayed on the host with a replica of the reference's heap
// (maxheap.h:31-106), but the expensive part of a split — partition, child histograms, split scan
// of both children — is done for several frontier nodes per kernel launch ("rounds").  That is
// legal because what a split produces depends only on the node, not on when it is split: the heap
// order merely decides WHICH nodes end up being split within the leaf budget.  A round expands the
// node the replay is blocked on plus the other frontier nodes that can still be reached with the
// remaining budget; expansions that the replay never reaches are simply dropped.
#pragma once

template <typename Fn>
static int dispatch_bins(const qr_ctx *c, Fn &&fn) {
  return c->bin_bytes == 1 ? fn(uint8_t()) : fn(uint16_t());
}

static int alloc_slot(qr_ctx *c) {
  if (c->free_slots.empty()) return -1;
  int s = c->free_slots.back();
  c->free_slots.pop_back();
  return s;
}
static void release_slot(qr_ctx *c, int &s) {
  if (s >= 0) c->free_slots.push_back(s);
  s = -1;
}

static int prepare_fixed_point(qr_ctx *c) {
  if (c->exact) return QR_OK;
  PhaseTimer pt(c, PH_HIST);
  QR_CUDA(cudaMemsetAsync(c->d_maxabs, 0, 2 * sizeof(unsigned long long), c->stream));
  QR_LAUNCH(c, PH_HIST, maxabs_kernel, 296, 256, 0, c->d_lambda, c->N, c->d_maxabs);
  if (c->lambda) QR_LAUNCH(c, PH_HIST, maxabs_kernel, 296, 256, 0, c->d_weight, c->N, c->d_maxabs + 1);   // leaf outputs' denominators
  if (c->comm) QR_TRY(comm_allreduce_max_u64(c->comm, c->d_maxabs, 2, c->stream));
  QR_LAUNCH(c, PH_HIST, choose_scale_kernel, 1, 2, 0, c->d_maxabs, ceil_log2(c->N_global) + 1, c->d_qexp);
  QR_LAUNCH(c, PH_HIST, quantize_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_lambda, c->N,
            c->d_qexp, c->d_lamq, c->d_node);
  return QR_OK;
}

// host replica of MaxHeap<RTNode*> (maxheap.h:31-106): same sift rules, so equal keys pop in
// the same order as in the reference
struct NodeHeap {
  struct Item { double key; int val; };
  std::vector<Item> arr;
  size_t size = 0;
  NodeHeap() { arr.push_back({DBL_MAX, -1}); }
  void push(double key, int val) {
    ++size;
    if (arr.size() <= size) arr.resize(size + 1);
    size_t p = size;
    while (key > arr[p >> 1].key) { arr[p] = arr[p >> 1]; p >>= 1; }
    arr[p] = {key, val};
  }
  int top() const { return arr[1].val; }
  void pop() {
    const Item last = arr[size--];
    size_t child, p = 1;
    while ((p << 1) <= size) {
      child = p << 1;
      if (child < size && arr[child + 1].key > arr[child].key) ++child;
      if (last.key < arr[child].key) arr[p] = arr[child];
      else break;
      p = child;
    }
    arr[p] = last;
  }
};

// QR_TRACE=1: GPU timeline of the growth rounds (events between the kernels of a round), printed per tree
struct RoundTrace {
  std::vector<cudaEvent_t> ev;   // 4 per round: start, after route/partition, after histogram, after the split scan
  size_t used = 0;
  cudaEvent_t next() {
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    return ev[used++];
  }
};
static RoundTrace g_trace;
static const bool g_trace_on = getenv("QR_TRACE") != nullptr;
static const bool g_pdl_on = getenv("QR_NO_PDL") == nullptr;
static const int g_ktrace_round = getenv("QR_KTRACE") ? atoi(getenv("QR_KTRACE")) : -1;
// host-side stamps of the traced round (ns, steady clock): 0 expand start, 1 tasks built, 2 route launched,
// 3 histogram launched, 4 scan launched, 5 first flag seen, 6 all flags seen, 7 results reduced
static int64_t g_hstamp[8];
static inline void hstamp(int i) {
  if (g_ktrace_round >= 0) g_hstamp[i] = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define QR_TRACE_MARK(c) do { if (g_trace_on) cudaEventRecord(g_trace.next(), (c)->stream); } while (0)

// wait for the k per-task flags of the current round (bounded spin; a launch failure surfaces through
// cudaStreamQuery, a peer that never arrived through the error word the split scan writes)
static int wait_round_flags(qr_ctx *c, uint32_t k) {
  volatile uint32_t *flags = c->h_flags;
  uint64_t spins = 0;
  for (uint32_t j = 0; j < k; ++j) {
    while (flags[j] != c->round_id) {
      if ((++spins & 0xfffff) == 0) {
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) {
          set_error("growth round failed: %s", cudaGetErrorString(e));
          return QR_ECUDA;
        }
        if (e == cudaSuccess && flags[j] != c->round_id) {
          set_error("internal: growth round finished without publishing task %u", j);
          return QR_ECUDA;
        }
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (c->h_err && *(volatile uint32_t *) c->h_err) {
    set_error("histogram exchange timed out waiting for a peer rank (is every rank still alive?)");
    return QR_ECOMM;
  }
  return QR_OK;
}

// Sharded training over peer memory: decides how this round's built histograms are exchanged, before the
// round's task records are made.  Small rounds: the split scan itself adds the peers' staging slots (fused,
// one barrier); wide rounds: the stand-alone in-place reduce-scatter + all-gather kernel (each rank moves
// 2(W-1)/W of the payload instead of reading W-1 copies of it).  Every rank takes the same decision.
constexpr uint32_t kSqRegion = 2048;   // squares partials: two regions of d_sq128, alternating by round
static void begin_exchange_round(qr_ctx *c, uint32_t k) {
  c->round_fused = false;
  c->round_sq_off = 0;
  c->round_parity = 0;
  if (!c->comm || comm_transport(c->comm) != 2) return;
  c->round_parity = c->xround++ & 1u;
  c->round_sq_off = c->round_parity * kSqRegion;
  // one-shot (every rank reads the W-1 other copies itself, one flag barrier) while that is at most a few MB per
  // rank; beyond, the stand-alone reduce-scatter + all-gather moves 2(W-1)/W of the payload instead of W-1 times it
  const size_t oneshot_bytes = (size_t) (comm_world(c->comm) - 1) * k * c->ncells * 12;
  c->round_fused = c->peer_fused && oneshot_bytes <= (size_t) c->oneshot_max << 20;
  // feature-sliced exchange: a rank only ever holds the cumulative histograms of its own features, so every round
  // takes the in-scan path (it reads 1 / world of every peer's bins, what a reduce-scatter would move)
  if (c->sliced) c->round_fused = true;
}
static uint32_t stage_of(const qr_ctx *c, uint32_t j) {
  if (c->fused_scan) return 1u + (uint32_t) c->stage_slot0 + j;   // one GPU: the raw slot of task j (cleared by its scan)
  return c->round_fused ? 1u + (uint32_t) c->stage_slot0 + c->round_parity * c->max_tasks + j : 0u;
}

// One GPU: waits for the (task, feature) flags of the round's split scan (scan_pub_kernel), then takes the first
// maximum over features of every child (rt.cc:297-306: ascending feature index, strict '>') and fills the node
// statistics of RTNode(sampleids, hist) (rtnode.h:97-107) into h_res, where the growth code reads them.
static int collect_fused_round(qr_ctx *c, uint32_t k) {
  const uint32_t tag = c->round_id;
  uint64_t spins = 0;
  bool first = true;
  for (uint32_t j = 0; j < k; ++j) {
    const NodeTask &t = c->h_tasks[j];
    const int nchild = t.whole ? 1 : 2;
    for (int child = 0; child < nchild; ++child) {
      ChildOut *o = c->h_out + (size_t) j * 2 + child;
      volatile uint32_t *tags[3] = {&o->split.tag, &o->where.tag, &o->stats.tag};
      for (int r = 0; r < 3; ++r) {
        while (*tags[r] != tag) {
          if ((++spins & 0xfffff) == 0) {   // bounded: a launch failure surfaces through cudaStreamQuery
            cudaError_t e = cudaStreamQuery(c->stream);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
              set_error("growth round failed: %s", cudaGetErrorString(e));
              return QR_ECUDA;
            }
            if (e == cudaSuccess && *tags[r] != tag) {
              set_error("internal: growth round finished without publishing task %u", j);
              return QR_ECUDA;
            }
          }
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      if (first) { hstamp(5); first = false; }
      const bool built = t.whole || ((child == 0) == (t.build_left != 0u));
      const double sqB = o->stats.v;
      SplitResult r;
      r.n = o->stats.a;
      r.sum = o->where.v;
      r.squares = built ? sqB : t.parent_squares - sqB;          // rtnode_histogram.cc:86,207
      r.deviance = r.squares - r.sum * r.sum / (double) r.n;      // rtnode.h:106
      r.score = o->split.v;
      r.valid = r.score != -1.0;
      r.feature = r.valid ? (o->where.a >> 16) : 0xffffffffu;
      r.threshold_idx = r.valid ? (o->where.a & 0xffffu) : 0xffffffffu;
      r.lcount = r.valid ? o->split.a : 0;
      r.pad = 0;
      c->h_res[(size_t) j * 2 + child] = r;
      *tags[0] = 0u; *tags[1] = 0u; *tags[2] = 0u;   // consumed
    }
  }
  hstamp(6);
  hstamp(7);
  if (c->h_err && *(volatile uint32_t *) c->h_err) {
    set_error("histogram exchange timed out waiting for a peer rank (is every rank still alive?)");
    return QR_ECOMM;
  }
  return QR_OK;
}

// QR_KTRACE=<round>: device-side timeline (%globaltimer stamps of every block, qr_fast_kernels.cuh) of growth
// round <round> of every 50th tree, printed on stderr.  Development aid.
constexpr uint32_t kTraceBlocks = 8192;
static unsigned long long *g_ktrace_dev = nullptr, *g_ktrace_host = nullptr;
static uint64_t g_ktrace_tree = 0;
static bool ktrace_active(const qr_ctx *c) {
  return g_ktrace_round >= 0 && (int) c->nrounds == g_ktrace_round && g_ktrace_tree % 50 == 49;
}
static unsigned long long *ktrace_buffer(const qr_ctx *c, int which) {
  if (!ktrace_active(c)) return nullptr;
  if (!g_ktrace_dev) {
    cudaMalloc((void **) &g_ktrace_dev, 3 * (size_t) kTraceBlocks * kTraceStamps * sizeof(unsigned long long));
    g_ktrace_host = (unsigned long long *) malloc(3 * (size_t) kTraceBlocks * kTraceStamps * sizeof(unsigned long long));
    cudaMemset(g_ktrace_dev, 0, 3 * (size_t) kTraceBlocks * kTraceStamps * sizeof(unsigned long long));
  }
  return g_ktrace_dev + (size_t) which * kTraceBlocks * kTraceStamps;
}
static void ktrace_report(qr_ctx *c, uint32_t route_blocks, uint32_t hist_blocks, uint32_t scan_blocks) {
  if (!ktrace_active(c) || !g_ktrace_dev) return;
  cudaStreamSynchronize(c->stream);
  cudaMemcpy(g_ktrace_host, g_ktrace_dev, 3 * (size_t) kTraceBlocks * kTraceStamps * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  const uint32_t nb[3] = {std::min(route_blocks, kTraceBlocks), std::min(hist_blocks, kTraceBlocks), std::min(scan_blocks, kTraceBlocks)};
  for (int w = 0; w < 3; ++w)
    for (uint32_t b = 0; b < nb[w]; ++b) {
      const unsigned long long v = g_ktrace_host[((size_t) w * kTraceBlocks + b) * kTraceStamps];
      if (v && v < t0) t0 = v;
    }
  const char *names[3] = {"route", "hist", "scan"};
  for (int w = 0; w < 3; ++w) {
    fprintf(stderr, "[ktrace] tree %llu round %u %s (%u blocks): ", (unsigned long long) g_ktrace_tree, c->nrounds, names[w], nb[w]);
    for (uint32_t i = 0; i < kTraceStamps; ++i) {
      unsigned long long mn = ~0ull, mx = 0, cnt = 0;
      for (uint32_t b = 0; b < nb[w]; ++b) {
        const unsigned long long v = g_ktrace_host[((size_t) w * kTraceBlocks + b) * kTraceStamps + i];
        if (!v) continue;
        mn = std::min(mn, v); mx = std::max(mx, v); ++cnt;
      }
      if (cnt) fprintf(stderr, " s%u[n=%llu %.1f..%.1f us]", i, cnt, (double) (mn - t0) * 1e-3, (double) (mx - t0) * 1e-3);
    }
    fprintf(stderr, "\n");
  }
  fprintf(stderr, "[ktrace] host: tasks +%.1f route-launched +%.1f hist-launched +%.1f scan-launched +%.1f first-flag +%.1f all-flags +%.1f reduced +%.1f us\n",
          (g_hstamp[1] - g_hstamp[0]) * 1e-3, (g_hstamp[2] - g_hstamp[0]) * 1e-3, (g_hstamp[3] - g_hstamp[0]) * 1e-3,
          (g_hstamp[4] - g_hstamp[0]) * 1e-3, (g_hstamp[5] - g_hstamp[0]) * 1e-3, (g_hstamp[6] - g_hstamp[0]) * 1e-3,
          (g_hstamp[7] - g_hstamp[0]) * 1e-3);
  cudaMemset(g_ktrace_dev, 0, 3 * (size_t) kTraceBlocks * kTraceStamps * sizeof(unsigned long long));
}

// small rounds carry their task records in the kernel parameters; otherwise they are copied to d_tasks
// (sharded training: only with the peer-memory exchange, whose kernel takes the records the same way)
static int publish_tasks(qr_ctx *c, uint32_t k) {
  c->pack.n = 0;
  if (!c->exact && k <= kPackTasks && (!c->comm || comm_transport(c->comm) == 2)) {
    c->pack.n = k;
    memcpy(c->pack.t, c->h_tasks, k * sizeof(NodeTask));
  } else {
    QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, k * sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  }
  return QR_OK;
}

// REFERENCE mode: which accumulation a task's built child takes (sizes are known on the host)
static uint32_t sq_chunks_host(uint64_t n) { return n > kSqSerialMax ? (uint32_t) ((n + kSqChunk - 1) / kSqChunk) : 0u; }
static uint32_t built_size_host(const NodeTask &t) { return t.whole ? t.n : (t.build_left ? t.lcount : t.n - t.lcount); }
static void plan_exact_tasks(qr_ctx *c, uint32_t k) {
  uint32_t chunk0 = 0;
  for (uint32_t j = 0; j < k; ++j) {
    NodeTask &t = c->h_tasks[j];
    const uint32_t built = built_size_host(t);
    t.walk = (c->d_perm != nullptr && built >= c->walk_min && t.slotB >= 0) ? 1u : 0u;
    t.sq_chunk0 = chunk0;
    chunk0 += sq_chunks_host(built);
  }
}

// REFERENCE mode: the built child of every task in the reference's accumulation order (qr_exact_kernels.cuh).  Small
// children: a warp per (feature, task) over the list; large ones (NodeTask::walk, set by the host, which knows the sizes):
// a thread per histogram cell over the dataset's sorted lists.  Squares: the chain for short lists, the parallel scheme
// for long ones (NodeTask::sq_chunk0).
static int launch_exact_hist(qr_ctx *c, uint32_t k) {
  const uint32_t F = (uint32_t) c->F;
  WalkList wl{};
  uint32_t total_chunks = 0;
  bool whole = false;
  for (uint32_t j = 0; j < k; ++j) {
    const NodeTask &t = c->h_tasks[j];
    if (t.walk) { if (wl.n == kWalkMax) { set_error("internal: more than %u large nodes in a round", kWalkMax); return QR_ECUDA; } wl.idx[wl.n++] = j; whole = whole || t.whole; }
    total_chunks = std::max(total_chunks, t.sq_chunk0 + sq_chunks_host(built_size_host(t)));
  }
  if (wl.n < k)
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      QR_LAUNCH(c, PH_HIST, hist_exact_list_kernel<B>, dim3((F + kExactWarps - 1) / kExactWarps, k), kExactWarps * 32, 0, c->d_tasks,
                c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lambda, c->d_thr_off, F, c->d_fskip, c->d_hist_sum,
                c->d_hist_cnt, c->ncells);
      return QR_OK;
    }));
  if (wl.n) {
    const unsigned cblocks = (c->ncells + kWalkWarps - 1) / kWalkWarps;
    if (whole) {   // the root: every document belongs to it
      QR_LAUNCH(c, PH_HIST, hist_exact_walk_kernel<0>, dim3(cblocks, 1), kWalkWarps * 32, 0, c->d_tasks, wl, c->d_perm, c->d_cell_pos,
                c->d_mark, 0u, c->d_lambda, c->d_thr_off, F, c->d_fskip, c->d_hist_sum, c->d_hist_cnt, c->ncells);
    } else {
      c->mark_tag = (c->mark_tag + 1u) & 0x0fffffffu;
      if (c->mark_tag == 0u) c->mark_tag = 1u;
      QR_LAUNCH(c, PH_HIST, mark_walk_kernel, dim3(296, wl.n), 256, 0, c->d_tasks, wl, c->d_lcount, c->d_ids[0], c->d_ids[1],
                c->d_mark, c->mark_tag);
#define QR_WALK_LAUNCH(KK)                                                                                                  \
  QR_LAUNCH(c, PH_HIST, hist_exact_walk_kernel<KK>, dim3(cblocks, (wl.n + KK - 1) / KK), kWalkWarps * 32, 0, c->d_tasks, wl,   \
            c->d_perm, c->d_cell_pos, c->d_mark, c->mark_tag, c->d_lambda, c->d_thr_off, F, c->d_fskip, c->d_hist_sum,      \
            c->d_hist_cnt, c->ncells)
      if (wl.n == 1) QR_WALK_LAUNCH(1);
      else QR_WALK_LAUNCH(2);
#undef QR_WALK_LAUNCH
    }
    QR_LAUNCH(c, PH_HIST, hist_exact_prefix_kernel, dim3((F + 63) / 64, wl.n), 64, 0, c->d_tasks, wl, c->d_thr_off, F,
              c->d_hist_sum, c->d_hist_cnt, c->ncells);
  }
  QR_LAUNCH(c, PH_HIST, squares_exact_kernel, k, 32, 0, c->d_tasks, c->d_lcount, c->d_lambda, c->d_ids[0], c->d_ids[1],
            c->d_partials, 0);
  if (total_chunks) {
    SqChunk *chunks = static_cast<SqChunk *>(c->d_sq_chunks);
    QR_LAUNCH(c, PH_HIST, ordered_squares_sums_kernel, (total_chunks + 3) / 4, 128, 0, c->d_tasks, k, c->d_lcount, c->d_lambda,
              c->d_ids[0], c->d_ids[1], chunks, total_chunks);
    QR_LAUNCH(c, PH_HIST, ordered_squares_binade_kernel, k, 256, 0, c->d_tasks, c->d_lcount, chunks);
    QR_LAUNCH(c, PH_HIST, ordered_squares_pairs_kernel, (total_chunks + 3) / 4, 128, 0, c->d_tasks, k, c->d_lcount, c->d_lambda,
              c->d_ids[0], c->d_ids[1], chunks, total_chunks);
    QR_LAUNCH(c, PH_HIST, ordered_squares_resolve_kernel, k, 32, 0, c->d_tasks, c->d_lcount, c->d_lambda, c->d_ids[0],
              c->d_ids[1], chunks, c->d_partials, c->d_sq_replayed);
  }
  return QR_OK;
}

// Launches the histogram + split-scan kernels of a round whose task records were published.
static int launch_hist_and_scan(qr_ctx *c, uint32_t k, uint32_t total_slices, bool root, double built_docs) {
  const uint32_t F = (uint32_t) c->F;
  {
    PhaseTimer pt(c, PH_HIST);
    const bool static_counts = root && !c->exact && c->d_root_cnt != nullptr;
    if ((root || c->exact) && !c->fused_scan)   // (FAST child rounds: route_kernel cleared the slots; fused scan: raw slots stay clear)
      QR_LAUNCH(c, PH_HIST, prep_slots_kernel, dim3(std::max<uint32_t>(1, std::min<uint32_t>(32, (c->ncells + 1023) / 1024)), k),
                256, 0, c->d_tasks, c->pack, c->d_hist_sum, c->d_hist_cnt, c->ncells, static_counts ? c->d_root_cnt : nullptr);
    if (c->exact) {
      QR_TRY(launch_exact_hist(c, k));
    } else {
      const size_t smem = (size_t) c->fpp * c->max_thr * 12;
      const bool use_smem = smem <= 200 * 1024;
      if (total_slices > (c->comm ? kSqRegion : c->max_slices - c->max_tasks)) { set_error("internal: %u histogram slices > capacity", total_slices); return QR_ECUDA; }
      const uint32_t *counts = c->d_counts + (size_t) c->count_parity * c->max_tasks;
      unsigned long long *kt = c->fused_scan ? ktrace_buffer(c, 1) : nullptr;
  // (a child round's histogram kernel follows the route kernel: programmatic dependent launch, unless timed)
  const bool pdl = !root && !c->profiling && !g_trace_on && g_pdl_on;
#define QR_HIST_LAUNCH(SMEMF, COUNTF, ACCF)                                                                   \
  do {                                                                                                        \
    if (pdl)                                                                                                  \
      QR_LAUNCH_PDL(c, PH_HIST, (hist_limb_kernel<B, SMEMF, COUNTF, ACCF>), dim3(total_slices, c->npanels), dim3(kHistThreads), \
                    SMEMF ? smem : 0, (const NodeTask *) c->d_tasks, k, c->pack, counts, (const uint4 *) c->d_panels, c->N, \
                    (const uint32_t *) c->d_cids, (const long long *) c->d_clamq, (const long long *) c->d_lamq, \
                    (const uint32_t *) c->d_thr_off, F, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_sq128 + c->round_sq_off, \
                    c->max_thr, c->d_sq_acc, kt, kspan, (const uint4 *) c->d_rows, c->npanels);               \
    else                                                                                                      \
      QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, SMEMF, COUNTF, ACCF>), dim3(total_slices, c->npanels), kHistThreads, \
                SMEMF ? smem : 0, c->d_tasks, k, c->pack, counts, c->d_panels, c->N, c->d_cids, c->d_clamq,   \
                c->d_lamq, c->d_thr_off, F, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_sq128 + c->round_sq_off, \
                c->max_thr, c->d_sq_acc, kt, kspan, c->d_rows, c->npanels);                                   \
  } while (0)
      unsigned long long *kspan = nullptr;
      if (c->profiling) {
        if (!c->d_kspan) QR_TRY(dev_alloc(&c->d_kspan, 2));
        kspan = c->d_kspan;
        QR_CUDA(cudaMemsetAsync(kspan, 0xff, sizeof(unsigned long long), c->stream));
        QR_CUDA(cudaMemsetAsync(kspan + 1, 0, sizeof(unsigned long long), c->stream));
      }
      QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
        using B = decltype(tag);
        if (use_smem) {
          if (c->fused_scan) { if (static_counts) QR_HIST_LAUNCH(true, false, true); else QR_HIST_LAUNCH(true, true, true); }
          else { if (static_counts) QR_HIST_LAUNCH(true, false, false); else QR_HIST_LAUNCH(true, true, false); }
        } else {
          if (c->fused_scan) { if (static_counts) QR_HIST_LAUNCH(false, false, true); else QR_HIST_LAUNCH(false, true, true); }
          else { if (static_counts) QR_HIST_LAUNCH(false, false, false); else QR_HIST_LAUNCH(false, true, false); }
        }
        return QR_OK;
      }));
#undef QR_HIST_LAUNCH
      if (c->profiling) {
        unsigned long long span[2] = {0, 0};
        QR_CUDA(cudaMemcpyAsync(span, kspan, sizeof(span), cudaMemcpyDeviceToHost, c->stream));
        QR_CUDA(cudaStreamSynchronize(c->stream));
        if (span[1] > span[0]) c->histk_ms += (double) (span[1] - span[0]) * 1e-6;
        c->histk_launches++;
        c->histk_docs += built_docs;
      }
    }
    if (c->comm && !c->round_fused) QR_TRY(comm_reduce_tasks(c, k, root));
  }
  QR_TRACE_MARK(c);
  hstamp(3);
  // scan_pub_kernel: one GPU, and sharded training over peer memory (the NCCL fallback keeps scan_kernel)
  const bool pub = c->fused_scan || (c->pub_ok && c->comm && comm_transport(c->comm) == 2);
  if (pub) {
    // the scan reduces over features on the device and publishes tagged records the host polls; launched behind
    // the histogram kernel with programmatic stream serialization (resident and waiting when that kernel ends)
    PhaseTimer pt(c, PH_SCAN);
    c->round_id++;
    ScanOut so{};
    so.out = c->d_out_mapped; so.cand = c->d_cand; so.node = c->d_noderec; so.sq_built = c->d_sq_built;
    so.done = c->d_task_done; so.sq_acc = c->fused_scan ? c->d_sq_acc : nullptr; so.root_cnt = c->d_root_cnt; so.qexp = c->d_qexp;
    so.round_id = c->round_id; so.minls = c->p.minleafsupport; so.ktrace = ktrace_buffer(c, 2);
    const bool static_counts = root && c->d_root_cnt != nullptr;
    // (PDL only directly behind the histogram kernel: a stand-alone exchange kernel in between is not PDL-aware)
    const bool pdl = !c->profiling && !g_trace_on && g_pdl_on && (!c->comm || c->round_fused);
    PeerView pv{};
    if (c->round_fused) comm_peer_view(c, !static_counts, &pv);
    const ulonglong2 *sqp = c->d_sq128 + c->round_sq_off;
    const uint32_t scan_f = pv.f_hi ? pv.f_hi - pv.f_lo : F;   // feature-sliced exchange: this rank's features only
#define QR_SCANPUB_LAUNCH(COUNTF, PEERF)                                                                      \
  do {                                                                                                        \
    if (pdl)                                                                                                  \
      QR_LAUNCH_PDL(c, PH_SCAN, (scan_pub_kernel<COUNTF, PEERF>), dim3(scan_f, k), dim3(kPubThreads), 0, (const NodeTask *) c->d_tasks, \
                    c->pack, c->d_hist_sum, c->d_hist_cnt, c->ncells, (const uint32_t *) c->d_thr_off, F, so, sqp, c->d_err_mapped, pv); \
    else                                                                                                      \
      QR_LAUNCH(c, PH_SCAN, (scan_pub_kernel<COUNTF, PEERF>), dim3(scan_f, k), kPubThreads, 0, c->d_tasks, c->pack, c->d_hist_sum, \
                c->d_hist_cnt, c->ncells, c->d_thr_off, F, so, sqp, c->d_err_mapped, pv);                     \
  } while (0)
    if (c->round_fused) { if (static_counts) QR_SCANPUB_LAUNCH(false, true); else QR_SCANPUB_LAUNCH(true, true); }
    else { if (static_counts) QR_SCANPUB_LAUNCH(false, false); else QR_SCANPUB_LAUNCH(true, false); }
#undef QR_SCANPUB_LAUNCH
    hstamp(4);
    QR_TRACE_MARK(c);
    return collect_fused_round(c, k);
  }
  {
    PhaseTimer pt(c, PH_SCAN);
    // results are written straight into mapped pinned host memory and announced through per-task
    // flags the host polls: no device-to-host copy and no stream synchronisation on the round path
    c->round_id++;
    PeerView pv{};
    if (c->round_fused) comm_peer_view(c, !(root && c->d_root_cnt != nullptr), &pv);
#define QR_SCAN_LAUNCH(PEERF, EXACTF)                                                                          \
  QR_LAUNCH(c, PH_SCAN, (scan_kernel<PEERF, EXACTF>), dim3(F, k), kScanThreads, 0, c->d_tasks, c->pack, c->d_hist_sum,  \
            c->d_hist_cnt, c->ncells, c->d_thr_off, F, c->p.minleafsupport, c->d_qexp, c->d_fbest_score, c->d_fbest_t, \
            c->d_fbest_lc, c->d_totals, c->d_sq_built, c->d_sq128 + c->round_sq_off, c->d_partials, c->d_task_done,    \
            c->d_res_mapped, c->d_flags_mapped, c->round_id, c->d_err_mapped, pv)
    if (c->exact) QR_SCAN_LAUNCH(false, true);
    else if (c->round_fused) QR_SCAN_LAUNCH(true, false);
    else QR_SCAN_LAUNCH(false, false);
#undef QR_SCAN_LAUNCH
    QR_TRACE_MARK(c);
    QR_TRY(wait_round_flags(c, k));
  }
  return QR_OK;
}

// documents per histogram slice so that a round's grid is about one block per SM (148 SMs): every
// block ends by flushing its shared-memory histogram with global atomics, a fixed cost that small
// rounds cannot amortise over more blocks; measured in scripts/hist_mb2.cu.
static uint32_t pick_hist_dpb(const qr_ctx *c, uint64_t total_docs) {
  const uint32_t want_slices = std::max<uint32_t>(1, 148u / c->npanels);
  uint64_t dpb = std::max<uint64_t>(2u * kHistThreads, (total_docs + want_slices - 1) / want_slices);
  dpb = (dpb + 255u) & ~(uint64_t) 255u;
  return (uint32_t) std::min<uint64_t>(dpb, 1u << 20);
}

// root histogram refresh (mart.cc:335) + root node statistics and best split
static int build_root(qr_ctx *c) {
  HostNode root;
  root.lo = 0; root.n = (uint32_t) c->N; root.buf = 2;
  root.hist = alloc_slot(c);
  NodeTask &t = c->h_tasks[0];
  memset(&t, 0, sizeof(t));
  begin_exchange_round(c, 1);
  t.lo = 0; t.n = root.n; t.src = 2; t.dst = 0; t.whole = 1; t.build_left = 1;
  t.slotP = -1; t.slotB = root.hist; t.slotD = -1;
  t.stage1 = stage_of(c, 0);
  // sharded: the slicing is derived from the largest shard, so that it is the same on every rank
  const uint32_t layout_n = c->comm ? (uint32_t) c->N_local_max : root.n;
  t.hist_dpb = pick_hist_dpb(c, layout_n);
  t.hist_blk0 = 0; t.part_blk0 = 0; t.sq0 = 0; t.fused_sq = 0;
  const uint32_t slices = std::max<uint32_t>(1, (layout_n + t.hist_dpb - 1) / t.hist_dpb);
  t.hist_nblk = slices;
  if (c->exact) plan_exact_tasks(c, 1);
  QR_TRACE_MARK(c); QR_TRACE_MARK(c);
  QR_TRY(publish_tasks(c, 1));
  QR_TRY(launch_hist_and_scan(c, 1, slices, true, (double) root.n));
  c->pack.n = 0;
  root.res = c->h_res[0];
  c->nodes.push_back(root);
  return QR_OK;
}

// Per-bin document counts of the whole dataset (they do not depend on the pseudo-responses, so the
// root refresh of every tree reuses them; the reference notes the same at rtnode_histogram.cc:149).
static int init_root_counts(qr_ctx *c) {
  if (c->exact) return QR_OK;
  QR_CUDA(cudaMemsetAsync(c->d_lamq, 0, c->N * sizeof(long long), c->stream));
  const int slot = alloc_slot(c);
  NodeTask &t = c->h_tasks[0];
  memset(&t, 0, sizeof(t));
  t.n = (uint32_t) c->N; t.src = 2; t.whole = 1; t.build_left = 1;
  t.slotP = -1; t.slotB = slot; t.slotD = -1;
  t.hist_dpb = pick_hist_dpb(c, c->N);
  t.hist_nblk = std::max<uint32_t>(1, (t.n + t.hist_dpb - 1) / t.hist_dpb);
  c->pack.n = 0;
  QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  QR_LAUNCH(c, PH_HIST, prep_slots_kernel, dim3(32, 1), 256, 0, c->d_tasks, c->pack, c->d_hist_sum, c->d_hist_cnt, c->ncells,
            (const uint32_t *) nullptr);
  const size_t smem = (size_t) c->fpp * c->max_thr * 12;
  const bool use_smem = smem <= 200 * 1024;
  const uint32_t F = (uint32_t) c->F;
  QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    if (use_smem)
      QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, true, true, false>), dim3(t.hist_nblk, c->npanels), kHistThreads, smem, c->d_tasks, 1u,
                c->pack, c->d_counts, c->d_panels, c->N, c->d_cids, c->d_clamq, c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (ulonglong2 *) nullptr, (unsigned long long *) nullptr,
                (unsigned long long *) nullptr, (const uint4 *) nullptr, c->npanels);
    else
      QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, false, true, false>), dim3(t.hist_nblk, c->npanels), kHistThreads, 0, c->d_tasks, 1u,
                c->pack, c->d_counts, c->d_panels, c->N, c->d_cids, c->d_clamq, c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (ulonglong2 *) nullptr, (unsigned long long *) nullptr,
                (unsigned long long *) nullptr, (const uint4 *) nullptr, c->npanels);
    return QR_OK;
  }));
  if (!c->d_root_cnt) QR_TRY(dev_alloc(&c->d_root_cnt, c->ncells));   // (qr_sample_redraw counts again into the same array)
  QR_CUDA(cudaMemcpyAsync(c->d_root_cnt, c->d_hist_cnt + (size_t) slot * c->ncells, c->ncells * sizeof(uint32_t),
                          cudaMemcpyDeviceToDevice, c->stream));
  if (c->comm) QR_TRY(comm_allreduce_sum_u32(c->comm, c->d_root_cnt, c->ncells, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  int s2 = slot;
  release_slot(c, s2);
  return QR_OK;
}

// links the two children of every expanded node (results of the round are in h_res)
static void link_children(qr_ctx *c, const std::vector<int> &S, bool build_child_hists, const uint32_t *lc_local) {
  for (uint32_t j = 0; j < (uint32_t) S.size(); ++j) {
    const int i = S[j];
    const NodeTask &t = c->h_tasks[j];
    const HostNode nd = c->nodes[i];
    HostNode L, R;
    L.parent = R.parent = i;
    if (lc_local) {   // REFERENCE mode: the children are segments of the id buffer
      L.lo = nd.lo; L.n = lc_local[j]; L.buf = (int) t.dst;
      R.lo = nd.lo + lc_local[j]; R.n = nd.n - lc_local[j]; R.buf = (int) t.dst;
    }
    if (build_child_hists) {
      L.hist = t.build_left ? t.slotB : t.slotD;
      R.hist = t.build_left ? t.slotD : t.slotB;
      L.res = c->h_res[2 * j];
      R.res = c->h_res[2 * j + 1];
    } else {
      // last oblivious level (ot.cc:142-149): plain leaves, no histogram, deviance left at 0
      L.res = SplitResult{}; R.res = SplitResult{};
      L.res.n = nd.res.lcount; R.res.n = nd.res.n - nd.res.lcount;
      L.res.score = R.res.score = -1.0;
    }
    const int li = (int) c->nodes.size();
    c->nodes.push_back(L);
    c->nodes.push_back(R);
    c->nodes[i].left = li;
    c->nodes[i].right = li + 1;
    c->nodes[i].expanded = true;
  }
}

// RegressionTree::split (rt.cc:209-362) for every node in `S` (their best splits are known), FAST mode:
// route -> histogram of the smaller child -> split scan of both children (qr_fast_kernels.cuh)
static int expand_nodes_fast(qr_ctx *c, const std::vector<int> &S, bool build_child_hists) {
  hstamp(0);
  const uint32_t k = (uint32_t) S.size();
  const uint64_t W = c->comm ? (uint64_t) comm_world(c->comm) : 1u;
  uint64_t built_total = 0;
  for (uint32_t j = 0; j < k; ++j) {
    const HostNode &nd = c->nodes[S[j]];
    // local sizes are only known exactly on a single GPU; with several ranks: the expected share
    built_total += std::min(nd.res.lcount, nd.res.n - nd.res.lcount) / W;
  }
  uint32_t dpb = pick_hist_dpb(c, built_total);
  if (build_child_hists) begin_exchange_round(c, k);
  if (c->comm) {
    // The local size of the built child is unknown until the route kernel has run, and the slicing must be
    // the same on every rank (the fused exchange reads the peers' per-slice squares partials): it is derived
    // from replicated quantities only.  bound = min(global size of the built child, largest shard) covers
    // any rank's share; queries are sharded without regard to their content, so the typical share is the
    // global size / W: slices are sized for that (plus a margin), their number for the bound (slices past
    // the real end return at once), capped so that a round never launches waves of empty blocks.
    uint64_t est_total = 0, bound_total = 0;
    for (uint32_t j = 0; j < k; ++j) {
      const HostNode &nd = c->nodes[S[j]];
      const uint64_t built = std::min(nd.res.lcount, nd.res.n - nd.res.lcount);
      const uint64_t bound = std::min<uint64_t>(built, c->N_local_max);
      est_total += std::min<uint64_t>(bound, built / W + built / (4 * W) + 256u);
      bound_total += bound;
    }
    const uint32_t max_slices = 4u * std::max<uint32_t>(1, 148u / c->npanels) + k;
    const uint64_t floor_dpb = ((bound_total + max_slices - 1) / max_slices + 255u) & ~(uint64_t) 255u;
    dpb = (uint32_t) std::min<uint64_t>(std::max<uint64_t>(pick_hist_dpb(c, est_total), floor_dpb), 1u << 20);
  }
  uint32_t hist_blk = 0;
  uint64_t region = 0;
  const uint32_t child_base = (uint32_t) c->nodes.size();
  if ((size_t) child_base + 2 * k > c->max_nodes) { set_error("internal: node table full (%u nodes)", child_base + 2 * k); return QR_ECUDA; }
  for (uint32_t j = 0; j < k; ++j) {
    HostNode &nd = c->nodes[S[j]];
    NodeTask &t = c->h_tasks[j];
    memset(&t, 0, sizeof(t));
    const uint64_t lc = nd.res.lcount, rc = nd.res.n - nd.res.lcount;
    t.f = nd.res.feature; t.t = nd.res.threshold_idx;
    t.build_left = lc <= rc ? 1u : 0u;      // the smaller child is built, its sibling is parent - built (exact in integers)
    t.whole = 0;
    t.slotP = nd.hist;
    t.slotB = t.slotD = -1;
    if (build_child_hists) {
      t.slotB = alloc_slot(c);
      t.slotD = alloc_slot(c);
      if (t.slotB < 0 || t.slotD < 0) { set_error("internal: histogram pool exhausted"); return QR_ECUDA; }
    }
    t.node = (uint32_t) S[j];
    t.child0 = child_base + 2 * j;
    t.region0 = (uint32_t) region;
    const uint64_t built = std::min(lc, rc);
    if (build_child_hists) region += std::min<uint64_t>(built, c->N);
    t.hist_blk0 = hist_blk;
    t.hist_dpb = dpb;
    const uint64_t layout_n = c->comm ? std::min<uint64_t>(built, c->N_local_max) : built;
    t.hist_nblk = std::max<uint32_t>(1, (uint32_t) ((layout_n + dpb - 1) / dpb));
    t.stage1 = build_child_hists ? stage_of(c, j) : 0u;
    hist_blk += t.hist_nblk;
    if (build_child_hists) c->beta += (double) built / (double) c->N_global;
    t.lcount = (uint32_t) lc;
    t.parent_squares = nd.res.squares;
  }
  if (region > c->compact_cap) { set_error("internal: compact lists need %llu entries > capacity %zu", (unsigned long long) region, c->compact_cap); return QR_ECUDA; }
  QR_TRACE_MARK(c);
  QR_TRY(publish_tasks(c, k));
  hstamp(1);
  {
    PhaseTimer pt(c, PH_PARTITION);
    const uint32_t parity = c->count_parity ^ 1u;   // this round's half of the counters; the other half is cleared
    c->count_parity = parity;
    const unsigned grid = (unsigned) ((c->N + kRouteDocs - 1) / kRouteDocs);
    for (uint32_t j0 = 0; j0 < k; j0 += kRouteTasks) {
      const uint32_t kc = std::min<uint32_t>(kRouteTasks, k - j0);
      uint32_t nmin = 0xffffffffu, nmax = 0;
      for (uint32_t j = j0; j < j0 + kc; ++j) { nmin = std::min(nmin, c->h_tasks[j].node); nmax = std::max(nmax, c->h_tasks[j].node); }
      const uint32_t range = nmax - nmin + 1;
      const size_t smem = ((size_t) range * 2 + 15) & ~(size_t) 15;
      TaskPack pk = c->pack;
      if (j0 != 0) pk.n = 0;   // (packed rounds have a single chunk)
      QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
        using B = decltype(tag);
        QR_LAUNCH(c, PH_PARTITION, route_kernel<B>, grid, kRouteThreads, smem, c->d_tasks + j0, kc, pk, c->d_panels, c->N,
                  c->d_node, nmin, range, (const long long *) c->d_lamq, c->d_cids, c->d_clamq,
                  c->d_counts + (size_t) parity * c->max_tasks + j0, c->d_counts + (size_t) (parity ^ 1u) * c->max_tasks,
                  j0 == 0 ? c->max_tasks : 0u, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->fused_scan ? 0 : 1,
                  ktrace_buffer(c, 0));
        return QR_OK;
      }));
    }
  }
  QR_TRACE_MARK(c);
  hstamp(2);
  if (build_child_hists) QR_TRY(launch_hist_and_scan(c, k, hist_blk, false, (double) built_total));
  c->pack.n = 0;
  ktrace_report(c, (uint32_t) ((c->N + kRouteDocs - 1) / kRouteDocs), hist_blk * c->npanels, k * (uint32_t) c->F);
  link_children(c, S, build_child_hists, nullptr);
  c->nrounds++;
  return QR_OK;
}

// The same in REFERENCE mode: stable partition of the node's sample-id list (count -> prefix -> scatter), the
// LEFT child's histogram accumulated in the reference's order, right = parent - left (rt.cc:325-347)
static int expand_nodes_exact(qr_ctx *c, const std::vector<int> &S, bool build_child_hists) {
  const uint32_t k = (uint32_t) S.size();
  uint32_t part_blk = 0;
  for (uint32_t j = 0; j < k; ++j) {
    HostNode &nd = c->nodes[S[j]];
    NodeTask &t = c->h_tasks[j];
    memset(&t, 0, sizeof(t));
    const uint64_t lc = nd.res.lcount;
    t.lo = nd.lo; t.n = nd.n; t.src = (uint32_t) nd.buf; t.dst = nd.buf == 2 ? 0u : (uint32_t) (1 - nd.buf);
    t.f = nd.res.feature; t.t = nd.res.threshold_idx;
    t.build_left = 1u;
    t.whole = 0;
    t.slotP = nd.hist;
    t.slotB = t.slotD = -1;
    if (build_child_hists) {
      t.slotB = alloc_slot(c);
      t.slotD = alloc_slot(c);
      if (t.slotB < 0 || t.slotD < 0) { set_error("internal: histogram pool exhausted"); return QR_ECUDA; }
    }
    t.part_blk0 = part_blk;
    part_blk += std::max<uint32_t>(1, (nd.n + kPartItems - 1) / kPartItems);
    if (build_child_hists) c->beta += (double) lc / (double) c->N_global;
    t.lcount = (uint32_t) lc;
    t.lc_known = 1u;
    t.sq0 = j;
    t.fused_sq = 1;
    t.parent_squares = nd.res.squares;
  }
  plan_exact_tasks(c, k);
  QR_TRACE_MARK(c);
  c->pack.n = 0;
  QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, k * sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  {
    PhaseTimer pt(c, PH_PARTITION);
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      QR_LAUNCH(c, PH_PARTITION, part_count_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                c->d_ids[0], c->d_ids[1], c->d_blockcnt);
      QR_LAUNCH(c, PH_PARTITION, part_prefix_kernel, k, 256, 0, c->d_tasks, c->d_blockcnt, c->d_lcount);
      QR_LAUNCH(c, PH_PARTITION, part_scatter_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                c->d_ids[0], c->d_ids[1], c->d_ids[0], c->d_ids[1], c->d_blockcnt, c->d_lcount);
      return QR_OK;
    }));
  }
  QR_TRACE_MARK(c);
  if (build_child_hists) QR_TRY(launch_hist_and_scan(c, k, 0, false, 0.0));
  std::vector<uint32_t> lcs(k);
  for (uint32_t j = 0; j < k; ++j) lcs[j] = (uint32_t) c->nodes[S[j]].res.lcount;
  link_children(c, S, build_child_hists, lcs.data());
  c->nrounds++;
  return QR_OK;
}

static int expand_nodes(qr_ctx *c, const std::vector<int> &S, bool build_child_hists) {
  const uint32_t k = (uint32_t) S.size();
  if (k == 0) return QR_OK;
  if (k > c->max_tasks) { set_error("internal: %u tasks > capacity %u", k, c->max_tasks); return QR_ECUDA; }
  return c->exact ? expand_nodes_exact(c, S, build_child_hists) : expand_nodes_fast(c, S, build_child_hists);
}

static bool can_split(const qr_ctx *c, int i) {
  const SplitResult &r = c->nodes[i].res;
  return r.deviance > 0.0 && r.valid;   // rt.cc:212, 312
}

// replay of RegressionTree::fit (rt.cc:49-84)
static int fit_leafwise(qr_ctx *c) {
  const size_t nleaves = c->p.nleaves;
  NodeHeap heap;
  size_t taken = 0;
  bool root_done = false;
  auto push_children = [&](int i) {
    c->nodes[i].pushed = true;
    const HostNode &nd = c->nodes[i];
    heap.push(c->nodes[nd.left].res.deviance, nd.left);      // rt.cc:59-60, 72-73
    heap.push(c->nodes[nd.right].res.deviance, nd.right);
    c->rho += (double) c->nodes[nd.left].res.n / (double) c->N_global;
    c->sigma += (double) nd.res.n / (double) c->N_global;
    c->nsplits++;
  };
  for (;;) {
    int need = -1;
    if (!root_done) {
      if (can_split(c, 0)) {
        if (!c->nodes[0].expanded) need = 0;
        else { push_children(0); root_done = true; }
      } else {
        root_done = true;
      }
    }
    if (need < 0 && root_done) {
      while (heap.size != 0 && (nleaves == 0 || taken + heap.size < nleaves)) {   // rt.cc:64-65
        const int i = heap.top();
        if (can_split(c, i)) {
          if (!c->nodes[i].expanded) { need = i; break; }
          heap.pop();
          push_children(i);
        } else {
          heap.pop();
          ++taken;                                                            // rt.cc:78-79
        }
        release_slot(c, c->nodes[i].hist);                                    // rt.cc:83-84
      }
    }
    if (need < 0) break;
    // expansion set: the blocking node plus the frontier nodes still reachable with the budget
    std::vector<int> S{need};
    if (need != 0) {
      const size_t budget = nleaves == 0 ? heap.size : nleaves - taken - heap.size;   // successes left
      std::vector<std::pair<double, int>> cand;
      for (size_t p = 1; p <= heap.size; ++p) {
        const int i = heap.arr[p].val;
        if (i != need && can_split(c, i) && !c->nodes[i].expanded) cand.push_back({heap.arr[p].key, i});
      }
      std::sort(cand.begin(), cand.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) {
        return a.first > b.first || (a.first == b.first && a.second < b.second);
      });
      for (size_t q = 0; q < cand.size() && S.size() < budget && S.size() < c->max_tasks; ++q) S.push_back(cand[q].second);
    }
    QR_TRY(expand_nodes(c, S, true));
  }
  return QR_OK;
}

static int fit_oblivious(qr_ctx *c) {
  const uint32_t depth = c->p.treedepth;
  std::vector<int> level{0};
  for (uint32_t d = 0; d < depth; ++d) {
    std::vector<int> slots;
    for (int i : level) slots.push_back(c->nodes[i].hist);
    const uint32_t nn = (uint32_t) slots.size();
    {
      PhaseTimer pt(c, PH_SCAN);
      QR_CUDA(cudaMemcpyAsync(c->d_obv_slots, slots.data(), nn * sizeof(int), cudaMemcpyHostToDevice, c->stream));
      if (c->exact) QR_LAUNCH(c, PH_SCAN, obv_level_kernel<true>, (unsigned) c->F, 256, 0, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_obv_slots, nn, c->d_thr_off, (uint32_t) c->F, c->p.minleafsupport, c->d_qexp, c->d_obv_scores);
      else QR_LAUNCH(c, PH_SCAN, obv_level_kernel<false>, (unsigned) c->F, 256, 0, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_obv_slots, nn, c->d_thr_off, (uint32_t) c->F, c->p.minleafsupport, c->d_qexp, c->d_obv_scores);
      QR_LAUNCH(c, PH_SCAN, obv_argmax_kernel, 1, 256, 0, c->d_obv_scores, c->d_thr_off, (uint32_t) c->F, c->d_hist_cnt, c->ncells, c->d_obv_slots, nn, c->d_res, c->d_obv_lcounts);
      QR_CUDA(cudaMemcpyAsync(c->h_res, c->d_res, sizeof(SplitResult), cudaMemcpyDeviceToHost, c->stream));
      QR_CUDA(cudaMemcpyAsync(c->h_obv_lcounts, c->d_obv_lcounts, nn * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
      QR_CUDA(cudaStreamSynchronize(c->stream));
    }
    const SplitResult best = c->h_res[0];
    if (!best.valid) break;                                   // ot.cc:96
    for (uint32_t k = 0; k < nn; ++k) {
      SplitResult &r = c->nodes[level[k]].res;
      r.feature = best.feature;
      r.threshold_idx = best.threshold_idx;
      r.lcount = c->h_obv_lcounts[k];
      r.valid = 1;
    }
    QR_TRY(expand_nodes(c, level, d != depth - 1));           // ot.cc:99-161; no histograms on the last level (:127)
    std::vector<int> next;
    for (int i : level) {
      next.push_back(c->nodes[i].left);
      next.push_back(c->nodes[i].right);
      c->rho += (double) c->nodes[c->nodes[i].left].res.n / (double) c->N_global;
      c->sigma += (double) c->nodes[i].res.n / (double) c->N_global;
      c->nsplits++;
      release_slot(c, c->nodes[i].hist);                      // ot.cc:157-160
    }
    level.swap(next);
  }
  return QR_OK;
}

static void collect_leaves(qr_ctx *c, int i) {
  if (c->nodes[i].is_leaf()) { c->leaves.push_back(i); return; }
  collect_leaves(c, c->nodes[i].left);      // rtnode.cc:34-46: left to right
  collect_leaves(c, c->nodes[i].right);
}

// a node that was expanded speculatively but never popped by the replay stays a leaf
static void prune_unreached(qr_ctx *c, const std::vector<char> &is_split) {
  for (size_t i = 0; i < c->nodes.size(); ++i)
    if (!is_split[i]) { c->nodes[i].left = c->nodes[i].right = -1; }
}

static void flatten(const qr_ctx *c, int i, qr_flat_tree *t, uint32_t *next) {
  const HostNode &nd = c->nodes[i];
  const uint32_t id = (*next)++;
  const bool leaf = nd.is_leaf();
  t->feature[id] = leaf ? -1 : (int32_t) nd.res.feature;
  t->threshold_idx[id] = leaf ? 0xffffffffu : nd.res.threshold_idx;
  t->threshold[id] = leaf ? 0.f : c->thr[nd.res.feature][nd.res.threshold_idx];   // rt.cc:317-318
  t->left[id] = t->right[id] = -1;
  if (t->value) t->value[id] = leaf ? nd.value : (nd.res.n ? nd.res.sum / (double) nd.res.n : 0.0);  // rtnode.h:105
  if (t->deviance) t->deviance[id] = nd.res.deviance;
  if (t->count) t->count[id] = nd.res.n;
  if (!leaf) {
    t->left[id] = (int32_t) *next;
    flatten(c, nd.left, t, next);
    t->right[id] = (int32_t) *next;
    flatten(c, nd.right, t, next);
  }
}

static uint32_t count_reachable(const qr_ctx *c, int i) {
  const HostNode &nd = c->nodes[i];
  return nd.is_leaf() ? 1u : 1u + count_reachable(c, nd.left) + count_reachable(c, nd.right);
}

// leaf outputs (rt.cc:165-207), REFERENCE mode: one warp per leaf, sums in list order
static int fit_leaves_exact(qr_ctx *c) {
  const size_t nl = c->leaves.size();
  PhaseTimer pt(c, PH_LEAF);
  for (size_t k = 0; k < nl; ++k) {
    const HostNode &nd = c->nodes[c->leaves[k]];
    c->h_segs[k] = LeafSeg{nd.lo, nd.n, (uint32_t) nd.buf, 0u};
  }
  QR_CUDA(cudaMemcpyAsync(c->d_segs, c->h_segs, nl * sizeof(LeafSeg), cudaMemcpyHostToDevice, c->stream));
  const double *w = c->lambda ? c->d_weight : nullptr;
  QR_LAUNCH(c, PH_LEAF, leaf_exact_kernel, (unsigned) nl, 32, 0, c->d_segs, c->d_ids[0], c->d_ids[1], c->d_lambda,
            w, c->d_leafval, c->d_leaf_of_doc);
  QR_CUDA(cudaMemcpyAsync(c->h_leafval, c->d_leafval, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k < nl; ++k) c->nodes[c->leaves[k]].value = c->h_leafval[k];
  return QR_OK;
}

// FAST mode: one pass over node_of_doc.  Every node id maps to the leaf it ended up under: a node that was
// expanded speculatively but never popped by the replay is a leaf whose documents already carry its
// children's ids.
static int fit_leaves_fast(qr_ctx *c) {
  const size_t nl = c->leaves.size(), nn = c->nodes.size();
  PhaseTimer pt(c, PH_LEAF);
  uint16_t *lut = reinterpret_cast<uint16_t *>(c->h_leafmeta);
  unsigned long long *leafn = reinterpret_cast<unsigned long long *>(c->h_leafmeta + c->leafn_off);
  for (size_t i = 0; i < nn; ++i) lut[i] = 0xffffu;
  for (size_t k = 0; k < nl; ++k) { lut[c->leaves[k]] = (uint16_t) k; leafn[k] = c->nodes[c->leaves[k]].res.n; }
  for (size_t i = 1; i < nn; ++i)   // parents precede their children
    if (lut[i] == 0xffffu && lut[c->nodes[i].parent] != 0xffffu) lut[i] = lut[c->nodes[i].parent];
  QR_CUDA(cudaMemcpyAsync(c->d_leafmeta, c->h_leafmeta, c->leafn_off + nl * sizeof(unsigned long long),
                          cudaMemcpyHostToDevice, c->stream));
  const uint16_t *d_lut = reinterpret_cast<const uint16_t *>(c->d_leafmeta);
  const unsigned long long *d_leafn = reinterpret_cast<const unsigned long long *>(c->d_leafmeta + c->leafn_off);
  // warps per block: as many as the per-warp accumulator tables leave room for
  const size_t lut_bytes = (nn * 2 + 15) & ~(size_t) 15;
  uint32_t wpb = 8;
  while (wpb > 1 && (size_t) wpb * (nl + 32) * 16 + lut_bytes > 160 * 1024) wpb >>= 1;
  const size_t smem = (size_t) wpb * (nl + 32) * 16 + lut_bytes;
  if (smem > 200 * 1024) { set_error("a tree of %zu leaves exceeds the leaf-fit kernel's shared-memory budget", nl); return QR_ELIMIT; }
  if (smem > c->leaf_smem_set) {
    QR_CUDA(cudaFuncSetAttribute(leaf_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) std::max<size_t>(smem, 48 * 1024)));
    c->leaf_smem_set = std::max<size_t>(smem, 48 * 1024);
  }
  const uint32_t blocks = (uint32_t) ((c->N + (size_t) wpb * kLeafWarpDocs - 1) / ((size_t) wpb * kLeafWarpDocs));
  if ((size_t) blocks * nl > c->leaf_part_cap) {
    if (c->d_leaf_partials) cudaFree(c->d_leaf_partials);
    c->d_leaf_partials = nullptr;
    c->leaf_part_cap = (size_t) blocks * nl * 2;
    QR_TRY(dev_alloc(&c->d_leaf_partials, c->leaf_part_cap));
  }
  const double *w = c->lambda ? c->d_weight : nullptr;
  QR_LAUNCH(c, PH_LEAF, leaf_node_kernel, blocks, wpb * 32, smem, c->d_node, d_lut, (uint32_t) nn, (uint32_t) nl, c->d_lamq,
            w, c->d_qexp, c->N, reinterpret_cast<longlong2 *>(c->d_leaf_partials), c->d_leaf_of_doc);
  QR_LAUNCH(c, PH_LEAF, leaf_reduce_kernel, (unsigned) ((nl + 7) / 8), 256, 0, reinterpret_cast<const longlong2 *>(c->d_leaf_partials),
            blocks, (uint32_t) nl, d_leafn, c->lambda, c->d_qexp, reinterpret_cast<longlong2 *>(c->d_leafsum), c->d_leafval);
  if (c->comm) QR_TRY(comm_leaf_values(c, (uint32_t) nl));
  QR_CUDA(cudaMemcpyAsync(c->h_leafval, c->d_leafval, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k < nl; ++k) c->nodes[c->leaves[k]].value = c->h_leafval[k];
  return QR_OK;
}

static int fit_leaves(qr_ctx *c) { return c->exact ? fit_leaves_exact(c) : fit_leaves_fast(c); }

static int fit_tree(qr_ctx *c, qr_flat_tree *out) {
  // release histograms still held by the previous tree
  for (auto &nd : c->nodes) release_slot(c, nd.hist);
  c->nodes.clear();
  c->leaves.clear();
  c->rho = c->sigma = 0;
  c->beta = 1.0;
  c->nsplits = 0;
  c->nrounds = 0;
  c->has_tree = false;

  ++g_ktrace_tree;
  QR_TRY(prepare_fixed_point(c));
  QR_TRY(build_root(c));
  QR_TRY(c->oblivious ? fit_oblivious(c) : fit_leafwise(c));

  if (!c->oblivious) {
    // keep only the splits the replay actually performed: children are linked at expansion time,
    // a node is an internal node of the final tree iff the replay pushed its children
    std::vector<char> is_split(c->nodes.size(), 0);
    for (size_t i = 0; i < c->nodes.size(); ++i) is_split[i] = c->nodes[i].pushed;
    prune_unreached(c, is_split);
  }
  collect_leaves(c, 0);
  QR_TRY(fit_leaves(c));
  c->has_tree = true;
  if (g_trace_on && g_trace.used >= 4) {
    cudaStreamSynchronize(c->stream);
    double tp = 0, th = 0, tf = 0, tg = 0;
    const size_t nr = g_trace.used / 4;
    for (size_t r = 0; r < nr; ++r) {
      float a = 0, b = 0, d = 0, g = 0;
      cudaEventElapsedTime(&a, g_trace.ev[4 * r], g_trace.ev[4 * r + 1]);
      cudaEventElapsedTime(&b, g_trace.ev[4 * r + 1], g_trace.ev[4 * r + 2]);
      cudaEventElapsedTime(&d, g_trace.ev[4 * r + 2], g_trace.ev[4 * r + 3]);
      if (r + 1 < nr) cudaEventElapsedTime(&g, g_trace.ev[4 * r + 3], g_trace.ev[4 * r + 4]);
      tp += a; th += b; tf += d; tg += g;
      if (getenv("QR_TRACE_ROUNDS")) fprintf(stderr, "[trace] round %2zu: route %6.1f hist %6.1f scan %6.1f gap-to-next %6.1f us\n", r, a * 1e3, b * 1e3, d * 1e3, g * 1e3);
    }
    fprintf(stderr, "[trace] %zu rounds: route %.0f hist %.0f scan %.0f gaps %.0f us\n", nr, tp * 1e3, th * 1e3, tf * 1e3, tg * 1e3);
    g_trace.used = 0;
  }

  if (out) {
    const uint32_t nn = count_reachable(c, 0);
    if (out->capacity < nn) { set_error("qr_flat_tree capacity %u < %u nodes", out->capacity, nn); return QR_EINVAL; }
    uint32_t next = 0;
    flatten(c, 0, out, &next);
    out->nnodes = nn;
    out->nleaves = (uint32_t) c->leaves.size();
  }
  return QR_OK;
}

static int update_modelscores(qr_ctx *c, double weight) {
  if (!c->has_tree) { set_error("qr_update_modelscores: no fitted tree"); return QR_EINVAL; }
  PhaseTimer pt(c, PH_LEAF);
  QR_LAUNCH(c, PH_LEAF, update_scores_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_leaf_of_doc,
            c->d_leafval, weight, c->N, c->d_scores);
  c->ranking_valid = false;
  return QR_OK;
}
