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
  QR_CUDA(cudaMemsetAsync(c->d_maxabs, 0, sizeof(unsigned long long), c->stream));
  QR_LAUNCH(c, PH_HIST, maxabs_kernel, 296, 256, 0, c->d_lambda, c->N, c->d_maxabs);
  if (c->comm) QR_TRY(comm_allreduce_max_u64(c->comm, c->d_maxabs, 1, c->stream));
  QR_LAUNCH(c, PH_HIST, choose_scale_kernel, 1, 1, 0, c->d_maxabs, ceil_log2(c->N_global) + 1, c->d_qexp);
  QR_LAUNCH(c, PH_HIST, quantize_kernel, (unsigned) ((c->N + 255) / 256), 256, 0, c->d_lambda, c->N,
            c->d_qexp, c->d_lamq);
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

// Launches the histogram + finalize kernels for the tasks already uploaded to c->d_tasks.
// `slots_ready`: the slots were already cleared by the one-pass partition kernel.
// QR_TRACE=1: GPU timeline of the growth rounds (events between the kernels of a round), printed per tree
struct RoundTrace {
  std::vector<cudaEvent_t> ev;   // 4 per round: start, after partition, after histogram, after finalize
  size_t used = 0;
  cudaEvent_t next() {
    if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    return ev[used++];
  }
};
static RoundTrace g_trace;
static const bool g_trace_on = getenv("QR_TRACE") != nullptr;
#define QR_TRACE_MARK(c) do { if (g_trace_on) cudaEventRecord(g_trace.next(), (c)->stream); } while (0)

// wait for the k per-task flags of the current round (bounded spin; a launch failure surfaces through
// cudaStreamQuery)
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
  c->round_fused = c->peer_fused && (uint32_t) (comm_world(c->comm) - 1) * k <= c->oneshot_max;
}
static uint32_t stage_of(const qr_ctx *c, uint32_t j) {
  return c->round_fused ? 1u + (uint32_t) c->stage_slot0 + c->round_parity * c->max_tasks + j : 0u;
}

static int launch_hist_and_scan(qr_ctx *c, uint32_t k, uint32_t total_slices, bool root, bool slots_ready,
                                double built_docs) {
  const uint32_t F = (uint32_t) c->F;
  if (root) { QR_TRACE_MARK(c); QR_TRACE_MARK(c); }
  {
    PhaseTimer pt(c, PH_HIST);
    const bool static_counts = root && !c->exact && c->d_root_cnt != nullptr;
    if (!slots_ready)
      QR_LAUNCH(c, PH_HIST, prep_slots_kernel, dim3(std::max<uint32_t>(1, std::min<uint32_t>(32, (c->ncells + 1023) / 1024)), k),
                256, 0, c->d_tasks, c->d_hist_sum, c->d_hist_cnt, c->ncells, static_counts ? c->d_root_cnt : nullptr);
    if (c->exact) {
      QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
        using B = decltype(tag);
        QR_LAUNCH(c, PH_HIST, hist_exact_kernel<B>, dim3((F + 3) / 4, k), 128, 0, c->d_tasks, c->d_lcount,
                  c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lambda, c->d_thr_off, F, c->d_hist_sum,
                  c->d_hist_cnt, c->ncells);
        return QR_OK;
      }));
      QR_LAUNCH(c, PH_HIST, squares_exact_kernel, k, 32, 0, c->d_tasks, c->d_lcount, c->d_lambda, c->d_ids[0],
                c->d_ids[1], c->d_partials);
    } else {
      const size_t smem = (size_t) c->fpp * c->max_thr * 12;
      const bool use_smem = smem <= 200 * 1024;
      if (total_slices > (c->comm ? kSqRegion : c->max_slices - c->max_tasks)) { set_error("internal: %u histogram slices > capacity", total_slices); return QR_ECUDA; }
#define QR_HIST_LAUNCH(SMEMF, COUNTF)                                                                        \
  QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, SMEMF, COUNTF>), dim3(total_slices, c->npanels), kHistThreads,   \
            SMEMF ? smem : 0, c->d_tasks, k, c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1],        \
            c->d_lamq, c->d_thr_off, F, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_sq128 + c->round_sq_off, c->max_thr, \
            (const RoundHdr *) nullptr, c->pack, (const long long *) (c->part_3pass ? nullptr : c->d_lamq_c))
      if (c->profiling) cudaEventRecord(c->ev_k0, c->stream);
      QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
        using B = decltype(tag);
        if (use_smem) {
          if (static_counts) QR_HIST_LAUNCH(true, false);
          else QR_HIST_LAUNCH(true, true);
        } else {
          if (static_counts) QR_HIST_LAUNCH(false, false);
          else QR_HIST_LAUNCH(false, true);
        }
        return QR_OK;
      }));
#undef QR_HIST_LAUNCH
      if (c->profiling) {
        cudaEventRecord(c->ev_k1, c->stream);
        cudaEventSynchronize(c->ev_k1);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1);
        c->histk_ms += ms;
        c->histk_launches++;
        c->histk_docs += built_docs;
      }
    }
    if (c->comm && !c->round_fused) QR_TRY(comm_reduce_tasks(c, k, root));
  }
  QR_TRACE_MARK(c);
  {
    PhaseTimer pt(c, PH_SCAN);
    // results are written straight into mapped pinned host memory and announced through per-task
    // flags the host polls: no device-to-host copy and no stream synchronisation on the round path
    c->round_id++;
    PeerView pv{};
    if (c->round_fused) comm_peer_view(c, !(root && c->d_root_cnt != nullptr), &pv);
    if (c->exact)
      QR_LAUNCH(c, PH_SCAN, finalize_kernel<true>, dim3(fin_blocks(F), k), kFinWarps * 32, 0, c->d_tasks, c->d_hist_sum, c->d_hist_cnt,
                c->ncells, c->d_thr_off, F, c->p.minleafsupport, c->d_qexp, c->d_fbest_score, c->d_fbest_t,
                c->d_fbest_lc, c->d_totals, c->d_sq128, c->d_partials, c->d_task_done, c->d_res_mapped,
                c->d_flags_mapped, c->round_id, (const RoundHdr *) nullptr, c->pack, pv);
    else if (c->round_fused)
      QR_LAUNCH(c, PH_SCAN, (finalize_kernel<false, true>), dim3(fin_blocks(F), k), kFinWarps * 32, 0, c->d_tasks, c->d_hist_sum, c->d_hist_cnt,
                c->ncells, c->d_thr_off, F, c->p.minleafsupport, c->d_qexp, c->d_fbest_score, c->d_fbest_t,
                c->d_fbest_lc, c->d_totals, c->d_sq128 + c->round_sq_off, c->d_partials, c->d_task_done, c->d_res_mapped,
                c->d_flags_mapped, c->round_id, (const RoundHdr *) nullptr, c->pack, pv);
    else
      QR_LAUNCH(c, PH_SCAN, finalize_kernel<false>, dim3(fin_blocks(F), k), kFinWarps * 32, 0, c->d_tasks, c->d_hist_sum, c->d_hist_cnt,
                c->ncells, c->d_thr_off, F, c->p.minleafsupport, c->d_qexp, c->d_fbest_score, c->d_fbest_t,
                c->d_fbest_lc, c->d_totals, c->d_sq128 + c->round_sq_off, c->d_partials, c->d_task_done, c->d_res_mapped,
                c->d_flags_mapped, c->round_id, (const RoundHdr *) nullptr, c->pack, pv);
    QR_TRACE_MARK(c);
    if (c->comm && c->part_3pass) {
      QR_CUDA(cudaMemcpyAsync(c->h_lcount, c->d_lcount, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      QR_CUDA(cudaStreamSynchronize(c->stream));
    }
    // (sharded, one-pass partition: the local left counts were written to mapped host memory by the
    // partition kernel, two kernels before the flags below)
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
  QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  QR_TRY(launch_hist_and_scan(c, 1, slices, true, false, (double) root.n));
  root.res = c->h_res[0];
  c->nodes.push_back(root);
  return QR_OK;
}

// Per-bin document counts of the whole dataset (they do not depend on the pseudo-responses, so the
// root refresh of every tree reuses them; the reference notes the same at rtnode_histogram.cc:149).
static int init_root_counts(qr_ctx *c) {
  if (c->exact) return QR_OK;
  QR_CUDA(cudaMemsetAsync(c->d_lamq, 0, c->N * sizeof(long long), c->stream));
  HostNode root;
  root.n = (uint32_t) c->N;
  const int slot = alloc_slot(c);
  NodeTask &t = c->h_tasks[0];
  memset(&t, 0, sizeof(t));
  t.n = root.n; t.src = 2; t.whole = 1; t.build_left = 1;
  t.slotP = -1; t.slotB = slot; t.slotD = -1;
  t.hist_dpb = pick_hist_dpb(c, root.n);
  t.hist_nblk = std::max<uint32_t>(1, (root.n + t.hist_dpb - 1) / t.hist_dpb);
  QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  QR_LAUNCH(c, PH_HIST, prep_slots_kernel, dim3(32, 1), 256, 0, c->d_tasks, c->d_hist_sum, c->d_hist_cnt, c->ncells,
            (const uint32_t *) nullptr);
  const size_t smem = (size_t) c->fpp * c->max_thr * 12;
  const bool use_smem = smem <= 200 * 1024;
  const uint32_t F = (uint32_t) c->F;
  QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
    using B = decltype(tag);
    if (use_smem)
      QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, true, true>), dim3(t.hist_nblk, c->npanels), kHistThreads, smem, c->d_tasks, 1u,
                c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) nullptr, c->pack, (const long long *) nullptr);
    else
      QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, false, true>), dim3(t.hist_nblk, c->npanels), kHistThreads, 0, c->d_tasks, 1u,
                c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) nullptr, c->pack, (const long long *) nullptr);
    return QR_OK;
  }));
  QR_TRY(dev_alloc(&c->d_root_cnt, c->ncells));
  QR_CUDA(cudaMemcpyAsync(c->d_root_cnt, c->d_hist_cnt + (size_t) slot * c->ncells, c->ncells * sizeof(uint32_t),
                          cudaMemcpyDeviceToDevice, c->stream));
  if (c->comm) QR_TRY(comm_allreduce_sum_u32(c->comm, c->d_root_cnt, c->ncells, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  int s2 = slot;
  release_slot(c, s2);
  return QR_OK;
}

// RegressionTree::split (rt.cc:209-362) for every node in `S` (their best splits are known)
static int expand_nodes(qr_ctx *c, const std::vector<int> &S, bool build_child_hists) {
  const uint32_t k = (uint32_t) S.size();
  if (k == 0) return QR_OK;
  if (k > c->max_tasks) { set_error("internal: %u tasks > capacity %u", k, c->max_tasks); return QR_ECUDA; }
  uint64_t built_total = 0;
  for (uint32_t j = 0; j < k; ++j) {
    const HostNode &nd = c->nodes[S[j]];
    const uint64_t lc = nd.res.lcount, rc = nd.res.n - nd.res.lcount;
    const bool build_left = c->exact ? true : lc <= rc;
    // local sizes are only known exactly on a single GPU; with several ranks: the expected share
    built_total += c->comm ? std::min(lc, rc) / (uint64_t) comm_world(c->comm) : (build_left ? lc : rc);
  }
  uint32_t dpb = pick_hist_dpb(c, built_total);
  if (build_child_hists) begin_exchange_round(c, k);
  if (c->comm) {
    // The local size of the built child is unknown until the partition has run, and the slicing must be the
    // same on every rank (the fused exchange reads the peers' per-slice squares partials): it is derived
    // from replicated quantities only.  bound = min(global size of the built child, largest shard) covers
    // any rank's share; queries are sharded without regard to their content, so the typical share is the
    // global size / W: slices are sized for that (plus a margin), their number for the bound (slices past
    // the real end return at once), capped so that a round never launches waves of empty blocks.
    const uint64_t W = (uint64_t) comm_world(c->comm);
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
  uint32_t part_blk = 0, hist_blk = 0;
  for (uint32_t j = 0; j < k; ++j) {
    HostNode &nd = c->nodes[S[j]];
    NodeTask &t = c->h_tasks[j];
    memset(&t, 0, sizeof(t));
    const uint64_t lc = nd.res.lcount, rc = nd.res.n - nd.res.lcount;
    t.lo = nd.lo; t.n = nd.n; t.src = (uint32_t) nd.buf; t.dst = nd.buf == 2 ? 0u : (uint32_t) (1 - nd.buf);
    t.f = nd.res.feature; t.t = nd.res.threshold_idx;
    t.build_left = c->exact ? 1u : (lc <= rc ? 1u : 0u);
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
    t.hist_blk0 = hist_blk;
    t.hist_dpb = dpb;
    const uint64_t built_n = c->comm ? std::min<uint64_t>(std::min(lc, rc), c->N_local_max) : (t.build_left ? lc : rc);
    t.hist_nblk = std::max<uint32_t>(1, (uint32_t) ((built_n + dpb - 1) / dpb));
    t.stage1 = build_child_hists ? stage_of(c, j) : 0u;
    hist_blk += t.hist_nblk;
    if (build_child_hists) c->beta += (double) (t.build_left ? lc : rc) / (double) c->N_global;
    t.lcount = (uint32_t) lc;
    t.lc_known = c->comm ? 0u : 1u;
    t.sq0 = j;
    t.fused_sq = 1;
    t.parent_squares = nd.res.squares;
  }
  QR_TRACE_MARK(c);
  const bool onepass = !c->part_3pass;
  c->pack.n = 0;
  // (sharded training: only with the peer-memory exchange, whose kernel takes the records the same way)
  if (onepass && !c->exact && (!c->comm || comm_transport(c->comm) == 2) && build_child_hists && k <= kPackTasks) {
    // small round: the task records travel in the kernel parameters
    c->pack.n = k;
    memcpy(c->pack.t, c->h_tasks, k * sizeof(NodeTask));
  } else {
    QR_CUDA(cudaMemcpyAsync(c->d_tasks, c->h_tasks, k * sizeof(NodeTask), cudaMemcpyHostToDevice, c->stream));
  }
  // one launch for the whole round (qr_round_kernel.cuh) when the per-phase breakdown is not asked for
  const size_t fused_smem = (size_t) c->fpp * c->max_thr * 12;
  const bool fused = c->fused_rounds && onepass && build_child_hists && !c->profiling && !g_trace_on &&
                     fused_smem <= 200 * 1024 && hist_blk <= c->max_slices - c->max_tasks;
  if (fused) {
    c->part_epoch++;
    c->round_id++;
    // QR_FUSE_PARTITION=1 also chains the partition inside the launch; by default it stays a kernel of its
    // own: its 256-thread blocks without shared memory fit 6 to an SM, the round kernel's blocks only 2
    const bool fuse_partition = c->fuse_partition;
    const uint32_t fused_part = fuse_partition ? part_blk : 0u;
    const uint32_t grid = fused_part + hist_blk * c->npanels;
    RoundCounters rc{c->d_part_done, c->d_panel_done, c->d_task_done};
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      if (!fuse_partition) {
        QR_LAUNCH(c, PH_PARTITION, partition_onepass_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                  c->d_ids[0], c->d_ids[1], c->d_ids[0], c->d_ids[1], c->d_part_status, c->d_ticket,
                  c->ticket_base, c->part_epoch, c->d_hist_sum, c->d_hist_cnt, c->ncells, (const RoundHdr *) nullptr, c->pack,
                  (const long long *) c->d_lamq, c->d_lamq_c, c->d_lcount, c->d_lcount_mapped);
        c->ticket_base += part_blk;
      }
      QR_LAUNCH(c, PH_HIST, round_kernel<B>, grid, kRoundThreads, fused_smem, c->d_tasks, k, fused_part, hist_blk, c->d_panels,
                c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, (uint32_t) c->F, c->npanels, c->d_hist_sum,
                c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, c->d_part_status, c->d_ticket, c->ticket_base,
                c->part_epoch, rc, c->p.minleafsupport, c->d_qexp, c->d_fbest_score, c->d_fbest_t, c->d_fbest_lc,
                c->d_totals, c->d_res_mapped, c->d_flags_mapped, c->round_id, c->pack);
      return QR_OK;
    }));
    c->ticket_base += grid;
    QR_TRY(wait_round_flags(c, k));
  } else {
  {
    PhaseTimer pt(c, PH_PARTITION);
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      if (onepass) {
        c->part_epoch++;
        QR_LAUNCH(c, PH_PARTITION, partition_onepass_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                  c->d_ids[0], c->d_ids[1], c->d_ids[0], c->d_ids[1], c->d_part_status, c->d_ticket,
                  c->ticket_base, c->part_epoch, c->d_hist_sum, c->d_hist_cnt, c->ncells, (const RoundHdr *) nullptr, c->pack,
                  (const long long *) c->d_lamq, c->d_lamq_c, c->d_lcount, c->d_lcount_mapped);
        c->ticket_base += part_blk;
      } else {
        QR_LAUNCH(c, PH_PARTITION, part_count_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                  c->d_ids[0], c->d_ids[1], c->d_blockcnt);
        QR_LAUNCH(c, PH_PARTITION, part_prefix_kernel, k, 256, 0, c->d_tasks, c->d_blockcnt, c->d_lcount);
        QR_LAUNCH(c, PH_PARTITION, part_scatter_kernel<B>, part_blk, 256, 0, c->d_tasks, k, c->d_panels, c->N,
                  c->d_ids[0], c->d_ids[1], c->d_ids[0], c->d_ids[1], c->d_blockcnt, c->d_lcount);
      }
      return QR_OK;
    }));
  }
  QR_TRACE_MARK(c);
  if (build_child_hists) {
    QR_TRY(launch_hist_and_scan(c, k, hist_blk, false, onepass, (double) built_total));
  } else if (c->comm) {
    // last oblivious level: no split scan follows whose flags could be waited for
    if (c->part_3pass)
      QR_CUDA(cudaMemcpyAsync(c->h_lcount, c->d_lcount, k * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
  }
  }
  c->pack.n = 0;
  for (uint32_t j = 0; j < k; ++j) {
    const int i = S[j];
    const NodeTask &t = c->h_tasks[j];
    const HostNode nd = c->nodes[i];
    const uint32_t lc_local = c->comm ? c->h_lcount[j] : (uint32_t) nd.res.lcount;
    HostNode L, R;
    L.lo = nd.lo; L.n = lc_local; L.buf = (int) t.dst;
    R.lo = nd.lo + lc_local; R.n = nd.n - lc_local; R.buf = (int) t.dst;
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
  c->nrounds++;
  return QR_OK;
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

static int fit_leaves(qr_ctx *c) {
  const size_t nl = c->leaves.size();
  PhaseTimer pt(c, PH_LEAF);
  uint32_t blk = 0;
  for (size_t k = 0; k < nl; ++k) {
    const HostNode &nd = c->nodes[c->leaves[k]];
    c->h_segs[k] = LeafSeg{nd.lo, nd.n, (uint32_t) nd.buf, blk};
    blk += (nd.n + kLeafItems - 1) / kLeafItems;
  }
  QR_CUDA(cudaMemcpyAsync(c->d_segs, c->h_segs, nl * sizeof(LeafSeg), cudaMemcpyHostToDevice, c->stream));
  const double *w = c->lambda ? c->d_weight : nullptr;
  if (c->exact && !c->comm) {
    QR_LAUNCH(c, PH_LEAF, leaf_exact_kernel, (unsigned) nl, 32, 0, c->d_segs, c->d_ids[0], c->d_ids[1], c->d_lambda,
              w, c->d_leafval, c->d_leaf_of_doc);
  } else {
    if (blk > 0)
      QR_LAUNCH(c, PH_LEAF, leaf_partial_kernel, blk, 256, 0, c->d_segs, (uint32_t) nl, c->d_ids[0], c->d_ids[1],
                c->d_lambda, w, c->d_leaf_partials, c->d_leaf_of_doc, (const RoundHdr *) nullptr);
    QR_LAUNCH(c, PH_LEAF, leaf_final_kernel, (unsigned) ((nl + 63) / 64), 64, 0, c->d_segs, (uint32_t) nl,
              c->d_leaf_partials, c->lambda, c->d_leafsum, c->d_leafval, (const RoundHdr *) nullptr);
    if (c->comm) QR_TRY(comm_leaf_values(c, (uint32_t) nl));
  }
  QR_CUDA(cudaMemcpyAsync(c->h_leafval, c->d_leafval, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  QR_CUDA(cudaStreamSynchronize(c->stream));
  for (size_t k = 0; k < nl; ++k) c->nodes[c->leaves[k]].value = c->h_leafval[k];
  return QR_OK;
}


// ------------------------------------------------------------------------------------------
// Device-driven leaf-wise growth (qr_grow.cuh): the host only keeps launches queued.
// ------------------------------------------------------------------------------------------
static int enqueue_device_round(qr_ctx *c, uint32_t round, bool root) {
  const uint32_t F = (uint32_t) c->F;
  const uint32_t mt = c->max_tasks;
  RoundHdr *hdr = c->d_hdr + (round & 1u), *next_hdr = c->d_hdr + ((round + 1u) & 1u);
  NodeTask *tasks = c->d_tasks + (size_t) (round & 1u) * mt, *next_tasks = c->d_tasks + (size_t) ((round + 1u) & 1u) * mt;
  const size_t smem = (size_t) c->fpp * c->max_thr * 12;
  const bool use_smem = smem <= 200 * 1024;
  const uint32_t want_slices = c->h_grow->want_slices;
  if (root) {
    const uint32_t slices = std::max<uint32_t>(1, ((uint32_t) c->N + c->root_dpb - 1) / c->root_dpb);
    QR_LAUNCH(c, PH_HIST, prep_slots_kernel, dim3(std::max<uint32_t>(1, std::min<uint32_t>(32, (c->ncells + 1023) / 1024)), 1),
              256, 0, tasks, c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_root_cnt);
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      if (use_smem)
        QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, true, false>), dim3(slices, c->npanels), kHistThreads, smem, tasks, 1u,
                  c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                  c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) hdr, c->pack, (const long long *) c->d_lamq_c);
      else
        QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, false, false>), dim3(slices, c->npanels), kHistThreads, 0, tasks, 1u,
                  c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                  c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) hdr, c->pack, (const long long *) c->d_lamq_c);
      return QR_OK;
    }));
  } else {
    const uint32_t part_grid = (uint32_t) ((c->N + kPartItems - 1) / kPartItems) + mt;
    const uint32_t hist_grid = want_slices + mt;
    c->part_epoch++;
    QR_TRY(dispatch_bins(c, [&](auto tag) -> int {
      using B = decltype(tag);
      QR_LAUNCH(c, PH_PARTITION, partition_onepass_kernel<B>, part_grid, 256, 0, tasks, 0u, c->d_panels, c->N,
                c->d_ids[0], c->d_ids[1], c->d_ids[0], c->d_ids[1], c->d_part_status, c->d_ticket, 0u, c->part_epoch,
                c->d_hist_sum, c->d_hist_cnt, c->ncells, (const RoundHdr *) hdr, c->pack, (const long long *) c->d_lamq, c->d_lamq_c,
                c->d_lcount, (uint32_t *) nullptr);
      if (use_smem)
        QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, true, true>), dim3(hist_grid, c->npanels), kHistThreads, smem, tasks, 0u,
                  c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                  c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) hdr, c->pack, (const long long *) c->d_lamq_c);
      else
        QR_LAUNCH(c, PH_HIST, (hist_limb_kernel<B, false, true>), dim3(hist_grid, c->npanels), kHistThreads, 0, tasks, 0u,
                  c->d_lcount, c->d_panels, c->N, c->d_ids[0], c->d_ids[1], c->d_lamq, c->d_thr_off, F, c->d_hist_sum,
                  c->d_hist_cnt, c->ncells, c->d_sq128, c->max_thr, (const RoundHdr *) hdr, c->pack, (const long long *) c->d_lamq_c);
      return QR_OK;
    }));
  }
  QR_LAUNCH(c, PH_SCAN, finalize_kernel<false>, dim3(fin_blocks(F), root ? 1u : mt), kFinWarps * 32, 0, tasks,
            c->d_hist_sum, c->d_hist_cnt, c->ncells, c->d_thr_off, F, c->p.minleafsupport, c->d_qexp, c->d_fbest_score,
            c->d_fbest_t, c->d_fbest_lc, c->d_totals, c->d_sq128, c->d_partials, c->d_task_done, c->d_res,
            (volatile uint32_t *) nullptr, 0u, (const RoundHdr *) hdr, c->pack, PeerView{});
  QR_LAUNCH(c, PH_SCAN, grow_step_kernel, 1, kGrowThreads, c->grow_smem, c->d_grow, hdr, next_hdr, tasks, next_tasks,
            c->d_res, c->d_ticket, c->d_segs, c->d_grow_out);
  return QR_OK;
}

static int fit_leafwise_device(qr_ctx *c, bool want_nodes) {
  constexpr uint32_t kAhead = 2;   // rounds kept queued beyond the last completed grow_step
  GrowOut *go = c->h_grow_out;
  go->steps = 0; go->done = 0; go->error = 0;
  std::atomic_thread_fence(std::memory_order_seq_cst);
  c->root_dpb = pick_hist_dpb(c, c->N);
  QR_LAUNCH(c, PH_HIST, grow_init_kernel, 1, 1, 0, c->d_grow, c->d_hdr, c->d_tasks, (uint32_t) c->N, c->root_dpb, c->d_ticket);
  QR_TRY(enqueue_device_round(c, 0, true));
  uint32_t enq = 0;   // growth rounds enqueued (the round after grow_step number enq + 1)
  uint64_t spins = 0;
  for (;;) {
    while (!go->done && enq >= go->steps + kAhead) {
      if ((++spins & 0xfffff) == 0) {
        cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) { set_error("tree growth failed: %s", cudaGetErrorString(e)); return QR_ECUDA; }
        if (e == cudaSuccess && !go->done && enq >= go->steps + kAhead) {
          set_error("internal: growth rounds finished without progress (steps %u, enqueued %u)", go->steps, enq);
          return QR_ECUDA;
        }
      }
    }
    if (go->done) break;
    ++enq;
    QR_TRY(enqueue_device_round(c, enq, false));
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (go->error) {
    set_error(go->error == 1 ? "internal: histogram pool exhausted" : go->error == 2 ? "internal: node table full"
                                                                                    : "internal: histogram slice table full");
    return QR_ECUDA;
  }
  c->rho = go->rho; c->sigma = go->sigma; c->beta = go->beta; c->nsplits = go->nsplits; c->nrounds = go->nrounds;
  // leaf outputs (rt.cc:165-207): the segments were written by the last grow_step
  const RoundHdr *hdr = c->d_hdr + ((go->steps) & 1u);   // header written by the last step
  {
    PhaseTimer pt(c, PH_LEAF);
    const size_t maxleaves = std::max<size_t>(c->p.nleaves, 1);
    const uint32_t leaf_grid = (uint32_t) ((c->N + kLeafItems - 1) / kLeafItems + maxleaves);
    const double *w = c->lambda ? c->d_weight : nullptr;
    QR_LAUNCH(c, PH_LEAF, leaf_partial_kernel, leaf_grid, 256, 0, c->d_segs, 0u, c->d_ids[0], c->d_ids[1], c->d_lambda, w,
              c->d_leaf_partials, c->d_leaf_of_doc, hdr);
    QR_LAUNCH(c, PH_LEAF, leaf_final_kernel, (unsigned) ((maxleaves + 63) / 64), 64, 0, c->d_segs, 0u, c->d_leaf_partials,
              c->lambda, c->d_leafsum, c->d_leafval, hdr);
  }
  c->nodes.clear();
  c->leaves.clear();
  if (want_nodes) {   // the caller wants the tree on the host
    const uint32_t nn = go->nnodes, nl = go->nleaves;
    QR_CUDA(cudaMemcpyAsync(c->h_nodes, c->d_nodes, nn * sizeof(DevNode), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaMemcpyAsync(c->h_leafval, c->d_leafval, nl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    QR_CUDA(cudaStreamSynchronize(c->stream));
    c->nodes.resize(nn);
    for (uint32_t i = 0; i < nn; ++i) {
      const DevNode &d = c->h_nodes[i];
      HostNode &h = c->nodes[i];
      h.lo = d.lo; h.n = d.n; h.buf = d.buf; h.hist = -1; h.left = d.left; h.right = d.right;
      h.expanded = d.expanded != 0; h.pushed = d.pushed != 0; h.res = d.res;
    }
    collect_leaves(c, 0);
    for (size_t k = 0; k < c->leaves.size(); ++k) c->nodes[c->leaves[k]].value = c->h_leafval[k];
  }
  return QR_OK;
}

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

  QR_TRY(prepare_fixed_point(c));
  if (c->device_growth && !c->profiling) {
    QR_TRY(fit_leafwise_device(c, out != nullptr));
    c->has_tree = true;
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
  if (c->device_growth) {   // the device-driven path leaves the partition ticket counter at an arbitrary value
    QR_CUDA(cudaMemsetAsync(c->d_ticket, 0, sizeof(uint32_t), c->stream));
    c->ticket_base = 0;
  }
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
      if (getenv("QR_TRACE_ROUNDS")) fprintf(stderr, "[trace] round %2zu: partition %6.1f hist %6.1f finalize %6.1f gap-to-next %6.1f us\n", r, a * 1e3, b * 1e3, d * 1e3, g * 1e3);
    }
    fprintf(stderr, "[trace] %zu rounds: partition %.0f hist %.0f finalize %.0f gaps %.0f us\n", nr, tp * 1e3, th * 1e3, tf * 1e3, tg * 1e3);
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
