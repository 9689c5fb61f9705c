#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: LambdaMART trees/sec; config 2).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one boosting iteration of Mart::learn's loop (reference mart.cc:331-347):
pseudo-responses -> root histogram -> tree fit -> leaf outputs -> score update -> NDCG@10, on the
synthetic config-2 workload (1M docs x 136 features x 10k queries, 64 leaves).  With N > 1 the
SAME 1M documents are sharded by query over the ranks, one process per GPU (torchrun; strong
scaling, as BASELINE.json's north_star asks: trees/sec on the 1M-doc input at 1/2/4/8 GPUs), and
the per-bin histograms of every growth round are all-reduced over NVLink peer memory (fused into
the split-scan kernel for small rounds; NCCL for the per-tree scalars).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each field means.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PARITY_TREES = 5   # trees of the job whose digest every bench line carries
WORKLOAD = dict(n_docs=1_000_000, n_features=136, n_queries=10_000, leaves=64, cutoff=10,
                shrinkage=0.1, nthresholds=0, minls=1, seed=20260102)
METRIC = "lambdamart_trees_per_sec"
UNIT = "trees/s"


def workload_string(w):
    """config.workload, the same string in both arms (ours and --impl reference)"""
    return ("LambdaMART 64 leaves, synthetic %d docs x %d feat x %d queries, NDCG@10, 256-level features (u8 bins), "
            "shrinkage 0.1 (BASELINE.json configs[1])" % (w["n_docs"], w["n_features"], w["n_queries"]))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            c = [v.strip() for v in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def tree_digest(trees):
    """SHA-1 over (split feature, threshold index, node size, leaf output bits) of the given trees, in pre-order.
    Histogram sums, leaf sums and the NDCG mean are exact integers in the benchmarked mode, so the digest — the whole
    model, leaf outputs included — is the same on 1, 2, 4 and 8 GPUs."""
    import hashlib
    h = hashlib.sha1()
    for t in trees:
        for k in ("feature", "threshold_idx", "count"):
            h.update(np.ascontiguousarray(t[k]).astype(np.int64).tobytes())
        leaves = np.asarray(t["feature"]) < 0
        h.update(np.ascontiguousarray(np.asarray(t["value"], np.float64)[leaves]).tobytes())
    return h.hexdigest()


def oracle_first_trees(x, labels, qoff, ntrees, w=WORKLOAD):
    """The first trees of the same job grown by the unmodified reference (oracle/_ref) on the host."""
    from oracle import pyref
    if not pyref.available():
        return None
    pyref.set_threads(cpu_threads())
    out = []
    with pyref.RefSession("LAMBDAMART", x, labels, qoff, ntrees=ntrees + 1, shrinkage=w["shrinkage"],
                          nthresholds=w["nthresholds"], nleaves=w["leaves"], minleafsupport=w["minls"],
                          cutoff=w["cutoff"]) as s:
        s.init()
        for m in range(ntrees):
            s.compute_pseudoresponses()
            s.fit_tree(True)
            out.append(s.tree(m))
    return out


def _same_tree(a, b):
    return bool(all(len(a[k]) == len(b[k]) and np.array_equal(np.asarray(a[k]).astype(np.int64), np.asarray(b[k]).astype(np.int64))
                    for k in ("feature", "threshold_idx", "count")))


def stagewise_vs_reference(x, labels, qoff, device, first_trees, ntrees=2, w=WORKLOAD):
    from oracle import pyref
    from quickrank_b200 import api
    if not pyref.available():
        return {"kind": "oracle/_ref not built", "trees_compared": 0}
    pyref.set_threads(cpu_threads())
    out = {"kind": "oracle/_ref (unmodified reference sources)", "trees_compared": ntrees,
           "how": "every tree fitted from the reference's pseudo-responses at the reference's scores"}
    kw = dict(algo="LAMBDAMART", nleaves=w["leaves"], minleafsupport=w["minls"], nthresholds=w["nthresholds"],
              cutoff=w["cutoff"], shrinkage=w["shrinkage"], device=device)
    with pyref.RefSession("LAMBDAMART", x, labels, qoff, ntrees=ntrees + 1, shrinkage=w["shrinkage"],
                          nthresholds=w["nthresholds"], nleaves=w["leaves"], minleafsupport=w["minls"],
                          cutoff=w["cutoff"]) as ref, \
            api.Trainer(x, labels, qoff, hist_mode=api.HIST_FAST, **kw) as tf, \
            api.Trainer(x, labels, qoff, hist_mode=api.HIST_REFERENCE, **kw) as te:
        ref.init()
        eq_fast, eq_exact, leaf_fast, leaf_exact, lam_err, free = [], [], [], [], [], []
        for m in range(ntrees):
            scores = ref.get_scores()
            ref.compute_pseudoresponses()
            lam, wt = ref.get_gradients()
            ref.fit_tree(True)
            want = ref.tree(m)
            lv = np.asarray(want["feature"]) < 0
            free.append(_same_tree(first_trees[m], want))
            for tr_, eq, le in ((tf, eq_fast, leaf_fast), (te, eq_exact, leaf_exact)):
                tr_.set_scores(scores)
                if tr_ is tf:
                    tr_.compute_pseudoresponses()
                    gl, _gw = tr_.get_pseudoresponses()
                    lam_err.append(float(np.max(np.abs(gl - lam)) / np.max(np.abs(lam))))
                tr_.set_pseudoresponses(lam, wt)
                got = tr_.fit_regressor_on_gradient()
                same = _same_tree(got, want)
                eq.append(same)
                if same:
                    d = np.abs(np.asarray(got["value"])[lv] - np.asarray(want["value"])[lv])
                    le.append(float(np.max(d / np.maximum(np.abs(np.asarray(want["value"])[lv]), 1e-300))))
                tr_.update_modelscores()
    out["split_indices_and_counts_equal"] = {"QR_HIST_REFERENCE": eq_exact, "QR_HIST_FAST": eq_fast,
                                             "QR_HIST_FAST free-running (own pseudo-responses)": free}
    out["max_rel_leaf_output_error"] = {"QR_HIST_REFERENCE": max(leaf_exact) if leaf_exact else None,
                                        "QR_HIST_FAST": max(leaf_fast) if leaf_fast else None}
    out["max_rel_pseudo_response_error"] = max(lam_err) if lam_err else None
    return out


def make_shard(rank, world, w=WORKLOAD, keep_global=False):
    """The rank's contiguous range of queries of the one global dataset, balanced by document count
    (lambdas need whole queries: lambdamart.cc:71-151)."""
    from quickrank_b200 import synth
    x, labels, qoff = synth.make_dataset(w["n_docs"], w["n_features"], w["n_queries"], seed=w["seed"])
    if keep_global:
        make_shard.global_data = (x, labels, qoff)
    if world == 1:
        return x, labels, qoff
    from quickrank_b200.sharding import query_shards
    q0, q1 = query_shards(qoff, world)[rank]
    d0, d1 = int(qoff[q0]), int(qoff[q1])
    return (np.ascontiguousarray(x[d0:d1]), np.ascontiguousarray(labels[d0:d1]),
            (qoff[q0:q1 + 1] - qoff[q0]).astype(np.uint64))


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/hist_full_summary.json, written by scripts/ncu_summary.py)."""
    p = os.path.join(ROOT, "profiles", "r02_hist_full_summary.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "hist_full_summary.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("dram_bytes_per_launch"), "profiles/%s (%s)" % (os.path.basename(p), d.get("capture", ""))
    return None, "no ncu capture committed"


def shared_atomics_view(hk_docs, root_docs, hk_ms, n_features, ceiling=3.1e12):
    """The histogram launches against the limit that actually binds them: shared-memory
    atomic lane-operations per second.  A (document, feature) update costs two limb atomics, plus a count atomic
    on child launches (the root refresh copies its counts); the ceiling is what scripts/hist_mb2.cu sustains on
    B200 (16 lanes per clock per SM)."""
    child_docs = max(hk_docs - root_docs, 0.0)
    atomics = (root_docs * 2.0 + child_docs * 3.0) * n_features
    rate = atomics / max(hk_ms * 1e-3, 1e-12)
    return {"achieved": round(rate / 1e12, 3), "ceiling": round(ceiling / 1e12, 2), "unit": "T lane-atomics/s",
            "frac": round(rate / ceiling, 3),
            "ceiling_source": "scripts/hist_mb2.cu microbenchmark on B200 (DESIGN.md section 4); small launches "
                              "sit below it because of their fixed costs (shared-memory clear, flush)"}


def hist_bytes_per_tree(n, f, rho, bin_bytes=1):
    """Algorithmic bytes of the histogram kernels for one tree:
    root: every bin once + lambda once; children: bins + doc id + gathered lambda."""
    return n * (f * bin_bytes + 8) + rho * n * (f * bin_bytes + 4 + 8)


def tree_bytes(n, f, rho, sigma, leaves, cells):
    h = cells * 12
    return (28 * n + hist_bytes_per_tree(n, f, rho) + sigma * n * (1 + 4 + 4) + 6 * h * (leaves - 1)
            + 20 * n + 20 * n + 12 * n)


def run_ours(args):
    from quickrank_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    if api.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback on the hot path)")
    w = WORKLOAD
    x, labels, qoff = make_shard(rank, world, keep_global=(rank == 0))

    comm = None
    if world > 1:
        import torch
        idt = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = (bytes(idt.cpu().numpy().tobytes()), rank, world)
    t0 = time.perf_counter()
    tr = api.Trainer(x, labels, qoff, algo="LAMBDAMART", nleaves=w["leaves"], minleafsupport=w["minls"],
                     nthresholds=w["nthresholds"], cutoff=w["cutoff"], shrinkage=w["shrinkage"],
                     hist_mode=api.HIST_FAST, device=local_rank, comm=comm)
    init_s = time.perf_counter() - t0
    exchange = {"none": "none (one GPU)", "nccl": "NCCL all-reduces per growth round",
                "peer": "peer memory over NVLink (CUDA IPC): the all-reduce of a round's built histograms is fused "
                        "into the split-scan kernel (peer loads, one flag barrier) when (W-1)*nodes <= 4, else one "
                        "stand-alone in-place reduce-scatter + all-gather kernel; NCCL only per tree (scale, leaf "
                        "sums, NDCG)"}[tr.comm_transport()]

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    # The workload is a 1000-tree run and its trees are not alike: the first ~25 trees rank queries full of tied
    # scores (sequential sort replica) and need ~11 growth rounds; from tree ~150 on LambdaMART's gradients grow
    # chain-like trees of ~32 rounds, which is what 85% of the run consists of.  The timed region (exactly
    # --steps trees after --settle + --warmup untimed ones) therefore sits in that deep regime (--settle 300);
    # `windows` reports the same measurement early (tree 30) and late (tree 900) in the run, and `e2e` is the
    # whole job.  Clocks are sampled from the first tree on: the same workload runs before and inside the timed
    # region, so every sample is a sample under load.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    windows = []

    def timed_window(first_tree, ntrees):
        barrier()
        tr.timer_start()
        rr = rho = sigma = 0.0
        for _ in range(ntrees):
            tr.boost_iteration(want_tree=False, want_metric=True)
            r, s_, _ns = tr.last_tree_stats()
            rho += r
            sigma += s_
            rr += tr.last_tree_rounds()[0]
        ms_ = tr.timer_stop()
        barrier()
        if dist is not None:
            import torch
            t_ = torch.tensor([ms_], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms_ = float(t_.item())
        windows.append({"first_tree": first_tree, "trees": ntrees, "ms_per_tree": round(ms_ / ntrees, 4),
                        "trees_per_s": round(1000.0 * ntrees / ms_, 2), "growth_rounds_per_tree": round(rr / ntrees, 1)})
        return ms_, rho / ntrees, sigma / ntrees

    done = 0
    # parity: the first trees of the job, digest identical for every N (and equal to the reference's, below)
    first_trees = []
    for _ in range(PARITY_TREES):
        tree, _m = tr.boost_iteration(want_tree=True, want_metric=True)
        first_trees.append(tree)
        done += 1
    early = min(30, args.settle)
    while done < early:
        tr.boost_iteration(want_tree=False, want_metric=True)
        done += 1
    if args.settle >= early + args.steps:
        timed_window(done, args.steps)
        done += args.steps
    while done < args.settle:
        tr.boost_iteration(want_tree=False, want_metric=True)
        done += 1
    for _ in range(max(args.warmup, 3)):
        tr.boost_iteration(want_tree=False, want_metric=True)
        done += 1
    launches0 = tr.launch_count()
    first_timed = done
    ms, rho_mean, sigma_mean = timed_window(done, args.steps)
    done += args.steps
    launches = tr.launch_count() - launches0
    main_window = windows[-1]
    if args.late_window and done + args.steps <= args.late_window + args.steps:
        while done < args.late_window:
            tr.boost_iteration(want_tree=False, want_metric=True)
            done += 1
        timed_window(done, args.steps)
        done += args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = 1000.0 / ms_per_step  # trees/s of the whole job (all ranks grow the same tree)
    rho_sum, sigma_sum = rho_mean * args.steps, sigma_mean * args.steps

    # ---- end to end: the whole job through the C ABI, starting from HOST buffers ----
    # A new context is created from the host feature matrix (host->device copy of the dataset, threshold
    # extraction and binning = Mart::init), then `e2e_trees` boosting iterations run through the hook-level
    # entry points, each copying the fitted tree and the metric back to host buffers — what quicklearn does
    # for a whole training run.  Everything is inside the timed region (wall clock around the calls, which
    # end with the device-to-host copy of the last metric).
    e2e_trees = args.e2e_trees
    d2h = 0
    barrier()
    t0 = time.perf_counter()
    tr2 = api.Trainer(x, labels, qoff, algo="LAMBDAMART", nleaves=w["leaves"], minleafsupport=w["minls"],
                      nthresholds=w["nthresholds"], cutoff=w["cutoff"], shrinkage=w["shrinkage"],
                      hist_mode=api.HIST_FAST, device=local_rank, comm=comm)   # same id: the process's NCCL communicator is reused
    e2e_init_s = time.perf_counter() - t0
    for _ in range(e2e_trees):
        tr2.compute_pseudoresponses()
        tree = tr2.fit_regressor_on_gradient(want_tree=True)      # flat tree copied to host arrays
        tr2.update_modelscores()
        _m = tr2.evaluate_dataset()                               # metric read back
        d2h += sum(tree[k].nbytes for k in tree if hasattr(tree[k], "nbytes")) + 8
    e2e_s = time.perf_counter() - t0
    tr2.close()
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_trees / e2e_s
    h2d_job = int(x.nbytes + labels.nbytes + qoff.nbytes)

    # ---- roofline of the dominant kernel (hist_limb_kernel), CUDA events around every launch ----
    prof_steps = min(args.steps, 10)
    tr.set_profiling(True)
    tr.phase_times(reset=True)
    tr.hist_kernel_time(reset=True)
    for _ in range(prof_steps):
        tr.boost_iteration(want_tree=False, want_metric=True)
    pms, pln = tr.phase_times()
    hk_ms, hk_launches, hk_docs = tr.hist_kernel_time()
    tr.set_profiling(False)
    peak, peak_src = load_peaks()
    n, f = len(labels), w["n_features"]
    # algorithmic bytes of one launch: every accumulated document contributes its
    # F bin bytes + its fixed-point pseudo-response (8) [+ its id (4) when the list is gathered]
    root_docs = float(n) * prof_steps
    alg_bytes = hk_docs * (f * 1 + 8) + (hk_docs - root_docs) * 4
    bytes_per_launch = alg_bytes / max(hk_launches, 1)
    us_per_launch = hk_ms * 1e3 / max(hk_launches, 1)
    achieved = alg_bytes / (hk_ms * 1e-3) / 1e9
    traffic, traffic_src = load_traffic()
    roofline = {"bound": "hbm", "kernel": "hist_limb_kernel (histogram build, root + child nodes)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(bytes_per_launch), "us_per_launch": round(us_per_launch, 2),
                "launches_per_tree": round(hk_launches / prof_steps, 2),
                "docs_per_launch": int(hk_docs / max(hk_launches, 1)),
                "kernel_ms_per_tree": round(hk_ms / prof_steps, 4),
                "kernel_share_of_step": round(hk_ms / prof_steps / ms_per_step, 3),
                "phase_ms_per_tree_with_sync": {k: round(v / prof_steps, 4) for k, v in pms.items()},
                "whole_tree_gbs": round(tree_bytes(n, f, rho_sum / args.steps, sigma_sum / args.steps,
                                                   w["leaves"], 0) / (ms_per_step * 1e-3) / 1e9, 1),
                "note": "bound in practice by the shared-memory atomic issue rate (16 lanes/clk/SM), "
                        "see DESIGN.md section 4"}

    try:   # an explanatory extra, never allowed to cost the line
        roofline["shared_atomics"] = shared_atomics_view(hk_docs, root_docs, hk_ms, f)
    except Exception as e:   # noqa: BLE001
        roofline["shared_atomics"] = {"error": str(e)}

    tr.close()

    # ---- parity block: digest of the job's first trees + the same trees from the unmodified reference ----
    parity = {"trees": PARITY_TREES, "digest_sha1": tree_digest(first_trees),
              "digest_of": "(split feature, threshold index, node size, leaf output bits) of the first %d trees, pre-order; "
                           "fixed-point histograms, leaf sums and NDCG mean make the model bit-identical for every "
                           "number of GPUs" % PARITY_TREES}
    reference_mode = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # Against the UNMODIFIED reference (oracle/_ref), stage by stage: each tree is fitted from the reference's own
        # pseudo-responses at the reference's own scores, so nothing but the tree fit is compared.  REFERENCE mode must
        # reproduce the reference bit for bit; FAST mode (the benchmarked one) may differ only inside exact-arithmetic
        # ties, which tests/test_gpu_parity_baseline_shapes.py audits node by node on this very dataset.
        parity["reference"] = stagewise_vs_reference(x, labels, qoff, local_rank, first_trees)
    # ---- the bit-exact mode (QR_HIST_REFERENCE: reference accumulation order, single GPU) next to FAST ----
    if world == 1 and args.reference_mode_trees > 0:
        tre = api.Trainer(x, labels, qoff, algo="LAMBDAMART", nleaves=w["leaves"], minleafsupport=w["minls"],
                          nthresholds=w["nthresholds"], cutoff=w["cutoff"], shrinkage=w["shrinkage"],
                          hist_mode=api.HIST_REFERENCE, device=local_rank)
        for i in range(2):
            tre.boost_iteration(want_tree=False, want_metric=True)
        tre.timer_start()
        for _ in range(args.reference_mode_trees):
            tre.boost_iteration(want_tree=False, want_metric=True)
        ems = tre.timer_stop()
        tre.close()
        reference_mode = {"hist_mode": "QR_HIST_REFERENCE (per-bin FP64 sums in the reference's order: bit-exact trees)",
                          "trees_per_s": round(1000.0 * args.reference_mode_trees / ems, 2),
                          "ms_per_tree": round(ems / args.reference_mode_trees, 3),
                          "trees_timed": args.reference_mode_trees, "first_tree_timed": 2}
    scoring = run_scoring(args, x, rank, world, local_rank, dist, barrier)

    out = None
    if rank == 0:
        cpu = cpu_baseline(args, quick=True) if world == 1 and not args.no_cpu_baseline else None
        if cpu:
            scoring["cpu_baseline"] = reference_scoring(x, 200_000)
        out = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(w),
                       "docs_per_gpu": int(len(labels)), "global_docs": w["n_docs"],
                       "hist_mode": "fixed-point int64 (FAST)", "parallelism": "query-sharded dp%d" % world,
                       "histogram_exchange": exchange,
                       "trees_before_timed_region": first_timed,
                       "growth_rounds_per_tree_in_timed_region": main_window["growth_rounds_per_tree"],
                       "l2": "inputs larger than L2 (136 MB bin matrix + 40 MB state per step)"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d_job / e2e_trees), "d2h_bytes_per_step": int(d2h / e2e_trees),
                    "trees": e2e_trees, "job_s": round(e2e_s, 3), "init_s": round(e2e_init_s, 3),
                    "h2d_bytes_job": h2d_job,
                    "note": "whole training job from host buffers: qr_ctx_create (dataset host->device, "
                            "thresholds, binning = Mart::init) + %d boosting iterations through the hook-level "
                            "C ABI, each copying the fitted tree and NDCG@10 to the host; the first ~25 trees "
                            "(tied scores, sequential sort replica) are inside" % e2e_trees
                            + ("" if world == 1 else "; the process's NCCL communicator (bootstrapped once, like the "
                               "CUDA context, when the first context was made) is reused; the CUDA IPC mapping of "
                               "the peers' histogram pools is inside")},
            "init_s": round(init_s, 3), "init_h2d_bytes": int(x.nbytes + labels.nbytes + qoff.nbytes),
            "gpu_launches": int(launches),
            "windows": windows,
            "parity": parity,
            "reference_mode": reference_mode,
            "roofline": roofline,
            "scoring": scoring,
        }
        if cpu:
            out["cpu_baseline"] = cpu
    if dist is not None:
        dist.destroy_process_group()
    return out


SCORING = dict(trees=1000, leaves=64, seed=11)


def scoring_ensemble(n_features):
    from quickrank_b200 import synth
    return synth.random_ensemble(SCORING["trees"], SCORING["leaves"], n_features, seed=SCORING["seed"])


def run_scoring(args, x, rank, world, local_rank, dist, barrier):
    """Second half of the headline metric: documents/s scored by a 1000-tree ensemble over this rank's
    shard of the same dataset (documents sharded, no collective)."""
    import torch
    from quickrank_b200 import api
    w = WORKLOAD
    trees, weights = scoring_ensemble(w["n_features"])
    sc = api.Scorer(trees, weights, w["n_features"], device=local_rank)
    n = x.shape[0]
    xp = torch.from_numpy(x).pin_memory()
    outp = torch.empty(n, dtype=torch.float64).pin_memory()
    xd = xp.cuda(non_blocking=False)
    outd = torch.empty(n, dtype=torch.float64, device="cuda")
    passes = max(args.steps, 5)
    for _ in range(3):
        sc.score_dataset_device(xd.data_ptr(), n, outd.data_ptr())
    sc.sync()
    l0 = sc.launch_count()
    barrier()
    sc.timer_start()
    for _ in range(passes):
        sc.score_dataset_device(xd.data_ptr(), n, outd.data_ptr())
    ms = sc.timer_stop() / passes
    barrier()
    launches = sc.launch_count() - l0
    # end to end: page-locked host rows in, host scores out, through qr_score_dataset
    xh, oh = xp.numpy(), outp.numpy()
    sc.score_dataset(xh, out=oh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        sc.score_dataset(xh, out=oh)
    e2e_ms = (time.perf_counter() - t0) / passes * 1e3
    barrier()
    same = bool(np.array_equal(oh, outd.cpu().numpy()))
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = (float(v) for v in t.tolist())
    sc.close()
    gdocs = w["n_docs"]
    peak, _src = load_peaks()
    alg = n * (w["n_features"] * 4 + 8)
    return {
        "metric": "ensemble_docs_per_sec", "value": round(gdocs / ms * 1e3, 1), "unit": "docs/s",
        "ms_per_pass": round(ms, 4), "passes": passes, "scaling": "strong",
        "doc_trees_per_sec": round(gdocs * SCORING["trees"] / ms * 1e3, 1),
        "config": {"workload": "%d-tree x %d-leaf random ensemble (synth.random_ensemble seed %d) over the same "
                               "%d docs x %d feat, documents sharded over ranks, no collective"
                               % (SCORING["trees"], SCORING["leaves"], SCORING["seed"], gdocs, w["n_features"]),
                   "docs_per_gpu": int(n)},
        "e2e": {"value": round(gdocs / e2e_ms * 1e3, 1), "unit": "docs/s", "h2d_bytes_per_step": int(x.nbytes),
                "d2h_bytes_per_step": int(n * 8), "host_equals_device_path": same},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": round(alg / (ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4), **score_traffic(),
                     "note": "the walk is bound by shared-memory wavefronts (one 8-byte node + one code per level at "
                             "data-dependent addresses), not by HBM: l1tex__data_pipe_lsu_wavefronts_mem_shared at 75% of "
                             "its peak, DRAM at 0.2% (profiles/r02_score_codes_counters.txt)"},
    }


def score_traffic():
    """DRAM bytes of one score_codes_kernel launch from the committed `ncu --set full` capture."""
    p = os.path.join(ROOT, "profiles", "r02_score_codes_full_summary.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"traffic": d.get("dram_bytes_per_launch"),
                "traffic_source": "profiles/r02_score_codes_full_summary.json (%s)" % d.get("capture", "")}
    return {"traffic": None, "traffic_source": "no ncu capture committed"}


def reference_scoring(x, sample):
    """The reference's LTR_Algorithm::score_dataset (OpenMP over documents) on a bounded sample of the
    same rows; model handed over as an XML file its own loader parses."""
    import tempfile
    from oracle import pyref
    from quickrank_b200 import modelxml
    nthreads = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthreads)   # (a launcher such as torchrun exports OMP_NUM_THREADS=1)
    if pyref.available():
        nthreads = pyref.set_threads(nthreads)
    trees, weights = scoring_ensemble(x.shape[1])
    xs = np.ascontiguousarray(x[:sample])
    if pyref.available():
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "ensemble.xml")
            modelxml.write_model(path, trees, weights)
            t0 = time.perf_counter()
            pyref.score_with_model(path, xs[:1])          # model load only
            load_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            pyref.score_with_model(path, xs)
            dt = max(time.perf_counter() - t0 - load_s, 1e-9)
        kind = "reference"
    else:
        from oracle import pyoracle as po
        po.score_dataset(trees[:2], weights[:2], xs[:16])
        t0 = time.perf_counter()
        po.score_dataset(trees, weights, xs)
        dt = time.perf_counter() - t0
        kind = "port"
    return {"value": round(len(xs) / dt, 1), "unit": "docs/s", "cores": nthreads, "kind": kind,
            "sample": "%d of the %d documents, %d trees (model load excluded)" % (len(xs), x.shape[0], SCORING["trees"])}


REF_BUILD = ("oracle/_ref: the unmodified reference sources, g++ -O3 -march=x86-64-v3 -fopenmp (the reference's Release "
             "flags are -O3 -march=native -fopenmp -D_GLIBCXX_PARALLEL: x86-64-v3 so that the library also runs on "
             "the GPU box, no libstdc++ parallel mode, which only affects std::sort of >= 1000 elements)")


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_steps(x, labels, qoff, warmup, steps, w=WORKLOAD):
    """Times the reference's own loop body on the host cores.  Uses the unmodified reference
    (oracle/_ref) when it is built, else the C restatement (oracle/qr_oracle.c)."""
    from oracle import pyref
    nthreads = cpu_threads()
    os.environ["OMP_NUM_THREADS"] = str(nthreads)   # (a launcher such as torchrun exports OMP_NUM_THREADS=1)
    if pyref.available():
        nthreads = pyref.set_threads(nthreads)       # what the OpenMP runtime will actually use
        s = pyref.RefSession("LAMBDAMART", x, labels, qoff, ntrees=warmup + steps + 1,
                             shrinkage=w["shrinkage"], nthresholds=w["nthresholds"], nleaves=w["leaves"],
                             minleafsupport=w["minls"], cutoff=w["cutoff"])
        t0 = time.perf_counter()
        s.init()
        init_s = time.perf_counter() - t0

        def step():
            s.compute_pseudoresponses()
            s.fit_tree(True)
            s.evaluate()
        kind = "reference"
    else:
        from oracle import pyoracle as po
        col = np.ascontiguousarray(x.T)
        t0 = time.perf_counter()
        ob = po.Binning(col, w["nthresholds"])
        init_s = time.perf_counter() - t0
        state = {"scores": np.zeros(len(labels))}

        def step():
            lam, wt = po.lambdas(state["scores"], labels, qoff, w["cutoff"])
            tree = ob.fit_tree(lam, wt, nleaves=w["leaves"], minls=w["minls"])
            state["scores"] = po.update_scores(tree, col, w["shrinkage"], state["scores"])
            po.ndcg_dataset(labels, state["scores"], qoff, w["cutoff"])
        kind = "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return dt / steps, init_s, kind, nthreads


def cpu_baseline(args, quick=True):
    """Bounded CPU sample of the same workload (rank 0, N=1 only)."""
    w = WORKLOAD
    from quickrank_b200 import synth
    n = w["n_docs"]
    x, labels, qoff = synth.make_dataset(n, w["n_features"], w["n_queries"], seed=w["seed"])
    sec_per_tree, init_s, kind, nthreads = reference_steps(x, labels, qoff, 1, 3)
    return {"value": round(1.0 / sec_per_tree, 4), "unit": UNIT, "cores": nthreads, "omp_max_threads": nthreads,
            "kind": kind, "build": REF_BUILD if kind == "reference" else "oracle/qr_oracle.c (restatement)",
            "sample": "full config-2 workload (%d docs), 1 warm-up + 3 timed boosting iterations; init "
                      "(transpose, argsort, binning) %.1f s excluded as in the reference's own Training Time"
                      % (n, init_s)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    w = WORKLOAD
    from quickrank_b200 import synth
    x, labels, qoff = synth.make_dataset(w["n_docs"], w["n_features"], w["n_queries"], seed=w["seed"])
    warm = min(max(args.warmup, 1), 10)   # (a reference iteration costs ~0.3 s on 16 cores)
    steps = min(args.steps, 20)
    sec_per_tree, init_s, kind, nthreads = reference_steps(x, labels, qoff, warm, steps)
    value = 1.0 / sec_per_tree
    ref_sc = reference_scoring(x, 200_000)
    return {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
        "ms_per_step": round(sec_per_tree * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(w), "where": "host CPU (the reference's OpenMP path), full workload"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": nthreads, "omp_max_threads": nthreads,
                         "kind": kind, "build": REF_BUILD if kind == "reference" else "oracle/qr_oracle.c (restatement)",
                         "sample": "full workload, %d timed iterations (capped at 20), init %.1f s excluded" % (steps, init_s)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "scoring": {"metric": "ensemble_docs_per_sec", "value": ref_sc["value"], "unit": "docs/s", "cpu_baseline": ref_sc},
    }


# ----------------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations (`--config 3|4`).  The default line stays configs[1] (config 2).
# ----------------------------------------------------------------------------------------------------------
CONFIG4 = dict(n_docs=4_000_000, n_features=220, n_queries=40_000, depth=6, cutoff=10, shrinkage=0.1, nthresholds=0,
               minls=1, seed=20260104, trees=2000)
CONFIG3 = dict(n_docs=10_000_000, n_features=700, trees=5000, leaves=64, seed=13, slice_docs=1_000_000)


def _dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return rank, world, local_rank, dist, barrier, maxr


def run_config4(args):
    """BASELINE.json configs[3]: oblivious LambdaMART depth 6 on 4M docs x 220 features x 40k queries, documents
    sharded by query over the ranks, the per-level histograms of every node all-reduced over NVLink peer memory."""
    from quickrank_b200 import api, synth
    rank, world, local_rank, dist, barrier, maxr = _dist_setup()
    w = dict(CONFIG4)
    if args.docs:
        w["n_docs"], w["n_queries"] = args.docs, max(1, args.docs // 100)
    x, labels, qoff = synth.make_dataset(w["n_docs"], w["n_features"], w["n_queries"], seed=w["seed"])
    gx = x[:400_000] if rank == 0 else None   # the reference leg's bounded sample (whole queries: see below)
    if world > 1:
        from quickrank_b200.sharding import query_shards
        q0, q1 = query_shards(qoff, world)[rank]
        d0, d1 = int(qoff[q0]), int(qoff[q1])
        gq = qoff
        gl = labels
        x, labels, qoff = (np.ascontiguousarray(x[d0:d1]), np.ascontiguousarray(labels[d0:d1]),
                           (qoff[q0:q1 + 1] - qoff[q0]).astype(np.uint64))
    else:
        gq, gl = qoff, labels
    comm = None
    if world > 1:
        import torch
        idt = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = (bytes(idt.cpu().numpy().tobytes()), rank, world)
    kw = dict(algo="OBVLAMBDAMART", treedepth=w["depth"], minleafsupport=w["minls"], nthresholds=w["nthresholds"],
              cutoff=w["cutoff"], shrinkage=w["shrinkage"], hist_mode=api.HIST_FAST, device=local_rank, comm=comm)
    tr = api.Trainer(x, labels, qoff, **kw)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    first = []
    for _ in range(3):
        t_, _m = tr.boost_iteration(want_tree=True, want_metric=True)
        first.append(t_)
    for _ in range(max(args.warmup, 3)):
        tr.boost_iteration(want_tree=False, want_metric=True)
    l0 = tr.launch_count()
    barrier()
    tr.timer_start()
    for _ in range(args.steps):
        tr.boost_iteration(want_tree=False, want_metric=True)
    ms = maxr(tr.timer_stop())
    barrier()
    launches = tr.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # roofline of the histogram kernel, timed on the device clock
    prof = min(args.steps, 5)
    tr.set_profiling(True)
    tr.hist_kernel_time(reset=True)
    for _ in range(prof):
        tr.boost_iteration(want_tree=False, want_metric=True)
    hk_ms, hk_launches, hk_docs = tr.hist_kernel_time()
    tr.set_profiling(False)
    tr.close()
    # end to end: context from host buffers + e2e trees, tree and metric back to the host every iteration
    e2e_trees = min(args.e2e_trees, 100)
    barrier()
    t0 = time.perf_counter()
    tr2 = api.Trainer(x, labels, qoff, **kw)
    d2h = 0
    for _ in range(e2e_trees):
        tr2.compute_pseudoresponses()
        tree = tr2.fit_regressor_on_gradient(want_tree=True)
        tr2.update_modelscores()
        tr2.evaluate_dataset()
        d2h += sum(tree[k].nbytes for k in tree if hasattr(tree[k], "nbytes")) + 8
    e2e_s = maxr(time.perf_counter() - t0)
    tr2.close()
    peak, peak_src = load_peaks()
    f = w["n_features"]
    n_local = len(labels)
    alg = hk_docs * (f + 8) + max(hk_docs - n_local * prof, 0) * 4
    ms_per_step = ms / args.steps
    out = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = reference_config4(gx, gl, gq, w)
        out = {"metric": "obv_lambdamart_trees_per_sec", "value": round(1000.0 / ms_per_step, 3), "unit": UNIT,
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "Oblivious LambdaMART depth %d, synthetic %d docs x %d feat x %d queries, NDCG@10 "
                                      "(BASELINE.json configs[3])" % (w["depth"], w["n_docs"], f, w["n_queries"]),
                          "docs_per_gpu": int(n_local), "parallelism": "query-sharded dp%d" % world,
                          "l2": "inputs larger than L2 (%d MB bin matrix per GPU)" % (n_local * ((f + 15) // 16) * 16 >> 20)},
               "clocks": clocks,
               "e2e": {"value": round(e2e_trees / e2e_s, 3), "unit": UNIT, "trees": e2e_trees, "job_s": round(e2e_s, 3),
                       "h2d_bytes_per_step": int((x.nbytes + labels.nbytes + qoff.nbytes) / e2e_trees),
                       "d2h_bytes_per_step": int(d2h / e2e_trees)},
               "gpu_launches": int(launches),
               "parity": {"trees": 3, "digest_sha1": tree_digest(first)},
               "roofline": {"bound": "hbm", "kernel": "hist_limb_kernel", "achieved": round(alg / (hk_ms * 1e-3) / 1e9, 1),
                            "peak": peak, "unit": "GB/s", "frac": round(alg / (hk_ms * 1e-3) / 1e9 / peak, 4), "traffic": None,
                            "peak_source": peak_src, "us_per_launch": round(hk_ms * 1e3 / max(hk_launches, 1), 2),
                            "kernel_ms_per_tree": round(hk_ms / prof, 4),
                            "kernel_share_of_step": round(hk_ms / prof / ms_per_step, 3),
                            "note": "bound by the shared-memory atomic rate (16 lanes/clk/SM), DESIGN.md section 4"}}
        if cpu:
            out["cpu_baseline"] = cpu
    if dist is not None:
        dist.destroy_process_group()
    return out


def reference_config4(x, labels, qoff, w):
    """The reference's OBVLAMBDAMART on a bounded sample (the first whole queries within 400k documents)."""
    from oracle import pyref
    if not pyref.available():
        return None
    nthreads = pyref.set_threads(cpu_threads())
    q = int(np.searchsorted(qoff, len(x), side="right") - 1)
    n = int(qoff[q])
    xs, ls, qs = np.ascontiguousarray(x[:n]), np.ascontiguousarray(labels[:n]), np.ascontiguousarray(qoff[:q + 1])
    with pyref.RefSession("OBVLAMBDAMART", xs, ls, qs, ntrees=4, shrinkage=w["shrinkage"], nthresholds=w["nthresholds"],
                          treedepth=w["depth"], minleafsupport=w["minls"], cutoff=w["cutoff"]) as s:
        s.init()
        s.compute_pseudoresponses(); s.fit_tree(True); s.evaluate()
        t0 = time.perf_counter()
        for _ in range(2):
            s.compute_pseudoresponses(); s.fit_tree(True); s.evaluate()
        dt = (time.perf_counter() - t0) / 2
    # the reference's cost per tree is linear in the documents: scaled to the full workload
    full = dt * w["n_docs"] / n
    return {"value": round(1.0 / full, 4), "unit": UNIT, "cores": nthreads, "omp_max_threads": nthreads, "kind": "reference",
            "build": REF_BUILD,
            "sample": "%d of the %d documents (%d whole queries), 1 warm-up + 2 timed boosting iterations at %.2f s each, "
                      "scaled linearly to the full workload" % (n, w["n_docs"], q, dt)}


def run_config3(args):
    """BASELINE.json configs[2]: a 5000-tree x 64-leaf ensemble over 10M docs x 700 features, documents sharded over
    the ranks, no collective.  The documents stream through qr_score_dataset from page-locked host memory in slices
    (upload of slice k+1 on a copy stream under the encode + walk of slice k); synthetic: each rank streams its share
    of the 10M documents as repeated passes over one 1M-document buffer."""
    import torch
    from quickrank_b200 import api, synth
    rank, world, local_rank, dist, barrier, maxr = _dist_setup()
    w = dict(CONFIG3)
    if args.docs:
        w["n_docs"] = args.docs
    f = w["n_features"]
    trees, weights = synth.random_ensemble(w["trees"], w["leaves"], f, seed=w["seed"])
    local_docs = w["n_docs"] // world
    passes = max(1, -(-local_docs // w["slice_docs"]))          # slices of at most slice_docs documents ...
    slice_docs = local_docs // passes                            # ... that add up to the rank's share
    x, _l, _q = synth.make_dataset(slice_docs, f, max(1, slice_docs // 100), seed=20260103 + rank)
    sc = api.Scorer(trees, weights, f, device=local_rank)
    xp = torch.from_numpy(x).pin_memory()
    outp = torch.empty(slice_docs, dtype=torch.float64).pin_memory()
    xd = xp.cuda()
    outd = torch.empty(slice_docs, dtype=torch.float64, device="cuda")
    for _ in range(max(args.warmup, 3) if slice_docs <= 200_000 else 1):
        sc.score_dataset_device(xd.data_ptr(), slice_docs, outd.data_ptr())
    sc.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = sc.launch_count()
    barrier()
    sc.timer_start()
    for _ in range(passes):
        sc.score_dataset_device(xd.data_ptr(), slice_docs, outd.data_ptr())
    ms = maxr(sc.timer_stop())
    barrier()
    launches = sc.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    xh, oh = xp.numpy(), outp.numpy()
    sc.score_dataset(xh[:4096], out=oh[:4096])
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        sc.score_dataset(xh, out=oh)
    e2e_s = maxr(time.perf_counter() - t0)
    same = bool(np.array_equal(oh, outd.cpu().numpy()))
    sc.close()
    docs = passes * slice_docs * world
    peak, peak_src = load_peaks()
    alg = passes * slice_docs * (4 * f + 8)
    out = None
    if rank == 0:
        out = {"metric": "ensemble_docs_per_sec", "value": round(docs / (ms * 1e-3), 1), "unit": "docs/s", "n_gpus": world,
               "steps": passes, "warmup": 3, "ms_per_step": round(ms / passes, 3), "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": "%d-tree x %d-leaf random ensemble over %d docs x %d feat (BASELINE.json configs[2]), "
                                      "documents sharded over ranks, no collective; a step = one pass over a %d-document slice"
                                      % (w["trees"], w["leaves"], docs, f, slice_docs),
                          "docs_per_gpu": int(passes * slice_docs), "l2": "inputs larger than L2 (%d MB of rows per pass)" % (x.nbytes >> 20)},
               "doc_trees_per_sec": round(docs * w["trees"] / (ms * 1e-3), 1),
               "clocks": clocks,
               "e2e": {"value": round(docs / e2e_s, 1), "unit": "docs/s", "h2d_bytes_per_step": int(x.nbytes),
                       "d2h_bytes_per_step": int(slice_docs * 8), "host_equals_device_path": same,
                       "pcie_bound_docs_per_s_at_50GBs": round(50e9 / (4 * f) * world, 1)},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": "score_codes_kernel", "achieved": round(alg / (ms * 1e-3) / 1e9, 1),
                            "peak": peak, "unit": "GB/s", "frac": round(alg / (ms * 1e-3) / 1e9 / peak, 4), **score_traffic(),
                            "peak_source": peak_src,
                            "note": "the walk visits one 8-byte node per level and tree in shared memory: bound by "
                                    "shared-memory wavefronts (75% of their peak in the ncu capture), not HBM"}}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = reference_scoring_ensemble(xh, 20_000, trees, weights, w["trees"])
    if dist is not None:
        dist.destroy_process_group()
    return out


def reference_scoring_ensemble(x, sample, trees, weights, ntrees):
    import tempfile
    from oracle import pyref
    from quickrank_b200 import modelxml
    if not pyref.available():
        return None
    nthreads = pyref.set_threads(cpu_threads())
    xs = np.ascontiguousarray(x[:sample])
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "ensemble.xml")
        modelxml.write_model(path, trees, weights)
        t0 = time.perf_counter()
        pyref.score_with_model(path, xs[:1])
        load_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        pyref.score_with_model(path, xs)
        dt = max(time.perf_counter() - t0 - load_s, 1e-9)
    return {"value": round(len(xs) / dt, 1), "unit": "docs/s", "cores": nthreads, "omp_max_threads": nthreads,
            "kind": "reference", "build": REF_BUILD,
            "sample": "%d documents, %d trees (model load excluded)" % (len(xs), ntrees)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-trees", type=int, default=1000,
                    help="boosting iterations of the end-to-end job (BASELINE.json configs[1]: 1000 trees)")
    ap.add_argument("--settle", type=int, default=300,
                    help="untimed boosting iterations before the warm-up: the timed region sits in the deep-tree "
                         "regime most of the 1000-tree job consists of (see run_ours)")
    ap.add_argument("--late-window", type=int, default=900,
                    help="first tree of the extra timed window late in the run (0: none)")
    ap.add_argument("--reference-mode-trees", type=int, default=5,
                    help="trees timed in the bit-exact QR_HIST_REFERENCE mode (0: skip)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="BASELINE.json configuration: 2 (default, the headline), 3 (ensemble scoring 10M x 700), "
                         "4 (oblivious LambdaMART 4M x 220)")
    ap.add_argument("--docs", type=int, default=0, help="override the document count of --config 3|4")
    args = ap.parse_args()
    if args.config == 4 and args.impl == "ours":
        out = run_config4(args)
    elif args.config == 3 and args.impl == "ours":
        out = run_config3(args)
    else:
        out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
