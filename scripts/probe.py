"""Development probe: phase breakdown of the training loop at a given size (not a benchmark)."""
import argparse
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from quickrank_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--f", type=int, default=136)
ap.add_argument("--q", type=int, default=10000)
ap.add_argument("--leaves", type=int, default=64)
ap.add_argument("--trees", type=int, default=8)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--algo", default="LAMBDAMART")
ap.add_argument("--depth", type=int, default=6)
ap.add_argument("--settle", type=int, default=0)
a = ap.parse_args()

t0 = time.time()
x, l, off = synth.make_dataset(a.n, a.f, a.q, seed=20260102)
print("gen %.1fs" % (time.time() - t0), flush=True)
t0 = time.time()
tr = api.Trainer(x, l, off, algo=a.algo, nleaves=a.leaves, treedepth=a.depth, nthresholds=0, cutoff=10,
                 hist_mode=a.mode)
print("init %.2fs" % (time.time() - t0), flush=True)
for i in range(a.settle):
    tr.boost_iteration(want_tree=False, want_metric=False)
    if i % 10 == 9:
        print("settle iter %d: stats %s rounds/beta %s" % (i, tr.last_tree_stats(), tr.last_tree_rounds()), flush=True)
for i in range(3):
    t0 = time.time()
    tree, m = tr.boost_iteration()
    print("warm iter %d: %.2f ms metric %.4f nodes %d" % (i, (time.time() - t0) * 1e3, m, len(tree["feature"])), flush=True)
tr.set_profiling(True)
tr.phase_times(reset=True)
t0 = time.time()
for i in range(a.trees):
    tr.boost_iteration(want_tree=False)
dt = (time.time() - t0) / a.trees
ms, ln = tr.phase_times()
print("profiled (sync per phase): %.2f ms/tree" % (dt * 1e3))
for k in ms:
    print("  %-10s %8.3f ms/tree  %6.1f launches/tree" % (k, ms[k] / a.trees, ln[k] / a.trees))
print("stats rho/sigma/splits", tr.last_tree_stats(), "rounds/beta", tr.last_tree_rounds())
tr.set_profiling(False)
t0 = time.time()
for i in range(a.trees):
    tr.boost_iteration(want_tree=False)
dt = (time.time() - t0) / a.trees
print("unprofiled: %.2f ms/tree = %.1f trees/s" % (dt * 1e3, 1 / dt))
