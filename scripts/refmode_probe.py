"""Development probe: QR_HIST_REFERENCE ms/tree with the per-phase device times, config-2 shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quickrank_b200 import api, synth
trees = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
x, l, off = synth.make_dataset(n, 136, n // 100, seed=20260102)
t0 = time.time()
tr = api.Trainer(x, l, off, algo="LAMBDAMART", nleaves=64, nthresholds=0, cutoff=10, hist_mode=api.HIST_REFERENCE)
print("create %.2f s (QR_EXACT_WALK_MIN=%s)" % (time.time() - t0, os.environ.get("QR_EXACT_WALK_MIN")), flush=True)
names = ["pseudo", "hist", "scan", "partition", "leaf", "rank"]
for i in range(trees):
    prof = i in (3, trees - 1)
    if prof:
        tr.set_profiling(True); tr.phase_times(reset=True)
    t0 = time.time()
    tr.boost_iteration(want_tree=False, want_metric=True)
    dt = (time.time() - t0) * 1e3
    msg = "tree %3d: %8.2f ms, rounds %d" % (i, dt, tr.last_tree_rounds()[0])
    if prof:
        ms, la = tr.phase_times(reset=True)
        msg += "  phases(ms, with syncs): " + " ".join("%s=%.2f/%d" % (nm, ms[nm], la[nm]) for nm in ms)
        tr.set_profiling(False)
    print(msg, flush=True)
