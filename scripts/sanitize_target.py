"""Target of the compute-sanitizer runs (scripts/r02_sanitize.sh): a few boosting iterations in both histogram
modes (leaf-wise and oblivious) and one scoring pass on a small dataset, so that every kernel of the hot path
executes at least once under the tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from quickrank_b200 import api, synth  # noqa: E402

x, labels, qoff = synth.make_dataset(6000, 24, 60, seed=5)
trees = []
for algo, kw in (("LAMBDAMART", dict(nleaves=16)), ("MART", dict(nleaves=8)), ("OBVLAMBDAMART", dict(treedepth=3))):
    for mode in (api.HIST_FAST, api.HIST_REFERENCE):
        with api.Trainer(x, labels, qoff, algo=algo, minleafsupport=1, cutoff=10, hist_mode=mode, **kw) as tr:
            for _ in range(3):
                tree, metric = tr.boost_iteration()
            if algo == "LAMBDAMART" and mode == api.HIST_FAST:
                trees = [tree]
            tr.get_scores()
with api.Scorer(trees, [0.1], x.shape[1]) as sc:
    out = sc.score_dataset(x)
print("sanitize target ok", float(np.sum(out)))
