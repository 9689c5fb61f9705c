"""Second target of the compute-sanitizer runs: the kernels added for the document-sampling trainers (sample contexts:
gathered bins, gathered scores / ranking keys, in-place redraw) and for CLEAVER's pruning passes (drop_points,
drop_column, score_loss), each executed at least once on a small dataset."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from quickrank_b200 import api, synth  # noqa: E402
from quickrank_b200.linesearch import Cleaver  # noqa: E402

x, labels, qoff = synth.make_dataset(5000, 20, 50, seed=7)
rng = np.random.default_rng(3)
trees = []
with api.Trainer(x, labels, qoff, nleaves=8, minleafsupport=5) as full:
    with full.sample_context(x, np.arange(len(labels))) as sm:
        for it in range(6):
            if it in (2, 4):
                sm.redraw(full, np.nonzero(rng.random(len(labels)) < (0.4 if it == 2 else 0.9))[0])
            sm.pull_scores(full)
            sm.compute_pseudoresponses()
            tree = sm.fit_regressor_on_gradient()
            full.apply_tree(tree, full.shrinkage)
            full.evaluate_dataset()
            trees.append(tree)
    with full.sample_context(x, np.arange(0, len(labels), 3), gather=False) as sm2:
        sm2.pull_scores(full)
        sm2.compute_pseudoresponses()
        sm2.fit_regressor_on_gradient()
with api.Scorer(trees, np.ones(len(trees)), x.shape[1]) as sc:
    part = sc.partial_scores(x)
w0 = np.full(len(trees), 0.1)
out = []
for method in ("QUALITY_LOSS", "QUALITY_LOSS_ADV", "SCORE_LOSS"):
    w, pruned = Cleaver(2, method, None).optimize(part, labels, qoff, w0)
    out.append(sorted(pruned))
print("sanitize target 2 ok", out)
