"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libqr_ref.so, built from
/root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/qr_oracle.c) where /root/reference is not available (the GPU box):
inputs are regenerated from the seeds stored in each file, outputs are the reference's own.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402
from quickrank_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    dict(name="lambdamart_small", algo="LAMBDAMART", n=1500, f=12, q=18, seed=101, gridded=True, nthr=0,
         leaves=8, depth=0, minls=1, trees=6, cutoff=10),
    dict(name="lambdamart_continuous", algo="LAMBDAMART", n=1200, f=9, q=14, seed=102, gridded=False, nthr=0,
         leaves=6, depth=0, minls=1, trees=4, cutoff=5),
    dict(name="lambdamart_equalwidth", algo="LAMBDAMART", n=2000, f=10, q=25, seed=103, gridded=False, nthr=16,
         leaves=10, depth=0, minls=3, trees=5, cutoff=10),
    dict(name="mart_small", algo="MART", n=1500, f=12, q=18, seed=104, gridded=True, nthr=0,
         leaves=8, depth=0, minls=1, trees=5, cutoff=10),
    dict(name="obvlambdamart_small", algo="OBVLAMBDAMART", n=1500, f=12, q=18, seed=105, gridded=True, nthr=0,
         leaves=8, depth=3, minls=1, trees=5, cutoff=10),
    dict(name="obvmart_small", algo="OBVMART", n=1500, f=12, q=18, seed=106, gridded=True, nthr=0,
         leaves=8, depth=3, minls=1, trees=4, cutoff=0),
]


def main():
    for c in CASES:
        x, l, off = synth.make_dataset(c["n"], c["f"], c["q"], seed=c["seed"], gridded=c["gridded"])
        out = {"case": np.array(repr(c))}
        with pyref.RefSession(c["algo"], x, l, off, ntrees=c["trees"], nthresholds=c["nthr"], nleaves=c["leaves"],
                              treedepth=c["depth"], minleafsupport=c["minls"], cutoff=c["cutoff"]) as s:
            s.learn(keep_gradients=True)
            out["metric"] = s.metric_history()
            for t in range(c["trees"]):
                tr = s.tree(t)
                for k in ("feature", "threshold_idx", "threshold", "left", "right", "value", "count"):
                    out["tree%d_%s" % (t, k)] = tr[k]
                out["lambda%d" % t] = s.recorded("lambdas", t)
                if c["algo"].endswith("LAMBDAMART"):
                    out["weight%d" % t] = s.recorded("weights", t)
            out["scores"] = s.recorded("scores", c["trees"] - 1)
            for f in (0, c["f"] - 1):
                out["thresholds%d" % f] = s.thresholds(f)
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **out)
        print("wrote", c["name"])
    # pure-function vectors: ties in std::sort, NDCG / DCG / jacobian
    rng = np.random.default_rng(7)
    vec = {}
    for i, n in enumerate((1, 2, 16, 17, 33, 64, 100, 257)):
        for lv in (1, 3, 1000):
            s = rng.integers(0, lv, size=n).astype(np.float64)
            vec["sort_in_%d_%d" % (n, lv)] = s
            vec["sort_out_%d_%d" % (n, lv)] = pyref.sort_indices(s)
    labels = rng.integers(0, 5, size=40).astype(np.float32)
    scores = np.round(rng.normal(size=40), 1)
    vec["m_labels"], vec["m_scores"] = labels, scores
    for k in (0, 3, 10, 100):
        vec["dcg_%d" % k] = np.array(pyref.dcg(labels, scores, k))
        vec["ndcg_%d" % k] = np.array(pyref.ndcg(labels, scores, k))
        vec["jac_%d" % k] = pyref.ndcg_jacobian(labels, scores, k)
    np.savez_compressed(os.path.join(HERE, "pure_functions.npz"), **vec)
    print("wrote pure_functions")


if __name__ == "__main__":
    main()
