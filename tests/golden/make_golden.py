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


def main_sampled():
    """sampled.npz: the document-sampling trainers (LambdaMartSelective) — masked pseudo-responses, draws of
    sampling_query_level, one whole learn() run.  Inputs that are not a pure function of a seed are stored too."""
    vec = {}
    x, l, off = synth.make_dataset(1500, 6, 18, seed=111)
    rng = np.random.default_rng(12)
    scores = rng.integers(0, 9, size=len(l)).astype(np.float64) * 0.25 + np.where(rng.random(len(l)) < 0.5, rng.normal(size=len(l)), 0.0)
    mask = (rng.random(len(l)) < 0.55).astype(np.uint8)
    with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=1, nleaves=4, cutoff=10) as s:
        s.init()
        s.set_scores(scores)
        s.compute_pseudoresponses_masked(mask)
        lam, w = s.get_gradients()
    vec["masked_scores"], vec["masked_presence"], vec["masked_lambda"], vec["masked_weight"] = scores, mask, lam, w
    lens = rng.integers(1, 50, size=60)
    doff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    dl = rng.choice([0, 0, 0, 1, 2, 3], size=int(doff[-1])).astype(np.float32)
    ds = np.round(rng.normal(size=int(doff[-1])), 1)
    vec["draw_labels"], vec["draw_scores"], vec["draw_offsets"] = dl, ds, doff
    for i, (rank, rnd, adaptive, negative, adapt) in enumerate([(0.3, 0.2, "NO", "RATIO", 1.0), (0.5, 0.25, "MIX", "POS", 0.37),
                                                               (1.5, 0.5, "FIXED", "MUL", 0.6), (0.2, 0.6, "RATIO", "RATIO", 0.8)]):
        nsel, ids = pyref.selective_sample(dl, ds, doff, rank, rnd, adaptive, negative, adapt)
        vec["draw%d_params" % i] = np.array(repr((rank, rnd, adaptive, negative, adapt)))
        vec["draw%d_n" % i], vec["draw%d_ids" % i] = np.array(nsel), ids
    c = dict(n=2500, f=10, q=25, seed=112, trees=8, leaves=8, minls=5, cutoff=10,
             selective=dict(sampling_iterations=3, rank_factor=0.3, random_factor=0.2))
    x, l, off = synth.make_dataset(c["n"], c["f"], c["q"], seed=c["seed"])
    vec["learn_case"] = np.array(repr(c))
    with pyref.RefSession("LAMBDAMART-SELECTIVE", x, l, off, ntrees=c["trees"], nleaves=c["leaves"], minleafsupport=c["minls"],
                          cutoff=c["cutoff"], selective=c["selective"]) as s:
        s.learn()
        vec["learn_metric"] = s.metric_history()
        vec["learn_log"] = np.array(s.log())
        for t in range(c["trees"]):
            tr = s.tree(t)
            for k in ("feature", "threshold_idx", "threshold", "left", "right", "value", "count"):
                vec["learn_tree%d_%s" % (t, k)] = tr[k]
    np.savez_compressed(os.path.join(HERE, "sampled.npz"), **vec)
    print("wrote sampled")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sampled":
        main_sampled()
    else:
        main()
        main_sampled()
