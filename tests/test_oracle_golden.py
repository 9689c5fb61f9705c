"""CPU: the oracle (oracle/qr_oracle.c) against
  * the known-answer vectors of the reference's own unit tests (catch-unit-tests/metric/ir/test-dcg.cc,
    test-ndcg.cc: labels {3,2,1,0,0}, scores {5,4,3,2,1}),
  * golden vectors produced by the unmodified reference (tests/golden/*.npz, made by
    tests/golden/make_golden.py from oracle/_ref).
Everything is compared bit for bit: the oracle replicates the reference's arithmetic order."""
import ast
import glob
import math
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from quickrank_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_unit_test_vectors_dcg():
    # catch-unit-tests/metric/ir/test-dcg.cc:35-98
    labels, scores = [3, 2, 1, 0, 0], [5, 4, 3, 2, 1]
    full = (2 ** 3 - 1) / math.log2(2) + (2 ** 2 - 1) / math.log2(3) + (2 ** 1 - 1) / math.log2(4)
    assert po.dcg_query(labels, scores, 5) == pytest.approx(full, rel=1e-15)
    assert po.dcg_query(labels, scores, 2) == pytest.approx(7 + 3 / math.log2(3), rel=1e-15)
    assert po.dcg_query(labels, scores, 0) == pytest.approx(full, rel=1e-15)        # 0 -> NO_CUTOFF (metric.h:65-67)
    assert po.dcg_query(labels, scores, 100) == pytest.approx(full, rel=1e-15)
    # a worse ranking
    assert po.dcg_query(labels, [1, 2, 3, 4, 5], 5) < full


def test_reference_unit_test_vectors_ndcg_and_jacobian():
    # catch-unit-tests/metric/ir/test-ndcg.cc:39-105
    labels = np.array([3, 2, 1, 0, 0], np.float32)
    scores = np.array([5, 4, 3, 2, 1], np.float64)
    for k in (5, 2, 0, 10):
        assert po.ndcg_query(labels, scores, k) == 1.0
    # jacobian(0, 2) equals the NDCG change of swapping ranks 0 and 2 (test-ndcg.cc:70-105)
    for k in (5, 3):
        idcg = po.idcg(labels, k)
        swapped = scores.copy()
        swapped[[0, 2]] = swapped[[2, 0]]
        delta = po.ndcg_query(labels, swapped, k) - po.ndcg_query(labels, scores, k)
        assert po.delta_ndcg(labels, k, idcg, 0, 2) == pytest.approx(delta, rel=1e-12)
    # all labels equal -> ideal DCG 0 -> NDCG 0 (ndcg.cc:54-57)
    assert po.ndcg_query([0, 0, 0], [3, 2, 1], 10) == 0.0
    assert po.ndcg_query([], [], 10) == 0.0


def test_sort_dcg_ndcg_jacobian_against_reference_vectors():
    v = np.load(os.path.join(GOLDEN, "pure_functions.npz"))
    n_sort = 0
    for key in v.files:
        if key.startswith("sort_in_"):
            out = v["sort_out_" + key[len("sort_in_"):]]
            assert np.array_equal(po.sort_desc(v[key]), out.astype(np.uint32)), key
            n_sort += 1
    assert n_sort >= 20
    labels, scores = v["m_labels"], v["m_scores"]
    order = po.sort_desc(scores)
    sl = labels[order]
    for k in (0, 3, 10, 100):
        assert po.dcg_query(labels, scores, k) == float(v["dcg_%d" % k])
        assert po.ndcg_query(labels, scores, k) == float(v["ndcg_%d" % k])
        jac = v["jac_%d" % k]
        idcg = po.idcg(sl, k)
        n = len(labels)
        for i in range(n):
            for j in range(i + 1, n):
                assert po.delta_ndcg(sl, k, idcg, i, j) == jac[i, j], (k, i, j)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*_*.npz"))))
def test_training_against_reference_vectors(path):
    if path.endswith("pure_functions.npz"):
        pytest.skip("not a training fixture")
    g = np.load(path)
    c = ast.literal_eval(str(g["case"]))
    x, l, off = synth.make_dataset(c["n"], c["f"], c["q"], seed=c["seed"], gridded=c["gridded"])
    trees, metric, scores = po.train(c["algo"], x, l, off, c["trees"], nthresholds=c["nthr"], nleaves=c["leaves"],
                                     depth=c["depth"], minls=c["minls"], cutoff=c["cutoff"])
    assert np.array_equal(metric, g["metric"])
    assert np.array_equal(scores, g["scores"])
    for t in range(c["trees"]):
        for k in ("feature", "threshold_idx", "threshold", "left", "right", "value"):
            assert np.array_equal(trees[t][k], g["tree%d_%s" % (t, k)]), (t, k)
        lv = trees[t]["feature"] < 0
        assert np.array_equal(trees[t]["count"][lv], g["tree%d_count" % t][lv])
    ob = po.Binning(np.ascontiguousarray(x.T), c["nthr"])
    for f in (0, c["f"] - 1):
        assert np.array_equal(ob.thresholds(f), g["thresholds%d" % f])
    # stage-wise: lambdas of every iteration from the reference's recorded scores trajectory
    if c["algo"].endswith("LAMBDAMART"):
        lam, w = po.lambdas(np.zeros(len(l)), l, off, c["cutoff"])
        assert np.array_equal(lam, g["lambda0"]) and np.array_equal(w, g["weight0"])


def test_edge_cases():
    # single-document and empty queries, one leaf, unsplittable data
    x, l, off = synth.make_dataset(300, 5, 6, seed=3)
    off2 = np.array([0, 1, 1, 50, 300], np.uint64)      # one-doc query, empty query
    lam, w = po.lambdas(np.zeros(300), l, off2, 10)
    assert lam[0] == 0.0 and w[0] == 0.0
    assert np.isfinite(po.ndcg_dataset(l, np.zeros(300), off2, 10))
    ob = po.Binning(np.ascontiguousarray(x.T), 0)
    t = ob.fit_tree(np.zeros(300), np.zeros(300), nleaves=8)       # zero gradients: deviance 0 -> root stays a leaf
    assert len(t["feature"]) == 1 and t["feature"][0] == -1 and t["value"][0] == 0.0
    t = ob.fit_tree(lam, w, nleaves=1)
    assert len(t["feature"]) == 3 or len(t["feature"]) == 1        # rt.cc:57-61: the root split precedes the budget test
