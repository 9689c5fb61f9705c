"""Line search / CLEAVER on the GPU against the UNMODIFIED reference (oracle/_ref,
LineSearch::learn, src/learning/linear/line_search.cc:153-416) and against the oracle's NDCG.

The learned weights must equal the reference's bit for bit: every candidate's score vector is formed in the
reference's arithmetic (separate multiply/add where its Release build has them, fused where it fuses), rankings use the
libstdc++ introsort replica, the metric is the sequential mean of per-query NDCG@k, and acceptance takes the first
maximum — so every comparison `metric > best` sees the same two doubles."""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref
from quickrank_b200 import api, synth
from quickrank_b200.linesearch import Cleaver, LineSearch
import qr_testlib as common

pytestmark = pytest.mark.gpu


def _features(n=3000, f=8, q=30, seed=3):
    return synth.make_dataset(n, f, q, seed=seed)


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("cfg", [
    dict(num_points=8, max_iterations=3),
    dict(num_points=21, max_iterations=4, window_size=2.0, reduction_factor=0.8),
    dict(num_points=10, max_iterations=6, adaptive=True),
    dict(num_points=12, max_iterations=3, last_only=3),
])
def test_line_search_learns_the_reference_weights(cfg):
    x, l, off = _features()
    want = pyref.linesearch(x, l, off, cutoff=10, **cfg)
    ls = LineSearch(**cfg)
    with api.LineSearchDevice(x, l, off, cutoff=10) as dev:
        got = ls.learn(dev)
        assert dev.launch_count() > 0
    assert np.array_equal(got, want), (got, want)
    # the learned weights are worth what the oracle says they are
    scores = (x.astype(np.float64) * got).sum(axis=1)
    assert abs(ls.metric_on_training - po.ndcg_dataset(l, scores, off, 10)) <= 1e-9


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
def test_line_search_on_the_partial_scores_of_an_ensemble():
    """The CLEAVER input: one column per tree (Driver::extract_partial_scores with unit weights, driver.cc:411-446),
    starting from the ensemble's own weights."""
    x, l, off = common.dataset(n=4000, f=12, q=40, seed=9)
    with api.Trainer(x, l, off, nleaves=8, shrinkage=0.1) as tr:
        trees = [tr.boost_iteration()[0] for _ in range(10)]
    with api.Scorer(trees, np.ones(len(trees)), x.shape[1]) as sc:
        part = sc.partial_scores(x)
    w0 = np.full(len(trees), 0.1)
    want = pyref.linesearch(part, l, off, cutoff=10, num_points=10, max_iterations=3, window_size=1.0, init_weights=w0)
    ls = LineSearch(num_points=10, max_iterations=3, window_size=1.0)
    ls.weights = w0.copy()
    with api.LineSearchDevice(part, l, off, cutoff=10) as dev:
        before = dev.evaluate(w0)
        got = ls.learn(dev)
    assert np.array_equal(got, want)
    assert ls.metric_on_training >= before


def test_device_metrics_match_the_oracle():
    """qr_ls_evaluate / feature_points / line_points: NDCG@10 of the candidate weight vectors (oracle: numpy scores in
    float64 + the restated NDCG; the sums differ from the device's in rounding only: 1e-12)."""
    x, l, off = _features(n=2500, f=6, q=25, seed=5)
    rng = np.random.default_rng(1)
    w = rng.random(6)
    xd = x.astype(np.float64)
    with api.LineSearchDevice(x, l, off, cutoff=10) as dev:
        assert abs(dev.evaluate(w) - po.ndcg_dataset(l, xd @ w, off, 10)) <= 1e-12
        pts = np.array([0.0, 0.3, 1.7])
        got = dev.feature_points(w, 2, pts)
        for p, g in zip(pts, got):
            w2 = w.copy(); w2[2] = p
            assert abs(g - po.ndcg_dataset(l, xd @ w2, off, 10)) <= 1e-12
        step = rng.normal(0, 0.05, 6)
        got = dev.line_points(w, step, 5)
        for p, g in enumerate(got):
            assert abs(g - po.ndcg_dataset(l, xd @ (w + step * p), off, 10)) <= 1e-12


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("with_ls", [True, False])
@pytest.mark.parametrize("method", ["LAST", "SKIP", "LOW_WEIGHTS", "QUALITY_LOSS", "QUALITY_LOSS_ADV", "SCORE_LOSS"])
def test_cleaver_equals_the_reference(method, with_ls):
    """Cleaver::optimize (cleaver.cc:166-412) on the partial scores of a trained ensemble: the pruned set and the
    re-learned weights equal the unmodified reference's bit for bit, with and without the line search."""
    x, l, off = common.dataset(n=4000, f=12, q=40, seed=11)
    with api.Trainer(x, l, off, nleaves=8, shrinkage=0.1) as tr:
        trees = [tr.boost_iteration()[0] for _ in range(20)]
    with api.Scorer(trees, np.ones(len(trees)), x.shape[1]) as sc:
        part = sc.partial_scores(x)
    w0 = np.full(len(trees), 0.1)
    kw = dict(num_points=8, max_iterations=2, window_size=1.0, reduction_factor=0.95)
    want = pyref.cleaver(method, part, l, off, w0, 0.3, cutoff=10, **(kw if with_ls else dict(num_points=0)))
    cl = Cleaver(0.3, method, LineSearch(**kw) if with_ls else None)
    w, pruned = cl.optimize(part, l, off, w0, cutoff=10)
    assert len(pruned) == 6 and all(w[f] == 0 for f in pruned)
    assert np.array_equal(w, want), (method, w, want)
    if method == "LAST":
        assert pruned == set(range(14, 20))
    # the optimised ensemble scores what the oracle says
    assert abs(cl.metric_after - po.ndcg_dataset(l, part.astype(np.float64) @ w, off, 10)) <= 1e-9


@pytest.mark.gpu
def test_cleaver_random_pruning_is_seeded():
    """RandomPruning (random_pruning.cc:47-55) draws trees with rand() % last + start until enough distinct ones are
    hit; the reference seeds from the wall clock, here `seed` makes the draw reproducible."""
    x, l, off = common.dataset(n=2000, f=8, q=20, seed=3)
    rng = np.random.default_rng(1)
    part = rng.normal(size=(len(l), 15)).astype(np.float32)
    w0 = np.full(15, 0.1)
    a = Cleaver(5, "RANDOM", None, seed=7).optimize(part, l, off, w0)
    b = Cleaver(5, "RANDOM", None, seed=7).optimize(part, l, off, w0)
    c = Cleaver(5, "RANDOM", None, seed=8).optimize(part, l, off, w0)
    assert len(a[1]) == 5 and a[1] == b[1] and np.array_equal(a[0], b[0])
    assert all(a[0][f] == 0 for f in a[1]) and np.count_nonzero(a[0]) == 10
    assert a[1] != c[1]
    # last_only: only the last trees are candidates
    d = Cleaver(3, "RANDOM", None, last_only=6, seed=1).optimize(part, l, off, w0)
    assert len(d[1]) == 3 and min(d[1]) >= 9


@pytest.mark.gpu
@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("cfg", [
    dict(num_points=8, max_iterations=6),
    dict(num_points=10, max_iterations=8, max_failed_vali=2, window_size=2.0),
    dict(num_points=12, max_iterations=5, adaptive=True, max_failed_vali=3),
])
def test_line_search_with_a_validation_set(cfg):
    """line_search.cc:360-383: the weights returned are those of the best iteration ON THE VALIDATION SET, and
    `max_failed_vali` iterations in a row without a new validation best end the search — same weights as the
    unmodified reference, bit for bit."""
    x, l, off = _features(n=3000, f=8, q=30, seed=3)
    xv, lv, offv = _features(n=2000, f=8, q=20, seed=4)
    want = pyref.linesearch(x, l, off, cutoff=10, valid=(xv, lv, offv), **cfg)
    ls = LineSearch(**cfg)
    with api.LineSearchDevice(x, l, off, cutoff=10) as dev, api.LineSearchDevice(xv, lv, offv, cutoff=10) as vdev:
        got = ls.learn(dev, vdev)
    assert np.array_equal(got, want), (got, want)
    assert all(h[2] is not None for h in ls.history)
