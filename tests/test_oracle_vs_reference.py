"""CPU: the oracle against the unmodified reference sources compiled into oracle/_ref (only where
that library exists, i.e. in the build container; the golden fixtures cover the same ground elsewhere)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref
from quickrank_b200 import synth

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref/libqr_ref.so not built")


def test_sort_matches_std_sort():
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 137, 500, 999, 3000]:
        for levels in (1, 2, 5, 10 ** 6):
            s = rng.integers(0, levels, size=n).astype(np.float64)
            assert np.array_equal(pyref.sort_indices(s).astype(np.uint32), po.sort_desc(s)), (n, levels)
    for n in (100, 1000, 5000):   # organ pipe: drives introsort to its depth limit
        s = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.float64)
        assert np.array_equal(pyref.sort_indices(s).astype(np.uint32), po.sort_desc(s))


def test_radix_argsort_matches():
    rng = np.random.default_rng(1)
    v = rng.normal(size=5000).astype(np.float32)
    v[::7] = 0.0
    v[::11] = -0.0
    v[::13] = v[3]
    assert np.array_equal(pyref.radix_argsort(v), po.radix_argsort(v))


@pytest.mark.parametrize("cutoff", [10, 3, 0])
def test_metric_and_lambdas(cutoff):
    rng = np.random.default_rng(cutoff)
    x, l, off = synth.make_dataset(3000, 6, 40, seed=5, qlen=(1, 150))
    with pyref.RefSession("LAMBDAMART", x, l, off, cutoff=cutoff) as s:
        s.init()
        for scores in (np.zeros(len(l)), np.round(rng.normal(size=len(l)), 1), rng.normal(size=len(l))):
            s.set_scores(scores)
            s.compute_pseudoresponses()
            lam, w = s.get_gradients()
            olam, ow = po.lambdas(scores, l, off, cutoff)
            assert np.array_equal(lam, olam) and np.array_equal(w, ow)
            assert s.evaluate() == po.ndcg_dataset(l, scores, off, cutoff)


@pytest.mark.parametrize("algo,depth", [("LAMBDAMART", 0), ("MART", 0), ("OBVLAMBDAMART", 4), ("OBVMART", 3)])
@pytest.mark.parametrize("gridded,nthr", [(True, 0), (False, 0), (False, 24)])
def test_training_is_bit_identical(algo, depth, gridded, nthr):
    T = 6
    x, l, off = synth.make_dataset(3000, 17, 30, seed=9, gridded=gridded)
    with pyref.RefSession(algo, x, l, off, ntrees=T, nthresholds=nthr, nleaves=12, treedepth=depth,
                          minleafsupport=2, cutoff=10) as s:
        s.learn(keep_gradients=True)
        trees, metric, scores = po.train(algo, x, l, off, T, nthresholds=nthr, nleaves=12, depth=depth, minls=2, cutoff=10)
        assert np.array_equal(metric, s.metric_history())
        assert np.array_equal(scores, s.recorded("scores", T - 1))
        for t in range(T):
            rt = s.tree(t)
            for k in ("feature", "threshold_idx", "threshold", "left", "right", "value"):
                assert np.array_equal(rt[k], trees[t][k]), (t, k)


def test_scoring_through_the_reference_xml_loader(tmp_path):
    """Model written by the reference, scored by the reference's score_dataset, equals the oracle's
    ensemble scoring of the same trees."""
    x, l, off = synth.make_dataset(2000, 10, 20, seed=2)
    with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=5, nleaves=8) as s:
        s.learn()
        trees = [s.tree(t) for t in range(5)]
        path = str(tmp_path / "model.xml")
        s.save_model(path)
    want = pyref.score_with_model(path, x)
    got = po.score_dataset(trees, [t["weight"] for t in trees], x)
    assert np.array_equal(got, want)
