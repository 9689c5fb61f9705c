"""Python model XML I/O (quickrank_b200/modelxml.py) against the reference's own loader."""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref
from quickrank_b200 import modelxml, synth


def test_round_trip(tmp_path):
    trees, weights = synth.random_ensemble(7, 12, 9, seed=3)
    trees.append(dict(feature=np.array([-1], np.int32), threshold=np.zeros(1, np.float32), left=np.array([-1], np.int32),
                      right=np.array([-1], np.int32), value=np.array([0.25])))          # a single-leaf tree
    weights = np.append(weights, 0.5)
    path = str(tmp_path / "m.xml")
    modelxml.write_model(path, trees, weights)
    info, got, w = modelxml.read_model(path)
    assert info["type"] == "LAMBDAMART" and int(info["trees"]) == len(trees)
    assert np.array_equal(w, weights)
    for a, b in zip(got, trees):
        for k in ("feature", "threshold", "left", "right", "value"):
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
def test_written_model_is_loaded_by_the_reference(tmp_path):
    trees, weights = synth.random_ensemble(20, 16, 11, seed=5)
    rng = np.random.default_rng(1)
    x = (rng.integers(0, 256, (500, 11)) / 255.0).astype(np.float32)
    path = str(tmp_path / "m.xml")
    modelxml.write_model(path, trees, weights)
    assert np.array_equal(pyref.score_with_model(path, x), po.score_dataset(trees, weights, x))


@pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")
def test_reads_a_model_written_by_the_reference(tmp_path):
    x, l, off = synth.make_dataset(1500, 8, 15, seed=4)
    with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=4, nleaves=6) as s:
        s.learn()
        want = [s.tree(t) for t in range(4)]
        path = str(tmp_path / "ref.xml")
        s.save_model(path)
    info, trees, w = modelxml.read_model(path)
    assert info["type"] == "LAMBDAMART"
    assert np.array_equal(po.score_dataset(trees, w, x), pyref.score_with_model(path, x))
    for a, b in zip(trees, want):
        assert np.array_equal(a["feature"], b["feature"]) and np.array_equal(a["threshold"], b["threshold"])
        leaf = b["feature"] < 0                      # the file holds outputs of leaves only
        assert np.array_equal(a["value"][leaf], b["value"][leaf])
