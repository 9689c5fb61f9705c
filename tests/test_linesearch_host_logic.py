"""CPU: the host logic of quickrank_b200/linesearch.py (LineSearch.learn, Cleaver.optimize) against the unmodified
reference, with the passes over documents done by a numpy stand-in for the device (tests/cpu_linesearch_device.py)
instead of the GPU.  The GPU versions of the same comparisons are in tests/test_linesearch.py."""
import numpy as np
import pytest

from oracle import pyref
from quickrank_b200 import linesearch, synth
from quickrank_b200.linesearch import Cleaver, LineSearch
from cpu_linesearch_device import CpuLineSearchDevice

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")


def _features(n=900, f=6, q=12, seed=3):
    return synth.make_dataset(n, f, q, seed=seed)


@pytest.mark.parametrize("cfg", [
    dict(num_points=6, max_iterations=2),
    dict(num_points=9, max_iterations=3, window_size=2.0, reduction_factor=0.8),
    dict(num_points=6, max_iterations=4, adaptive=True),
    dict(num_points=6, max_iterations=2, last_only=2),
])
def test_line_search_host_logic(cfg):
    x, l, off = _features()
    want = pyref.linesearch(x, l, off, cutoff=10, **cfg)
    got = LineSearch(**cfg).learn(CpuLineSearchDevice(x, l, off, cutoff=10))
    assert np.array_equal(got, want), (got, want)


def test_line_search_host_logic_with_validation():
    x, l, off = _features()
    xv, lv, offv = _features(n=600, f=6, q=8, seed=4)
    cfg = dict(num_points=6, max_iterations=5, max_failed_vali=2, window_size=2.0)
    want = pyref.linesearch(x, l, off, cutoff=10, valid=(xv, lv, offv), **cfg)
    got = LineSearch(**cfg).learn(CpuLineSearchDevice(x, l, off, cutoff=10), CpuLineSearchDevice(xv, lv, offv, cutoff=10))
    assert np.array_equal(got, want), (got, want)


@pytest.mark.parametrize("with_ls", [True, False])
@pytest.mark.parametrize("method", ["LAST", "SKIP", "LOW_WEIGHTS", "QUALITY_LOSS", "QUALITY_LOSS_ADV", "SCORE_LOSS"])
def test_cleaver_host_logic(monkeypatch, method, with_ls):
    """Cleaver::optimize (cleaver.cc:166-412): pruned set and re-learned weights equal the reference's."""
    monkeypatch.setattr(linesearch.api, "LineSearchDevice", CpuLineSearchDevice)
    x, l, off = _features(n=900, f=10, q=12, seed=5)
    part = (x - 0.4).astype(np.float32)          # a partial-score-like matrix: signed columns
    w0 = np.full(part.shape[1], 0.1)
    kw = dict(num_points=6, max_iterations=2, window_size=1.0, reduction_factor=0.95)
    want = pyref.cleaver(method, part, l, off, w0, 0.3, cutoff=10, **(kw if with_ls else dict(num_points=0)))
    w, pruned = Cleaver(0.3, method, LineSearch(**kw) if with_ls else None).optimize(part, l, off, w0, cutoff=10)
    assert len(pruned) == 3 and all(w[f] == 0 for f in pruned)
    assert np.array_equal(w, want), (method, w, want)
