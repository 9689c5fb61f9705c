"""quickrank_b200.linesearch.std_sort restates libstdc++'s std::sort; the order it leaves tied elements in must be the
one the reference's own `std::sort` produces (checked through QueryResults::indexing_of_sorted_labels of the
unmodified reference, queryresults.cc:47-53: an index sort by descending score)."""
import numpy as np
import pytest

from oracle import pyref
from quickrank_b200.linesearch import std_sort

pytestmark = pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")


@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 33, 100, 1000, 5000])
@pytest.mark.parametrize("levels", [2, 7, 0])
def test_std_sort_replica_orders_ties_like_libstdcxx(n, levels):
    rng = np.random.default_rng(n * 31 + levels)
    scores = rng.integers(0, levels, n).astype(np.float64) if levels else rng.random(n)
    want = pyref.sort_indices(scores)
    got = std_sort(list(range(n)), lambda a, b: scores[a] > scores[b])
    assert list(want) == got


def test_std_sort_replica_on_adversarial_patterns():
    for scores in (np.arange(3000.0), np.arange(3000.0)[::-1].copy(), np.zeros(3000), np.tile([1.0, 0.0], 1500),
                   np.concatenate([np.arange(1500.0), np.arange(1500.0)])):
        want = pyref.sort_indices(scores)
        got = std_sort(list(range(len(scores))), lambda a, b: scores[a] > scores[b])
        assert list(want) == got
