"""The host's Dataset / QueryResults / VerticalDataset / Dcg / Ndcg classes against the assertions of the
reference's own dataset unit test (catch-unit-tests/data/test-hdata.cc:33-105).  That test reads the MSN1
5k sample, which is not distributed with the sources; here the same checks run on a synthetic file of the
same shape (5000 documents x 136 features x 43 queries, first queries of 86 and 106 documents — SURVEY.md
section 8d, config 1).  CPU only."""
import os
import subprocess

import numpy as np
import pytest

from test_host_cli import write_svml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "host", "bin", "hdata_check")

pytestmark = pytest.mark.skipif(not os.path.exists(CHECK), reason="host/bin/hdata_check not built")


def test_hdata_assertions_of_the_reference(tmp_path):
    rng = np.random.default_rng(12)
    n, f = 5000, 136
    lens = [86, 106] + [0] * 41
    rest = n - 192
    cuts = np.sort(rng.choice(np.arange(1, rest), size=40, replace=False))
    lens[2:] = np.diff(np.concatenate([[0], cuts, [rest]])).tolist()
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    assert off[-1] == n and len(off) == 44
    x = rng.integers(0, 9, size=(n, f)).astype(np.float32)
    x[:, f - 1] = rng.integers(1, 9, size=n)                 # the last feature is always present: 136 columns
    labels = rng.integers(0, 5, size=n).astype(np.float32)
    labels[:3] = [2, 2, 0]                                   # the values test-hdata.cc reads from MSN1
    path = str(tmp_path / "hdata.txt")
    write_svml(path, x, labels, off)
    out = subprocess.run([CHECK, path], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = dict(line.split() for line in out.stdout.strip().split("\n"))
    g = lambda k: float(got[k])   # noqa: E731
    # test-hdata.cc:43-45
    assert (int(got["num_features"]), int(got["num_instances"]), int(got["num_queries"])) == (136, 5000, 43)
    # :48-69: query results of queries 0 and 1
    assert int(got["q0.num_results"]) == 86 and int(got["q1.num_results"]) == 106
    for q, d0 in ((0, 0), (1, 86)):
        for i in range(3):
            assert g("q%d.label%d" % (q, i)) == labels[d0 + i]
            assert g("q%d.feature_%d_%d" % (q, i, i)) == x[d0 + i, i]
    # :72-94: DCG@3 / NDCG@3 closed forms for scores {3,2,1} and {1,2,3} on query 0
    l0 = labels[:86]
    gain = lambda v: 2.0 ** v - 1.0   # noqa: E731
    dcg_a = gain(l0[0]) + gain(l0[1]) / np.log2(3) + gain(l0[2]) / 2
    dcg_b = gain(l0[2]) + gain(l0[1]) / np.log2(3) + gain(l0[0]) / 2
    ideal = np.sort(l0)[::-1][:3]
    idcg = gain(ideal[0]) + gain(ideal[1]) / np.log2(3) + gain(ideal[2]) / 2
    assert g("dcg3.a") == pytest.approx(dcg_a, rel=1e-12) and g("dcg3.b") == pytest.approx(dcg_b, rel=1e-12)
    assert g("ndcg3.a") == pytest.approx(dcg_a / idcg, rel=1e-12) and g("ndcg3.b") == pytest.approx(dcg_b / idcg, rel=1e-12)
    # :96-104: the vertical dataset has the same shape and a column-major layout
    assert (int(got["v.num_features"]), int(got["v.num_instances"]), int(got["v.num_queries"])) == (136, 5000, 43)
    assert int(got["v.q0.num_results"]) == 86
    for i in range(3):
        assert g("v.q0.feature_%d_%d" % (i, i)) == x[i, i]
    assert g("v.q1.feature_2_2") == x[86 + 2, 2]
    assert g("v.at_5_1") == x[5, 1]
