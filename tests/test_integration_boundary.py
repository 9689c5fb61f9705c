"""The drop-in boundary, compiled and run: oracle/integration/gpulambdamart.h is the subclass INTEGRATION.md
describes, built against the UNMODIFIED reference headers and linked with the reference's own objects
(oracle/Makefile target `integration`).  The reference's stock Mart::learn (mart.cc:208-416) then drives the GPU
through the Mart hooks (mart.h:118-147) and its own XML writer saves the model.  Compared with this repository's
quicklearn on the same file: the same stdout table and the same trees."""
import os
import re
import subprocess

import numpy as np
import pytest

import qr_testlib as common
from oracle import pyref
from test_host_cli import QL, write_svml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "oracle", "_ref", "gpu_lambdamart_demo")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(DEMO), reason="oracle/_ref/gpu_lambdamart_demo is not built")]


def _trees(xml):
    """(feature, threshold, output) lists of every tree of a QuickRank XML model, in document order."""
    out = []
    for tree in re.findall(r"<tree .*?</tree>", xml, flags=re.S):
        out.append((re.findall(r"<feature>\s*(\d+)\s*</feature>", tree),
                    [float(v) for v in re.findall(r"<threshold>\s*([^<\s]+)\s*</threshold>", tree)],
                    [float(v) for v in re.findall(r"<output>\s*([^<\s]+)\s*</output>", tree)]))
    return out


@pytest.mark.parametrize("mode", ["fast", "reference"])
def test_stock_mart_learn_through_the_gpu_subclass(tmp_path, mode):
    x, l, off = common.dataset(n=4000, f=16, q=40, seed=21)
    train = str(tmp_path / "train.txt")
    write_svml(train, x, l, off)
    m_ref, m_host = str(tmp_path / "via_reference_learn.xml"), str(tmp_path / "via_quicklearn.xml")
    a = subprocess.run([DEMO, train, m_ref, "12", "10"] + (["reference"] if mode == "reference" else []),
                       capture_output=True, text=True)
    assert a.returncode == 0, a.stderr + a.stdout
    b = subprocess.run([QL, "--algo", "LAMBDAMART", "--train", train, "--num-trees", "12", "--num-leaves", "10",
                        "--model-out", m_host, "--hist-mode", mode], capture_output=True, text=True)
    assert b.returncode == 0, b.stderr + b.stdout
    # the reference's learn() printed its own table from the device-side NDCG: same rows as quicklearn's
    rows = lambda s: re.findall(r"^\s+(\d+)\s+([0-9.]+)( \*)?$", s, flags=re.M)
    assert len(rows(a.stdout)) == 12 and rows(a.stdout) == rows(b.stdout)
    ta, tb = _trees(open(m_ref).read()), _trees(open(m_host).read())
    assert len(ta) == 12 and len(tb) == 12
    for (fa, tha, oa), (fb, thb, ob) in zip(ta, tb):
        assert fa == fb
        assert np.array_equal(np.float32(tha), np.float32(thb))
        assert np.array_equal(oa, ob)
    # and the model the reference wrote scores through the reference's own loader
    if pyref.available():
        s = pyref.score_with_model(m_ref, x)
        assert np.isfinite(s).all() and np.ptp(s) > 0
