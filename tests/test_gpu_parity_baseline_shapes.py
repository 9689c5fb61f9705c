"""GPU parity at the shapes BASELINE.json names, against the UNMODIFIED reference (oracle/_ref), stage by stage:

* config 1: 5 000 docs x 136 continuous features x 43 queries, LAMBDAMART 50 trees / 10 leaves (the shape of the
  reference's own plumbing test, catch-unit-tests/learning/forests/test-lambdamart.cc:135-137) — every tree;
* config 2: 1 000 000 docs x 136 features x 10 000 queries, LAMBDAMART 64 leaves (the benchmarked workload,
  bench.py WORKLOAD, same seed) — the first 3 trees.

Each boosting iteration starts from the reference's own scores: the GPU's pseudo-responses must equal the
reference's bit for bit, then the tree is fitted from the REFERENCE's pseudo-responses:
QR_HIST_REFERENCE must reproduce the reference's tree bit for bit (split feature / threshold index, counts,
leaf outputs), QR_HIST_FAST up to audited exact-arithmetic ties (qr_testlib.audit_tree) with leaf outputs within
1e-5 relative; NDCG@10 within 1e-5 relative.  The restatement (oracle/qr_oracle.c, bit-identical to the reference:
tests/test_oracle_vs_reference.py) supplies the bins and the candidate scores the audit needs."""
import numpy as np
import pytest

from oracle import pyoracle as po
from oracle import pyref
from quickrank_b200 import api, synth
import qr_testlib as common

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not pyref.available(), reason="oracle/_ref is not built")]

REL = 1e-5  # north_star tolerance for leaf outputs / NDCG


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0


def _stagewise(x, l, off, leaves, ntrees, mode, max_near):
    col = np.ascontiguousarray(x.T)
    ob = po.Binning(col, 0) if mode == api.HIST_FAST else None
    bins = ob.bins() if ob is not None else None
    near_total = equiv_total = 0
    with pyref.RefSession("LAMBDAMART", x, l, off, ntrees=ntrees + 1, nleaves=leaves, minleafsupport=1,
                          cutoff=10) as ref, \
            api.Trainer(x, l, off, algo="LAMBDAMART", nleaves=leaves, minleafsupport=1, cutoff=10,
                        hist_mode=mode) as tr:
        ref.init()
        for f in (0, 5, x.shape[1] - 1):
            assert np.array_equal(tr.thresholds(f), ref.thresholds(f)), "thresholds of feature %d" % f
        for m in range(ntrees):
            scores = ref.get_scores()
            ref.compute_pseudoresponses()
            lam, w = ref.get_gradients()
            tr.set_scores(scores)
            tr.compute_pseudoresponses()
            glam, gw = tr.get_pseudoresponses()
            assert np.max(np.abs(glam - lam)) <= 1e-13 * np.max(np.abs(lam)), "lambda, tree %d" % m
            assert np.max(np.abs(gw - w)) <= 1e-13 * np.max(np.abs(w)), "weights, tree %d" % m
            # bit for bit at every iteration: the kernel's exp() is glibc's algorithm restated (qr_kernels.cuh, exp_lambda),
            # everything else on the path is IEEE arithmetic in the reference's order
            assert np.array_equal(glam, lam) and np.array_equal(gw, w), "pseudo-responses differ in the last bits, tree %d" % m
            tr.set_pseudoresponses(lam, w)
            got = tr.fit_regressor_on_gradient()
            ref.fit_tree(True)
            want = ref.tree(m)
            if mode == api.HIST_REFERENCE:
                assert common.same_structure(got, want), "tree %d: %s" % (m, common.describe_tree_diff(got, want))
                assert np.array_equal(got["count"], want["count"]), "tree %d" % m
                lv = common.leaves_mask(want)
                assert np.array_equal(got["value"][lv], want["value"][lv]), "tree %d leaf outputs" % m
                clean = np.ones(len(l), bool)
            else:
                equiv, near, clean = common.audit_tree(got, want, ob, bins, lam, 1)
                equiv_total += equiv
                near_total += near
                og, ow = common.tree_outputs(got, bins), common.tree_outputs(want, bins)
                assert rel_err(og[clean], ow[clean]) <= REL, "tree %d leaf outputs" % m
            tr.update_modelscores()
            gs, rs = tr.get_scores(), ref.get_scores()
            assert np.max(np.abs(gs[clean] - rs[clean])) <= REL * np.max(np.abs(rs)), "scores, tree %d" % m
            if clean.all():
                metric, want_metric = tr.evaluate_dataset(), ref.evaluate()
                assert abs(metric - want_metric) <= REL * want_metric, "NDCG@10, tree %d" % m
    assert near_total <= max_near, "%d rounding-decided near-ties" % near_total
    return equiv_total, near_total


@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_config1_shape_every_tree(mode):
    # MSN1-5k shape: continuous features (one threshold per distinct value: 16-bit bins), 43 queries
    x, l, off = synth.make_dataset(5000, 136, 43, seed=20260101, gridded=False, qlen=(86, 140))
    _stagewise(x, l, off, leaves=10, ntrees=50, mode=mode, max_near=3)


@pytest.mark.parametrize("mode", [api.HIST_REFERENCE, api.HIST_FAST])
def test_config2_size_first_trees(mode):
    # the benchmarked workload itself (bench.py WORKLOAD)
    x, l, off = synth.make_dataset(1_000_000, 136, 10_000, seed=20260102)
    _stagewise(x, l, off, leaves=64, ntrees=3, mode=mode, max_near=1)
