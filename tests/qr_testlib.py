"""Shared helpers for the parity tests."""
import numpy as np

from quickrank_b200 import synth

STRUCT_FIELDS = ("feature", "threshold_idx", "threshold", "left", "right")


def dataset(n=3000, f=20, q=30, seed=11, gridded=True, qlen=(60, 140)):
    x, l, off = synth.make_dataset(n, f, q, seed=seed, gridded=gridded, qlen=qlen)
    return x, l, off


def tie_heavy_scores(n, rng, levels=7):
    return rng.integers(0, levels, size=n).astype(np.float64) * 0.25


def same_structure(a, b):
    return all(np.array_equal(a[k], b[k]) for k in STRUCT_FIELDS)


def describe_tree_diff(a, b):
    out = []
    for k in STRUCT_FIELDS + ("value", "count", "deviance"):
        if k in a and k in b and not np.array_equal(a[k], b[k]):
            if len(a[k]) != len(b[k]):
                out.append("%s: length %d vs %d" % (k, len(a[k]), len(b[k])))
            else:
                idx = np.nonzero(a[k] != b[k])[0]
                out.append("%s differs at nodes %s: %s vs %s" % (k, idx[:6], a[k][idx[:6]], b[k][idx[:6]]))
    return "; ".join(out)


def leaves_mask(t):
    return t["feature"] < 0


def audit_tree(got, want, binning, bins, lam, minls, rel=1e-12):
    """Tie audit (hard parts 2 and 6).

    Walks the two pre-order trees together.  Where the split (feature, threshold index) differs,
    the mismatch is tolerated only if the ORACLE's own score of the two candidates on that node
    differs by less than `rel` relative, i.e. the reference's choice was decided by rounding.
    Returns (tolerated, docs_compared): the number of tolerated ties and a boolean mask of the
    documents whose path never crossed a tolerated node (their leaves must agree exactly)."""
    n_docs = bins.shape[1]
    clean = np.ones(n_docs, bool)
    tolerated = 0
    stack = [(0, 0, np.arange(n_docs))]
    while stack:
        ig, iw, ids = stack.pop()
        fg, fw = int(got["feature"][ig]), int(want["feature"][iw])
        if fg < 0 and fw < 0:
            continue
        tg = int(got["threshold_idx"][ig]) if fg >= 0 else -1
        tw = int(want["threshold_idx"][iw]) if fw >= 0 else -1
        if fg == fw and tg == tw:
            left = bins[fg, ids] <= tg
            stack.append((int(got["right"][ig]), int(want["right"][iw]), ids[~left]))
            stack.append((int(got["left"][ig]), int(want["left"][iw]), ids[left]))
            continue
        if fg < 0 or fw < 0:
            raise AssertionError("node %d/%d: one side is a leaf, the other splits (%d,%d) vs (%d,%d)"
                                 % (ig, iw, fg, tg, fw, tw))
        sc = binning.split_scores(lam, ids, minls, [(fg, tg), (fw, tw)])
        gap = abs(sc[0] - sc[1]) / max(abs(sc[1]), 1e-300)
        if not (sc[0] >= 0 and gap <= rel):
            raise AssertionError(
                "node %d (%d docs): got split (f=%d,t=%d) score %.17g, oracle (f=%d,t=%d) score %.17g, "
                "relative gap %.3g > %.1g" % (iw, len(ids), fg, tg, sc[0], fw, tw, sc[1], gap, rel))
        tolerated += 1
        clean[ids] = False
    return tolerated, clean


def tree_outputs(tree, bins):
    """Leaf output of every document, walking the flat tree on bins."""
    n = bins.shape[1]
    node = np.zeros(n, np.int64)
    active = tree["feature"][node] >= 0
    while active.any():
        f = tree["feature"][node[active]]
        t = tree["threshold_idx"][node[active]]
        idx = np.nonzero(active)[0]
        go_left = bins[f, idx] <= t
        node[idx] = np.where(go_left, tree["left"][node[idx]], tree["right"][node[idx]])
        active = tree["feature"][node] >= 0
    return tree["value"][node]
