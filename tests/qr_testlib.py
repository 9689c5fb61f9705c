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

    Walks the two pre-order trees together over the training documents.  Where the split
    (feature, threshold index) differs, the mismatch is tolerated only if
      (a) both candidates cut the node's documents into the SAME two sets (possibly mirrored): a tie
          in exact arithmetic, which the reference breaks by the rounding noise of its cumulative
          sums (plateaus of empty bins in a right child = parent - left, mirrored single-document
          cuts, ...).  The walk continues below with the children matched by document set; or
      (b) the ORACLE's own scores of the two candidates on that node differ by less than `rel`
          relative (a near-tie the reference decided by rounding).  The walk stops there.
    Anything else raises.  Returns (equivalent, near, clean): counts of (a) and (b) and the mask
    of documents whose path never crossed a (b) node."""
    n_docs = bins.shape[1]
    clean = np.ones(n_docs, bool)
    equivalent = near = 0
    stack = [(0, 0, np.arange(n_docs))]
    while stack:
        ig, iw, ids = stack.pop()
        fg, fw = int(got["feature"][ig]), int(want["feature"][iw])
        if fg < 0 and fw < 0:
            continue
        if fg < 0 or fw < 0:
            raise AssertionError("node %d/%d (%d docs): one side is a leaf, the other splits on %d"
                                 % (ig, iw, len(ids), max(fg, fw)))
        tg, tw = int(got["threshold_idx"][ig]), int(want["threshold_idx"][iw])
        lg = bins[fg, ids] <= tg
        kids_g = (int(got["left"][ig]), int(got["right"][ig]))
        kids_w = (int(want["left"][iw]), int(want["right"][iw]))
        if fg == fw and tg == tw:
            stack.append((kids_g[1], kids_w[1], ids[~lg]))
            stack.append((kids_g[0], kids_w[0], ids[lg]))
            continue
        lw = bins[fw, ids] <= tw
        if np.array_equal(lg, lw):
            equivalent += 1
            stack.append((kids_g[1], kids_w[1], ids[~lg]))
            stack.append((kids_g[0], kids_w[0], ids[lg]))
            continue
        if np.array_equal(lg, ~lw):
            equivalent += 1
            stack.append((kids_g[1], kids_w[0], ids[~lg]))
            stack.append((kids_g[0], kids_w[1], ids[lg]))
            continue
        sc = binning.split_scores(lam, ids, minls, [(fg, tg), (fw, tw)])
        gap = abs(sc[0] - sc[1]) / max(abs(sc[1]), 1e-300)
        if not (sc[0] >= 0 and gap <= rel):
            raise AssertionError(
                "node %d (%d docs): got split (f=%d,t=%d) score %.17g, oracle (f=%d,t=%d) score %.17g, "
                "relative gap %.3g > %.1g" % (iw, len(ids), fg, tg, sc[0], fw, tw, sc[1], gap, rel))
        near += 1
        clean[ids] = False
    return equivalent, near, clean


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
