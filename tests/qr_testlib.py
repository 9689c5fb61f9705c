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
