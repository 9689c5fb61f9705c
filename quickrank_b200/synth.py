"""Seeded synthetic learning-to-rank data.

The reference's own fixture (a 5k-row MSLR-WEB10K sample) is cloned at build time
from a host that is unreachable here (reference CMakeLists.txt:80-84), so every
workload in this repo is generated.  The same bytes are fed to the oracle and to
the CUDA path.

Layout returned: features row-major float32 [N, F] (the reference's
``data::Dataset`` layout, include/data/dataset.h:65-66), labels float32 [N],
query offsets uint64 [Q + 1] (``Dataset::offset``), docs of a query contiguous.
"""
from __future__ import annotations

import numpy as np

INFORMATIVE = tuple(range(0, 50, 5))


def query_offsets(n_docs: int, n_queries: int, rng: np.random.Generator,
                  lo: int = 60, hi: int = 140) -> np.ndarray:
    """Query lengths ~ uniform{lo..hi}, adjusted so they sum to n_docs; all < 1000."""
    if n_queries <= 0:
        return np.zeros(1, dtype=np.uint64)
    lens = rng.integers(lo, hi + 1, size=n_queries).astype(np.int64)
    # rescale to the requested total, keep every query non-empty
    lens = np.maximum(1, np.floor(lens * (n_docs / lens.sum())).astype(np.int64))
    diff = int(n_docs - lens.sum())
    i = 0
    while diff != 0:
        step = 1 if diff > 0 else -1
        if lens[i % n_queries] + step >= 1:
            lens[i % n_queries] += step
            diff -= step
        i += 1
    assert lens.sum() == n_docs and lens.min() >= 1
    off = np.zeros(n_queries + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    return off


def make_dataset(n_docs: int, n_features: int, n_queries: int, seed: int = 20260101,
                 gridded: bool = True, levels: int = 256, qlen=(60, 140)):
    """Returns (features[N,F] f32 row-major, labels[N] f32, qoffsets[Q+1] u64).

    gridded=True draws every feature from a ``levels``-level grid so that the number of
    distinct values (and so the reference's threshold count, mart.cc:144-158) is <= levels.
    Features with f % 20 == 19 are constant 0; features 1<->2 and 3<->4 are exact duplicates
    (they exercise the first-maximum tie rule of the split scan, rt.cc:285-306).
    """
    rng = np.random.default_rng(seed)
    off = query_offsets(n_docs, n_queries, rng, qlen[0], qlen[1])
    x = np.empty((n_docs, n_features), dtype=np.float32)
    chunk = max(1, (1 << 24) // max(1, n_features))
    for s in range(0, n_docs, chunk):
        e = min(n_docs, s + chunk)
        u = rng.random((e - s, n_features))
        expo = 1.0 + (np.arange(n_features) % 3)
        v = u ** expo
        if gridded:
            v = np.round((levels - 1) * v) / (levels - 1)
        x[s:e] = v.astype(np.float32)
    for f in range(n_features):
        if f % 20 == 19:
            x[:, f] = 0.0
    if n_features > 2:
        x[:, 2] = x[:, 1]
    if n_features > 4:
        x[:, 4] = x[:, 3]
    inf = [f for f in INFORMATIVE if f < n_features]
    if not inf:
        inf = [0]
    z = x[:, inf].astype(np.float64).mean(axis=1) + 0.08 * rng.standard_normal(n_docs)
    # thresholds follow the mean of z so the label skew is stable across F
    base = float(np.mean(z))
    sd = float(np.std(z))
    cuts = [base + k * sd for k in (1.15, 1.7, 2.2, 2.7)]
    labels = np.zeros(n_docs, dtype=np.float32)
    for c in cuts:
        labels += (z > c).astype(np.float32)
    return np.ascontiguousarray(x), labels, off


def random_ensemble(n_trees: int, n_leaves: int, n_features: int, seed: int = 7,
                    levels: int = 256, weight: float = 0.1):
    """Seeded random forest for the scoring benchmark.

    Returns a list of flat pre-order trees: dict(feature i32[], threshold f32[], left i32[],
    right i32[], value f64[]) plus weights f64[n_trees].  Trees are grown by splitting a random
    current leaf until n_leaves is reached, which gives the unbalanced shapes of leaf-wise growth.
    """
    rng = np.random.default_rng(seed)
    trees = []
    for _ in range(n_trees):
        # build as linked nodes then flatten pre-order
        feat = [-1]
        thr = [0.0]
        left = [-1]
        right = [-1]
        leaves = [0]
        while len(leaves) < n_leaves:
            k = int(rng.integers(len(leaves)))
            node = leaves.pop(k)
            feat[node] = int(rng.integers(n_features))
            thr[node] = float(np.float32(rng.integers(1, levels - 1) / (levels - 1)))
            for side in (left, right):
                side[node] = len(feat)
                feat.append(-1)
                thr.append(0.0)
                left.append(-1)
                right.append(-1)
                leaves.append(len(feat) - 1)
        # pre-order relabel
        order = []
        stack = [0]
        while stack:
            n = stack.pop()
            order.append(n)
            if feat[n] >= 0:
                stack.append(right[n])
                stack.append(left[n])
        new_id = {old: i for i, old in enumerate(order)}
        m = len(order)
        t = dict(feature=np.full(m, -1, np.int32), threshold=np.zeros(m, np.float32),
                 left=np.full(m, -1, np.int32), right=np.full(m, -1, np.int32),
                 value=np.zeros(m, np.float64))
        for old in order:
            i = new_id[old]
            t["feature"][i] = feat[old]
            t["threshold"][i] = thr[old]
            if feat[old] >= 0:
                t["left"][i] = new_id[left[old]]
                t["right"][i] = new_id[right[old]]
            else:
                t["value"][i] = rng.normal(0.0, 0.1)
        trees.append(t)
    weights = np.full(n_trees, weight, dtype=np.float64)
    return trees, weights
