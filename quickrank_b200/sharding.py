"""Query-aligned document sharding for multi-GPU training: lambdas need
all documents of a query on one rank (reference lambdamart.cc:71-151), so ranks own contiguous
query ranges balanced by document count."""
from __future__ import annotations

import numpy as np


def query_shards(qoff, world: int):
    """Returns [(q_begin, q_end)] * world: contiguous query ranges whose document counts are as
    close as possible to N / world.  Every rank gets at least one query when Q >= world."""
    qoff = np.asarray(qoff, dtype=np.int64)
    nq = len(qoff) - 1
    if world < 1:
        raise ValueError("world must be >= 1")
    if nq < world:
        raise ValueError("cannot shard %d queries over %d ranks" % (nq, world))
    n = int(qoff[-1])
    bounds = [0]
    for r in range(1, world):
        target = n * r / world
        q = int(np.searchsorted(qoff, target, side="left"))
        # pick the boundary closest to the target
        if q > 0 and abs(qoff[q - 1] - target) <= abs(qoff[min(q, nq)] - target):
            q -= 1
        q = max(q, bounds[-1] + 1)
        q = min(q, nq - (world - r))
        bounds.append(q)
    bounds.append(nq)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]
