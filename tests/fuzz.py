"""Seeded random CSR matrices with adversarial row-length mixes, shared by the CPU and GPU fuzz tests."""
import numpy as np


def random_case(seed):
    rng = np.random.default_rng(seed)
    m = int(rng.integers(1, 3000))
    n = int(rng.integers(1, 200000)) if rng.random() < 0.3 else int(rng.integers(1, 5000))
    kind = rng.integers(0, 5)
    if kind == 0:      # short rows dominate
        lens = rng.choice([0, 1, 2, 3, 4], m, p=[0.05, 0.35, 0.2, 0.25, 0.15])
    elif kind == 1:    # medium rows around the tile / block_longest edges
        lens = rng.choice([4, 5, 6, 7, 8, 9, 23, 24, 25, 31, 32, 33, 63, 64, 65, 127, 255, 256, 257], m)
    elif kind == 2:    # heavy tail
        lens = np.minimum((rng.pareto(0.9, m) + 1).astype(np.int64), 5000)
    elif kind == 3:    # everything
        lens = rng.integers(0, 40, m)
        lens[rng.integers(0, m, max(1, m // 50))] = rng.integers(256, 3000)
    else:              # almost empty
        lens = (rng.random(m) < 0.1).astype(np.int64) * rng.integers(1, 300, m)
    lens = np.minimum(lens, n).astype(np.int64)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    # columns: duplicates allowed (the reference keeps them), not sorted
    colidx = rng.integers(0, n, nnz).astype(np.int32)
    val = rng.uniform(-1, 1, nnz)
    threshold = float(rng.choice([0.75, 0.75, 0.5, 1.0, 0.25, 0.9]))
    block_longest = int(rng.choice([256, 256, 64, 1000, 17, 5]))
    return m, n, rowptr.astype(np.int32), colidx, val, threshold, block_longest
