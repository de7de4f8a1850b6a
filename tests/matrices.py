"""Small seeded CSR generators shared by the tests (numpy twins of the bench generators).

Every generator returns (m, n, rowptr[int32], colidx[int32], val[float64]).  Column order inside a
row is deliberately NOT sorted unless stated: the reference keeps CSR order.
"""
from __future__ import annotations

import numpy as np


def from_lengths(lens, n, seed, window=None):
    """Rows with the given lengths; distinct random columns per row (optionally inside a window
    around the diagonal), values U(-1,1)."""
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int64)
    m = len(lens)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    colidx = np.empty(nnz, dtype=np.int32)
    for i, l in enumerate(lens):
        if l == 0:
            continue
        if window is None or l > 2 * window:
            cols = rng.choice(n, int(l), replace=False)
        else:
            c = int(i * n / max(m, 1))
            lo = max(0, min(c - window, n - 2 * window))
            cols = lo + rng.choice(min(2 * window, n), int(l), replace=False)
        colidx[rowptr[i]:rowptr[i + 1]] = cols
    val = rng.uniform(-1.0, 1.0, nnz)
    return m, n, rowptr.astype(np.int32), colidx, val


def mixed(seed=1234, n_rows=4000, n_cols=5000, zeros=37, ones=300, twos=217, threes=450, fours=333,
          medium=2600, longs=63, long_max=1500):
    """Fixture F1 of SURVEY.md Appendix A: every category populated, rows shuffled."""
    rng = np.random.default_rng(seed)
    lens = ([0] * zeros + [1] * ones + [2] * twos + [3] * threes + [4] * fours
            + list(rng.integers(5, 256, medium)) + list(rng.integers(256, long_max, longs)))
    lens = np.array(lens[:n_rows] if len(lens) > n_rows else lens, dtype=np.int64)
    rng.shuffle(lens)
    return from_lengths(lens, n_cols, seed + 1)


def stencil27(nx, ny=None, nz=None, seed=20240004):
    """27-point stencil on an nx*ny*nz grid, natural ordering (x fastest), columns ascending."""
    ny = ny or nx
    nz = nz or nx
    m = nx * ny * nz
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # row index = x + nx*(y + ny*z): iterate z slowest
    X = ix.transpose(2, 1, 0).ravel()
    Y = iy.transpose(2, 1, 0).ravel()
    Z = iz.transpose(2, 1, 0).ravel()
    cols = []
    valid = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                x2, y2, z2 = X + dx, Y + dy, Z + dz
                ok = (x2 >= 0) & (x2 < nx) & (y2 >= 0) & (y2 < ny) & (z2 >= 0) & (z2 < nz)
                cols.append(np.where(ok, x2 + nx * (y2 + ny * z2), 0))
                valid.append(ok)
    cols = np.stack(cols, axis=1)
    valid = np.stack(valid, axis=1)
    lens = valid.sum(axis=1)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = cols[valid].astype(np.int32)
    rng = np.random.default_rng(seed)
    val = rng.uniform(-1.0, 1.0, int(rowptr[-1]))
    return m, m, rowptr.astype(np.int32), colidx, val


def powerlaw(m=20000, alpha=0.95, lmax=20000, seed=20240001, window=512):
    """Config C3 scaled down: L = min(floor(u^(-1/alpha)), lmax); interleaved categories."""
    rng = np.random.default_rng(seed)
    u = 1.0 - rng.random(m)
    lens = np.minimum(np.floor(u ** (-1.0 / alpha)), min(lmax, m)).astype(np.int64)
    return from_lengths(lens, m, seed + 1, window=window)


def skewed(n_long=6, long_len=5000, n_short=30000, seed=20240005, window=512):
    """Config C5 scaled down: a few very long rows first, then short rows with L in {1,2,3,4}."""
    rng = np.random.default_rng(seed)
    m = n_long + n_short
    lens = np.concatenate([np.full(n_long, long_len), rng.integers(1, 5, n_short)])
    return from_lengths(lens, m, seed + 1, window=window)


def only_lengths(length, count, n=None, seed=5):
    """count rows all of the same length."""
    n = n or max(count, length + 8, 64)
    return from_lengths(np.full(count, length), n, seed)


def symmetric_like(m=3000, mean_len=21, seed=7):
    """cop20k_A stand-in scaled down: banded-ish symmetric pattern, row lengths ~ 5..80."""
    rng = np.random.default_rng(seed)
    rows = []
    cols = []
    for i in range(m):
        k = int(rng.integers(2, mean_len))
        j = np.unique(np.clip(i + rng.integers(-300, 301, k), 0, m - 1))
        rows.append(np.full(len(j), i))
        cols.append(j)
    r = np.concatenate(rows + cols)
    c = np.concatenate(cols + rows)
    key = np.unique(r.astype(np.int64) * m + c)
    r, c = key // m, key % m
    lens = np.bincount(r, minlength=m)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    val = rng.uniform(-1.0, 1.0, len(c))
    return m, m, rowptr.astype(np.int32), c.astype(np.int32), val
