"""Named parity cases shared by the CPU (oracle) and GPU (C-ABI) tests.  All are small enough for
the oracle to finish in well under a second each."""
from __future__ import annotations

import numpy as np

import matrices as M


def _lens(lens, n=None, seed=11, window=None):
    lens = np.asarray(lens, dtype=np.int64)
    n = n or max(int(lens.max(initial=1)) + 16, len(lens), 64)
    return M.from_lengths(lens, n, seed, window=window)


def _pairs(k):
    """k one-rows and k three-rows interleaved with a few others: around the common_13 >= 128 switch"""
    rng = np.random.default_rng(k)
    lens = np.array([1] * k + [3] * k + [2] * 37 + [4] * 21 + [0] * 5 + [7] * 40)
    rng.shuffle(lens)
    return _lens(lens, seed=k)


def _wide_mixed():
    """Medium rows whose columns sit in a narrow window (compressible tiles), rows spanning > 65535 columns (wide
    blocks), rows that contain column 0, and a row with columns exactly 65534 / 65535 / 65536 apart."""
    rng = np.random.default_rng(33)
    n = 200000
    rows = []
    for i in range(400):
        L = int(rng.integers(5, 60))
        if i % 7 == 0:
            cols = rng.choice(n, L, replace=False)                       # wide
        elif i % 7 == 1:
            cols = np.concatenate([[0], 1 + rng.choice(3000, L - 1, replace=False)])  # column 0 + narrow
        else:
            lo = int(rng.integers(0, n - 4000))
            cols = lo + rng.choice(4000, L, replace=False)               # narrow
        rows.append(cols)
    for span in (65534, 65535, 65536):
        rows += [np.array([1000, 1000 + span, 1001, 1002, 1003, 1004, 1005, 1006])] * 8
    lens = np.array([len(r) for r in rows])
    rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = np.concatenate(rows).astype(np.int32)
    val = rng.uniform(-1, 1, len(colidx))
    return len(rows), n, rowptr.astype(np.int32), colidx, val


def _lcb_row_deltas():
    """Long rows for the column-blocked kernel's 16-bit index stream: 260 long rows over 30 column blocks of 8192, row i
    living mostly in blocks i % 13 and (i + 5) % 13 + 13, so that inside a block consecutive entries jump 13 rows (chunks
    flagged wide, read through the 32-bit indices), plus 40 long rows spread over all columns (small deltas) and a few medium /
    short rows."""
    rng = np.random.default_rng(77)
    n = 30 * 8192
    rows = []
    for i in range(260):
        b0, b1 = i % 13, (i + 5) % 13 + 13
        cols = np.concatenate([b0 * 8192 + rng.choice(8192, 180, replace=False), b1 * 8192 + rng.choice(8192, 150, replace=False)])
        rows.append(np.sort(cols))
    for i in range(40):
        rows.append(np.sort(rng.choice(n, 300 + 7 * i, replace=False)))
    for i in range(50):
        rows.append(np.sort(rng.choice(n, int(rng.integers(1, 40)), replace=False)))
    order = rng.permutation(len(rows))
    rows = [rows[k] for k in order]
    lens = np.array([len(r) for r in rows])
    rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = np.concatenate(rows).astype(np.int32)
    val = rng.uniform(-1, 1, len(colidx))
    return len(rows), n, rowptr.astype(np.int32), colidx, val


CASES = {
    # every category populated (fixture F1 of SURVEY.md Appendix A)
    "mixed_f1": lambda: M.mixed(),
    "stencil27_12": lambda: M.stencil27(12),
    "stencil27_9x7x5": lambda: M.stencil27(9, 7, 5),
    "powerlaw_20k": lambda: M.powerlaw(),
    "skewed_long5000": lambda: M.skewed(),
    "symmetric_like": lambda: M.symmetric_like(),
    # category edges
    "only_1": lambda: M.only_lengths(1, 700),
    "only_2": lambda: M.only_lengths(2, 333),
    "only_3": lambda: M.only_lengths(3, 129),
    "only_4": lambda: M.only_lengths(4, 1001),
    "only_5": lambda: M.only_lengths(5, 77),
    "len_255_256": lambda: _lens([255] * 9 + [256] * 7 + [257] * 3 + [4, 5] * 10, n=4096),
    "pairs_127": lambda: _pairs(127),
    "pairs_128": lambda: _pairs(128),
    "pairs_129": lambda: _pairs(129),
    "pairs_1000_vs_300": lambda: _lens([1] * 1000 + [3] * 300 + [2] * 65, seed=3),
    "empty_rows_only": lambda: _lens([0] * 50, n=64),
    "zero_rows_mixed": lambda: _lens([0, 9, 0, 0, 300, 1, 0, 2, 0, 3, 4, 0] * 30, n=2048),
    "one_row_70000": lambda: _lens([70000, 3, 1, 12], n=80000, seed=2),
    "long_pad_64k": lambda: _lens([64, 65, 128, 129, 4096, 4097] * 3 + [300] * 50, n=8192, seed=4),
    "ragged_tail_blocks": lambda: _lens(list(range(5, 90)) + [200, 199, 17] * 11, n=1024, seed=6),
    # column spans around the 16-bit limit of the compact index form: every tile wide / mixed / column 0 present
    "wide_span_all": lambda: _lens([40] * 200 + [7] * 30, n=300000, seed=21),
    "wide_span_mixed": lambda: _wide_mixed(),
    "lcb_row_deltas": lambda: _lcb_row_deltas(),
    "rowloop_59989": lambda: _lens([5] * 59989 + [1, 2, 3], n=70000, seed=8, window=64),
    "rowloop_59990": lambda: _lens([5] * 59990 + [1, 2, 3], n=70000, seed=8, window=64),
    "rowloop_399999": lambda: _lens([6] * 399999, n=400000, seed=9, window=64),
    "rowloop_400000": lambda: _lens([6] * 400000, n=400000, seed=9, window=64),
}

_cache = {}


def get(name):
    """Generate a case once per process (the big ones take seconds in numpy)."""
    if name not in _cache:
        _cache[name] = CASES[name]()
    return _cache[name]


# cheap subset for the no-GPU suite (the big rowloop cases only matter for array lengths)
CPU_CASES = [k for k in CASES if k not in ("rowloop_400000", "rowloop_399999")]


def x_for(n, seed=7):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)
