"""Host-side plan of the 1.5-D partition of the iterated workload (bench.py --exchange hybrid).

Every rank owns a row slab [r0, r1) AND the matching column slab of x.  Short / medium rows stay with their row slab; the
LONG rows of the whole matrix are split by COLUMNS: rank p multiplies, for every long row, the entries whose columns lie in
its slab (they become ordinary long rows of its local DASP matrix) and the partial sums are merged by one small all-reduce.
Pure torch, device-agnostic: the GPU bench and the two-rank gloo test (tests/test_partition_gloo.py) run the same code.
"""
from __future__ import annotations

import torch


def hybrid_local_matrix(rp, ci, v, r0, r1, long_local, pieces):
    """Local matrix of one rank.

    rp, ci, v   : the rank's row slab as CSR (rp zero-based, GLOBAL column indices), torch tensors on one device
    long_local  : int64 tensor, slab-local indices of the slab's long rows
    pieces      : for EVERY long row of the matrix, in ascending global row order, a pair (cols, vals) holding the entries of
                  that row whose columns lie in [r0, r1)
    Returns (rp_l, ci_l, v_l, cmin, cmax): CSR of `rows + len(pieces)` rows - the slab's rows with the long rows emptied, then
    one row per long-row piece - and the column range the short part reads (for the halo plan)."""
    dev = rp.device
    rows = r1 - r0
    lens = (rp[1:] - rp[:-1]).long()
    row_of = torch.repeat_interleave(torch.arange(rows, device=dev), lens)
    is_long_row = torch.zeros(rows, dtype=torch.bool, device=dev)
    is_long_row[long_local] = True
    keep = ~is_long_row[row_of]
    del row_of
    lens_local = torch.where(is_long_row, torch.zeros_like(lens), lens)
    plen = [int(c.numel()) for c, _ in pieces]
    all_lens = torch.cat([lens_local, torch.tensor(plen, dtype=torch.int64, device=dev)])
    rp_l = torch.zeros(rows + len(pieces) + 1, dtype=torch.int64, device=dev)
    torch.cumsum(all_lens, 0, out=rp_l[1:])
    cshort = ci[keep]
    ci_l = torch.cat([cshort] + [c for c, _ in pieces])
    v_l = torch.cat([v[keep]] + [w for _, w in pieces])
    cmin = int(cshort.min().item()) if cshort.numel() else r0
    cmax = int(cshort.max().item()) if cshort.numel() else r0
    return rp_l, ci_l, v_l, cmin, cmax


def halo_plan(need, cuts, rank):
    """need[q] = [lo, hi) of the columns of x rank q's short part reads; cuts = slab boundaries.  Returns (sends, recvs):
    lists of (peer, lo, hi) in global indices - what this rank sends from its own slab, and what it receives from whom."""
    r0, r1 = cuts[rank], cuts[rank + 1]
    sends, recvs = [], []
    for q in range(len(cuts) - 1):
        if q == rank:
            continue
        lo, hi = max(r0, need[q][0]), min(r1, need[q][1])  # what q needs from my slab
        if lo < hi:
            sends.append((q, lo, hi))
        lo, hi = max(cuts[q], need[rank][0]), min(cuts[q + 1], need[rank][1])  # what I need from q's slab
        if lo < hi:
            recvs.append((q, lo, hi))
    return sends, recvs
