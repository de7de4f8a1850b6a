"""CPU suite, part 3: the N>1 path's host logic with world_size-2 gloo processes.

Each rank takes its nnz-balanced row slab (dasp_partition_rows), x is replicated, the slab product is
computed (here by the oracle standing in for the GPU kernel: this test is about partition + gather
plumbing, not arithmetic) and the y slabs are all-gathered as the power-iteration workload does."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dasp_b200
    import oracle
    from cases import get, x_for

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, n, rp, ci, v = get("powerlaw_20k")
    cuts = dasp_b200.partition_rows(rp, world)
    r0, r1 = int(cuts[rank]), int(cuts[rank + 1])
    # slab CSR with its own zero-based row pointers and GLOBAL column indices
    rp_s = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    ci_s, v_s = ci[rp[r0]:rp[r1]], v[rp[r0]:rp[r1]]
    x = x_for(n)
    ys = []
    for _ in range(3):  # three power-iteration steps: y = A x ; x = y / ||y||
        y_slab = oracle.csr_spmv_f64(r1 - r0, rp_s, ci_s, v_s, x)
        sizes = [int(cuts[p + 1] - cuts[p]) for p in range(world)]
        pad = max(sizes)
        buf = torch.zeros(pad, dtype=torch.float64)
        buf[: r1 - r0] = torch.from_numpy(y_slab)
        out = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, buf)
        y = np.concatenate([o.numpy()[: sizes[p]] for p, o in enumerate(out)])
        nrm2 = torch.tensor([float(np.dot(y_slab, y_slab))], dtype=torch.float64)
        dist.all_reduce(nrm2)
        x = y / np.sqrt(nrm2.item())
        ys.append(y)
    if rank == 0:
        q.put((cuts.tolist(), [a.tolist() for a in ys]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_gather_match_single_process():
    import oracle
    from cases import get, x_for

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    cuts, ys = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, n, rp, ci, v = get("powerlaw_20k")
    assert cuts[0] == 0 and cuts[-1] == m
    nnz = int(rp[m])
    assert abs(int(rp[cuts[1]]) - nnz // 2) <= int(np.diff(rp).max())
    x = x_for(n)
    for step in range(3):
        y = oracle.csr_spmv_f64(m, rp, ci, v, x)
        assert np.array_equal(y, np.array(ys[step])), f"step {step}"
        x = y / np.sqrt(np.dot(y, y))


# ------------------------------------------------------------------------------------------------------------------
# 1.5-D partition of the iterated workload (bench.py --exchange hybrid): short rows by row slab + halo exchange of x, long
# rows split by COLUMNS and merged by one all-reduce.  The plan code (dasp_b200/partition.py) is the code the GPU bench runs;
# here it runs on CPU tensors, two gloo ranks, with the oracle's CSR product standing in for the kernel.

def _skewed_square(m=6000, n_long=12, long_len=1500, window=64, seed=5):
    rng = np.random.default_rng(seed)
    long_rows = set(rng.choice(m, n_long, replace=False).tolist())
    rows = []
    for i in range(m):
        if i in long_rows:
            cols = np.sort(rng.choice(m, long_len, replace=False))
        else:
            k = int(rng.integers(1, 5))
            lo, hi = max(0, i - window), min(m, i + window + 1)
            cols = np.sort(rng.choice(np.arange(lo, hi), k, replace=False))
        rows.append(cols)
    lens = np.array([len(r) for r in rows])
    rp = np.zeros(m + 1, dtype=np.int32)
    np.cumsum(lens, out=rp[1:])
    ci = np.concatenate(rows).astype(np.int32)
    v = rng.uniform(-1, 1, len(ci))
    return m, rp, ci, v


def _hybrid_worker(rank, world, port, q, steps):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dasp_b200
    import oracle
    from dasp_b200 import partition

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, rp, ci, v = _skewed_square()
    cuts = [int(c) for c in dasp_b200.partition_rows(rp, world)]
    r0, r1 = cuts[rank], cuts[rank + 1]
    rows = r1 - r0
    rp_s = torch.from_numpy((rp[r0:r1 + 1] - rp[r0]).astype(np.int32))
    ci_s, v_s = torch.from_numpy(ci[rp[r0]:rp[r1]].copy()), torch.from_numpy(v[rp[r0]:rp[r1]].copy())
    lens = (rp_s[1:] - rp_s[:-1]).long()
    long_local = torch.nonzero(lens >= 256).flatten()
    counts = torch.zeros(world, dtype=torch.int64)
    counts[rank] = long_local.numel()
    dist.all_reduce(counts)
    nl, off = int(counts.sum()), int(counts[:rank].sum())
    glong = torch.zeros(max(nl, 1), dtype=torch.int64)
    glong[off:off + long_local.numel()] = long_local + r0
    dist.all_reduce(glong)
    glong = glong[:nl]
    pieces = []
    for g in glong.tolist():  # every long row of the matrix: its entries inside this rank's column slab
        gc, gv = torch.from_numpy(ci[rp[g]:rp[g + 1]].copy()), torch.from_numpy(v[rp[g]:rp[g + 1]].copy())
        keep = (gc >= r0) & (gc < r1)
        pieces.append((gc[keep], gv[keep]))
    rp_l, ci_l, v_l, cmin, cmax = partition.hybrid_local_matrix(rp_s, ci_s, v_s, r0, r1, long_local, pieces)
    need = torch.zeros(world, 2, dtype=torch.int64)
    need[rank, 0], need[rank, 1] = cmin, cmax + 1
    dist.all_reduce(need)
    need = need.tolist()
    sends, recvs = partition.halo_plan(need, cuts, rank)
    mine = (glong >= r0) & (glong < r1)
    my_pos = torch.nonzero(mine).flatten()
    my_rows = glong[my_pos] - r0

    # x: own slab + what the halo plan delivers; everything else is NaN, so a column the plan forgot poisons the result
    x0 = np.random.default_rng(7).uniform(-1, 1, m)
    x = torch.full((m,), float("nan"), dtype=torch.float64)
    x[r0:r1] = torch.from_numpy(x0[r0:r1])
    lo, hi = need[rank]
    x[lo:hi] = torch.from_numpy(x0[lo:hi])
    rp_n, ci_n, v_n = rp_l.numpy().astype(np.int32), ci_l.numpy().astype(np.int32), v_l.numpy()
    for _ in range(steps):
        y = torch.from_numpy(oracle.csr_spmv_f64(rows + nl, rp_n, ci_n, v_n, np.nan_to_num(x.numpy(), nan=np.nan)))
        red = torch.zeros(nl + 1, dtype=torch.float64)
        red[:nl] = y[rows:]
        red[nl] = (y[:rows] * y[:rows]).sum()
        dist.all_reduce(red)  # the ONE collective of a step
        norm2 = red[nl] + (red[:nl] * red[:nl]).sum()
        y[my_rows] = red[my_pos]
        x[r0:r1] = y[:rows] / torch.sqrt(norm2)
        reqs = [dist.isend(x[a:b].clone(), peer) for peer, a, b in sends]
        bufs = [(torch.empty(b - a, dtype=torch.float64), a, b, peer) for peer, a, b in recvs]
        reqs += [dist.irecv(buf, peer) for buf, a, b, peer in bufs]
        for r in reqs:
            r.wait()
        for buf, a, b, _ in bufs:
            x[a:b] = buf
    out = [torch.zeros(m, dtype=torch.float64) for _ in range(world)]
    slab = torch.zeros(m, dtype=torch.float64)
    slab[r0:r1] = x[r0:r1]
    dist.all_gather(out, slab)
    if rank == 0:
        q.put((cuts, sum(out).tolist(), nl, [len(sends), len(recvs)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_hybrid_partition_matches_single_process_power_iteration():
    import oracle

    world, steps = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_hybrid_worker, args=(r, world, port, q, steps)) for r in range(world)]
    for p in procs:
        p.start()
    cuts, x_par, nl, plan = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, rp, ci, v = _skewed_square()
    assert nl == 12 and plan[0] >= 1 and plan[1] >= 1  # long rows found on both ranks, a halo in each direction
    x = np.random.default_rng(7).uniform(-1, 1, m)
    for _ in range(steps):
        y = oracle.csr_spmv_f64(m, rp, ci, v, x)
        x = y / np.sqrt(np.dot(y, y))
    x_par = np.array(x_par)
    assert np.all(np.isfinite(x_par)), "the halo plan missed a column (NaN reached the iterate)"
    assert np.linalg.norm(x_par - x) <= 1e-12 * np.linalg.norm(x)


def test_halo_plan_is_symmetric():
    from dasp_b200 import partition

    cuts = [0, 100, 250, 400, 1000]
    need = [[0, 130], [80, 260], [240, 420], [390, 1000]]  # every rank reads a little beyond its slab
    plans = [partition.halo_plan(need, cuts, r) for r in range(4)]
    for r in range(4):
        for peer, lo, hi in plans[r][0]:  # what r sends to peer is exactly what peer expects from r
            assert (r, lo, hi) in plans[peer][1]
        for peer, lo, hi in plans[r][1]:
            assert (peer, lo, hi) == (peer, max(cuts[peer], need[r][0]), min(cuts[peer + 1], need[r][1]))
            assert (r, lo, hi) in plans[peer][0]
    assert plans[0] == ([(1, 80, 100)], [(1, 100, 130)])
