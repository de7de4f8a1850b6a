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
