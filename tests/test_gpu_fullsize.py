"""GPU: BASELINE.json's configurations at FULL size, checked through size-independent properties (the oracle
cannot run 10^8-10^9 entries in a unit test):
  * order_rid is a permutation; all-ones matrix and vector => y_perm[k] == length of row order_rid[k] (the
    reference's shipped run, src/main_f64.cu:131-132), exact in FP64 up to 2^53;
  * checksum: sum(y) == sum_k val[k] * x[col[k]] computed by an independent torch expression;
  * y in original order == y in permuted order scattered through order_rid, bit for bit;
  * a row slab preprocessed on its own (the multi-GPU path) reproduces the corresponding rows of the full product."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spec(name):
    from dasp_b200 import synth

    return {"c4": lambda: synth.stencil27(256), "c3": lambda: synth.powerlaw(),
            "c5_quarter": lambda: synth.skewed(n_long=250, n_short=12_500_000)}[name]()


@pytest.mark.parametrize("name", ["c4", "c3", "c5_quarter"])
def test_full_size_properties(cuda_device, name):
    import torch

    import dasp_b200
    from dasp_b200 import synth

    spec = _spec(name)
    m, n = int(spec.m), int(spec.n)
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device)
    s = torch.cuda.current_stream().cuda_stream
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v, nnz=nnz)
    st = h.stats()
    assert st["nnz_long"] + st["nnz_short"] + st["origin_nnz_reg"] + st["nnz_irreg"] == nnz
    order = torch.from_numpy(h.export("order_rid")).to(cuda_device).long()
    assert int(torch.bincount(order, minlength=m).max().item()) == 1 and order.numel() == m

    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(3)
    x = torch.rand(n, generator=gen, device=cuda_device, dtype=torch.float64) * 2 - 1
    y = torch.full((m,), float("nan"), dtype=torch.float64, device=cuda_device)
    yo = torch.full((m,), float("nan"), dtype=torch.float64, device=cuda_device)
    h.spmv(x, y, s)
    h.spmv_unpermuted(x, yo, s)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(y).all()) and bool(torch.equal(yo[order], y))
    # checksum against an independent evaluation, chunked to bound memory
    total, scale = 0.0, 0.0
    for a in range(0, nnz, 1 << 27):
        b = min(nnz, a + (1 << 27))
        p = v[a:b] * x[ci[a:b].long()]
        total += float(p.sum().item())
        scale += float(p.abs().sum().item())
    assert abs(float(y.sum().item()) - total) <= 1e-11 * scale
    h.close()

    # all ones: y_perm == row lengths
    v.fill_(1.0)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v, nnz=nnz)
    ones = torch.ones(n, dtype=torch.float64, device=cuda_device)
    h.spmv(ones, y, s)
    torch.cuda.synchronize()
    lens = (rp[1:] - rp[:-1]).double()
    assert bool(torch.equal(y, lens[order]))
    h.close()


def test_row_slab_equals_rows_of_the_full_product(cuda_device):
    import torch

    import dasp_b200
    from dasp_b200 import synth

    spec = synth.stencil27(160)
    m = int(spec.m)
    s = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(5)
    x = torch.rand(m, generator=gen, device=cuda_device, dtype=torch.float64) * 2 - 1
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device)
    cuts = dasp_b200.partition_rows(rp.cpu().numpy(), 3)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, m, rp, ci, v, nnz=nnz)
    y_full = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    h.spmv_unpermuted(x, y_full, s)
    torch.cuda.synchronize()
    h.close()
    for p in range(3):
        r0, r1 = int(cuts[p]), int(cuts[p + 1])
        rps, cis, vs, k = synth.generate(spec, r0, r1, cuda_device)
        hs = dasp_b200.Dasp(dasp_b200.DASP_F64, r1 - r0, m, rps, cis, vs, nnz=k)
        ys = torch.zeros(r1 - r0, dtype=torch.float64, device=cuda_device)
        hs.spmv_unpermuted(x, ys, s)
        torch.cuda.synchronize()
        assert bool(torch.equal(ys, y_full[r0:r1])), f"slab {p}"
        hs.close()
