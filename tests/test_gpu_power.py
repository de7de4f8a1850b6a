"""GPU: the iterated workload's building blocks (y in original order written into a slab of the next x,
device-side norm and scaling) against a numpy power iteration on the oracle's CSR product."""
import numpy as np
import pytest

import oracle
from cases import get, x_for

pytestmark = pytest.mark.gpu


def test_power_iteration_matches_oracle(cuda_device):
    import torch

    import dasp_b200

    m, n, rp, ci, v = get("symmetric_like")
    x0 = x_for(n)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    s = torch.cuda.current_stream().cuda_stream
    xa = torch.from_numpy(x0).to(cuda_device)
    xb = torch.zeros_like(xa)
    norm2 = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    x_ref = x0.copy()
    for step in range(6):
        h.spmv_unpermuted(xa, xb, s)
        dasp_b200.sumsq(xb, m, norm2, s)
        dasp_b200.scale_rsqrt(xb, m, norm2, s)
        xa, xb = xb, xa
        y = oracle.csr_spmv_f64(m, rp, ci, v, x_ref)
        nrm = np.sqrt(np.dot(y, y))
        x_ref = y / nrm
        torch.cuda.synchronize()
        assert abs(np.sqrt(norm2.item()) - nrm) <= 1e-12 * nrm
        got = xa.cpu().numpy()
        assert np.linalg.norm(got - x_ref) <= 1e-12 * np.linalg.norm(x_ref), f"step {step}"
    h.close()


def test_sumsq_large_and_empty(cuda_device):
    import torch

    import dasp_b200

    v = torch.arange(1, 300001, dtype=torch.float64, device=cuda_device) * 1e-3
    out = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    dasp_b200.sumsq(v, v.numel(), out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = float((v.cpu().numpy() ** 2).sum())
    assert abs(out.item() - want) <= 1e-12 * want
    dasp_b200.sumsq(v, 0, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert out.item() == 0.0


def test_scatter_to_several_destinations(cuda_device):
    """The fused-exchange entry on one GPU: the slab product lands, scaled, at row_offset of every destination."""
    import torch

    import dasp_b200

    m, n, rp, ci, v = get("symmetric_like")
    x0 = x_for(n)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    s = torch.cuda.current_stream().cuda_stream
    dx = torch.from_numpy(x0).to(cuda_device)
    norm2 = torch.tensor([6.25], dtype=torch.float64, device=cuda_device)
    off = 17
    dests = [torch.full((m + 40,), -7.0, dtype=torch.float64, device=cuda_device) for _ in range(3)]
    h.spmv_scatter_to(dx, dests, off, norm2, s)
    torch.cuda.synchronize()
    want = oracle.csr_spmv_f64(m, rp, ci, v, x0) / 2.5
    for d in dests:
        got = d.cpu().numpy()
        assert np.all(got[:off] == -7.0) and np.all(got[off + m:] == -7.0)
        assert np.linalg.norm(got[off:off + m] - want) <= 1e-12 * np.linalg.norm(want)
    h.spmv_scatter_to(dx, dests[:1], 0, None, s)
    torch.cuda.synchronize()
    assert np.linalg.norm(dests[0].cpu().numpy()[:m] - want * 2.5) <= 1e-12 * np.linalg.norm(want * 2.5)
    h.close()


def test_unpermute_to_matches_unpermuted_product(cuda_device):
    import torch

    import dasp_b200

    m, n, rp, ci, v = get("mixed_f1")
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    s = torch.cuda.current_stream().cuda_stream
    dx = torch.from_numpy(x_for(n)).to(cuda_device)
    yp = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    yo = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    h.spmv(dx, yp, s)
    h.spmv_unpermuted(dx, yo, s)
    norm2 = torch.tensor([16.0], dtype=torch.float64, device=cuda_device)
    dests = [torch.full((m + 9,), 3.0, dtype=torch.float64, device=cuda_device) for _ in range(2)]
    for rep in range(2):
        h.unpermute_to(yp, dests, 5, norm2, s)
    torch.cuda.synchronize()
    for d in dests:
        assert bool(torch.equal(d[5:5 + m], yo * 0.25)) and bool((d[:5] == 3.0).all()) and bool((d[5 + m:] == 3.0).all())
    h.close()


@pytest.mark.parametrize("count,offset", [(1000, 6), (1001, 6), (1000, 7), (1, 0), (0, 3)])
def test_scale_copy_to(cuda_device, count, offset):
    import torch

    import dasp_b200

    s = torch.cuda.current_stream().cuda_stream
    v = torch.arange(1, count + 1, dtype=torch.float64, device=cuda_device)
    norm2 = torch.tensor([4.0], dtype=torch.float64, device=cuda_device)
    dests = [torch.full((count + 16,), -1.0, dtype=torch.float64, device=cuda_device) for _ in range(3)]
    dasp_b200.scale_copy_to(v if count else dests[0], count, dests, offset, norm2, s)
    torch.cuda.synchronize()
    for d in dests:
        assert bool(torch.equal(d[offset:offset + count], v * 0.5))
        assert bool((d[:offset] == -1.0).all()) and bool((d[offset + count:] == -1.0).all())


def test_two_handles_on_two_streams_interleaved(cuda_device):
    """Handles are independent: two matrices multiplied and reduced on two streams at the same time."""
    import torch

    import dasp_b200

    cases_ = [get("powerlaw_20k"), get("mixed_f1")]
    hs = [dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v) for (m, n, rp, ci, v) in cases_]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    xs = [torch.from_numpy(x_for(c[1], seed=9 + i)).to(cuda_device) for i, c in enumerate(cases_)]
    ys = [torch.zeros(c[0], dtype=torch.float64, device=cuda_device) for c in cases_]
    n2 = [torch.zeros(1, dtype=torch.float64, device=cuda_device) for _ in cases_]
    torch.cuda.synchronize()
    for rep in range(20):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                hs[i].spmv_unpermuted(xs[i], ys[i], streams[i].cuda_stream)
                dasp_b200.sumsq(ys[i], cases_[i][0], n2[i], streams[i].cuda_stream)
    torch.cuda.synchronize()
    for i, (m, n, rp, ci, v) in enumerate(cases_):
        y_ref = oracle.csr_spmv_f64(m, rp, ci, v, xs[i].cpu().numpy())
        assert np.linalg.norm(ys[i].cpu().numpy() - y_ref) <= 1e-12 * np.linalg.norm(y_ref)
        assert abs(n2[i].item() - float(np.dot(y_ref, y_ref))) <= 1e-11 * float(np.dot(y_ref, y_ref))
    for h in hs:
        h.close()


@pytest.mark.parametrize("name", ["mixed_f1", "symmetric_like", "powerlaw_20k_square"])
@pytest.mark.parametrize("long_variant", ["chunked", "blocked"])
def test_relabelled_mode_equals_the_plain_product(cuda_device, name, long_variant):
    """P*A*P^T mode (dasp_relabel_columns with the inverse permutation): x given in PERMUTED order, y produced in permuted
    order; equals the plain product through order_rid, and a power iteration that feeds y_perm straight back as x
    (dasp_spmv_permuted_to, no scatter, no un-permute pass) reproduces the eigenvalue estimate and the iterate of the
    unpermuted path.  Two slabs with a global relabelling reproduce the single-handle product."""
    import torch

    import dasp_b200
    import matrices

    if name == "powerlaw_20k_square":
        m, n, rp, ci, v = matrices.powerlaw(m=20000, lmax=5000)
    else:
        m, n, rp, ci, v = get(name)
    if m != n:  # square it: columns folded into [0, m)
        ci = (ci % m).astype(np.int32)
        n = m
    lv = dasp_b200.VARIANT_BLOCKED if long_variant == "blocked" else dasp_b200.VARIANT_CUDA_CORE
    s = torch.cuda.current_stream().cuda_stream
    x0 = x_for(n)
    y_ref = oracle.csr_spmv_f64(m, rp, ci, v, x0)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    h.set_variant(0, lv, 0)
    order = h.export("order_rid")
    inv = np.empty(m, dtype=np.int32)
    inv[order] = np.arange(m, dtype=np.int32)
    d_inv = torch.from_numpy(inv).to(cuda_device)
    before = {a: h.export(a) for a in ("reg_cid", "long_cid", "short_cid", "irreg_cid")}
    h.relabel_columns(d_inv, m)
    for a, arr in before.items():  # the reference arrays are untouched by the relabelling
        assert np.array_equal(h.export(a), arr), a
    xp = torch.from_numpy(x0[order]).to(cuda_device)
    yp = torch.full((m,), float("nan"), dtype=torch.float64, device=cuda_device)
    for rep in range(2):
        h.spmv(xp, yp, s)
        torch.cuda.synchronize()
        got = yp.cpu().numpy()
        assert np.linalg.norm(got - y_ref[order]) <= 1e-12 * np.linalg.norm(y_ref), (name, rep)
    # power iteration entirely in permuted space
    xa, xb = xp.clone(), torch.zeros_like(xp)
    nrm2 = [torch.ones(1, dtype=torch.float64, device=cuda_device), torch.ones(1, dtype=torch.float64, device=cuda_device)]
    x_ref = x0.copy()
    lam_ref = 0.0
    for k in range(5):
        # x_k is stored un-normalised; the product scales by 1/sqrt(||x_k||^2) read on the device
        h.spmv_permuted_to(xa, [xb], 0, nrm2[k & 1], s)
        dasp_b200.sumsq(xb, m, nrm2[(k + 1) & 1], s)
        xa, xb = xb, xa
        y = oracle.csr_spmv_f64(m, rp, ci, v, x_ref)
        lam_ref = np.sqrt(np.dot(y, y))
        x_ref = y / lam_ref
    torch.cuda.synchronize()
    lam = float(torch.sqrt(nrm2[5 & 1]).item())
    assert abs(lam - lam_ref) <= 1e-11 * lam_ref
    got = (xa / lam).cpu().numpy()
    assert np.linalg.norm(got - x_ref[order]) <= 1e-11
    h.close()

    # two row slabs, global relabelling: new index of column j = slab offset of its owner + that slab's inverse order
    cuts = dasp_b200.partition_rows(rp, 2)
    hs, invs = [], []
    for p in range(2):
        r0, r1 = int(cuts[p]), int(cuts[p + 1])
        rps = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
        hp = dasp_b200.Dasp(dasp_b200.DASP_F64, r1 - r0, n, rps, ci[rp[r0]:rp[r1]], v[rp[r0]:rp[r1]])
        hp.set_variant(0, lv, 0)
        o = hp.export("order_rid")
        iv = np.empty(r1 - r0, dtype=np.int32)
        iv[o] = np.arange(r1 - r0, dtype=np.int32)
        hs.append((hp, o, r0, r1))
        invs.append(iv + r0)
    gmap = torch.from_numpy(np.concatenate(invs).astype(np.int32)).to(cuda_device)
    gorder = np.concatenate([o + r0 for (_, o, r0, _) in hs])
    xg = torch.from_numpy(x0[gorder]).to(cuda_device)
    yg = torch.full((m,), float("nan"), dtype=torch.float64, device=cuda_device)
    for hp, o, r0, r1 in hs:
        hp.relabel_columns(gmap, m)
        hp.spmv_permuted_to(xg, [yg], r0, None, s)
    torch.cuda.synchronize()
    assert np.linalg.norm(yg.cpu().numpy() - y_ref[gorder]) <= 1e-12 * np.linalg.norm(y_ref)
    for hp, *_ in hs:
        hp.close()
