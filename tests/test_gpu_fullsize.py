"""GPU: BASELINE.json's configurations at FULL size against the oracle.

For every benchmarked shape (and for both the CSR-sorted and the SURVEY §8(d)-literal generators of C3 / C5) the CSR is
generated on the device, copied to the host once, preprocessed by the C oracle (oracle/, O(nnz), views without a second
copy) and by the GPU library; every scalar and every one of the 12 exported arrays must be bit-identical (array by array,
so only one exported array is resident at a time; its SHA-256 is recorded in gpurun_out/fullsize_digests.json when that
directory exists).  y is checked on ALL rows against an independent float64 evaluation of the CSR definition on the
device (torch index_add_, chunked), in permuted and in original order.  Fixture F2 (128^3 stencil) is
additionally compared with the unmodified reference build (oracle/_ref), host preprocessing and its closed-form scalars.
Size-independent properties stay: order_rid is a permutation, all-ones gives the row lengths, a row slab reproduces the
rows of the full product."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALARS = ["row_long", "row_block", "row_zero", "short_row_1", "short_row_3", "short_row_2", "short_row_4",
           "common_13", "short_row_34", "rowloop", "blocknum", "warp_number", "BlockNum_long", "fill0_nnz_long",
           "fill0_nnz_reg", "nnz_irreg", "origin_nnz_reg", "fill0_nnz_short", "fill0_nnz_short13",
           "fill0_nnz_short34", "fill0_nnz_short22", "threadblock13", "threadblock34", "threadblock22",
           "nnz_short", "nnz_long", "BlockNum", "BlockNum_short_1", "BlockNum_all", "sumBlockNum", "fill0_nnz_irreg"]


def _spec(name):
    from dasp_b200 import synth

    return {"f2_128": lambda: synth.stencil27(128), "c4": lambda: synth.stencil27(256),
            "c3": lambda: synth.powerlaw(), "c3_spec": lambda: synth.powerlaw_spec(),
            "c5": lambda: synth.skewed(), "c5_spec": lambda: synth.skewed_spec()}[name]()


def _host_gb_free():
    import psutil

    return psutil.virtual_memory().available / 2 ** 30


def reference_y(rp, ci, v, x, m):
    """y (original order, float64) from the CSR definition, evaluated by torch on the device in row chunks."""
    import torch

    y = torch.zeros(m, dtype=torch.float64, device=rp.device)
    xd = x.double()
    r0 = 0
    while r0 < m:  # row chunks of at most 2^27 entries (at least one row)
        target = rp[r0:r0 + 1].long() + (1 << 27)
        r1 = int(torch.searchsorted(rp, target.to(rp.dtype), right=True).item()) - 1
        r1 = min(m, max(r1, r0 + 1))
        k0, k1 = int(rp[r0].item()), int(rp[r1].item())
        if k1 > k0:
            lens = (rp[r0 + 1:r1 + 1] - rp[r0:r1]).long()
            rows = torch.repeat_interleave(torch.arange(r0, r1, device=rp.device), lens)
            y.index_add_(0, rows, v[k0:k1].double() * xd[ci[k0:k1].long()])
            del lens, rows
        r0 = r1
    return y


def _record(name, dtype, digests):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "fullsize_digests.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[f"{name}/{'f16' if dtype == oracle.F16 else 'f64'}"] = digests
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


CONFIGS = [("f2_128", oracle.F64, 2), ("f2_128", oracle.F16, 2), ("c4", oracle.F64, 16), ("c4", oracle.F16, 12),
           ("c3", oracle.F64, 10), ("c3_spec", oracle.F64, 10), ("c5", oracle.F64, 60), ("c5_spec", oracle.F64, 60)]


@pytest.mark.parametrize("name,dtype,need_gb", CONFIGS, ids=[f"{c[0]}-{'f16' if c[1] else 'f64'}" for c in CONFIGS])
def test_full_size_bit_exact_vs_oracle_and_all_rows(cuda_device, name, dtype, need_gb):
    import torch

    import dasp_b200
    from dasp_b200 import synth

    if _host_gb_free() < need_gb:
        pytest.skip(f"{name}: needs about {need_gb} GB of free host memory for the oracle, {_host_gb_free():.0f} available")
    half = dtype == oracle.F16
    spec = _spec(name)
    m, n = int(spec.m), int(spec.n)
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device, half=half)
    s = torch.cuda.current_stream().cuda_stream
    h = dasp_b200.Dasp(dtype, m, n, rp, ci, v, nnz=nnz)
    st = h.stats()
    assert st["nnz_long"] + st["nnz_short"] + st["origin_nnz_reg"] + st["nnz_irreg"] == nnz

    # ---- preprocessing, bit for bit, against the oracle on the same CSR ----
    rp_h, ci_h, v_h = rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy()
    ref = oracle.preprocess(dtype, m, n, rp_h, ci_h, v_h, copy=False)
    del ci_h, v_h
    for k in SCALARS:
        assert st[k] == ref[k], f"{name}: scalar {k}: {st[k]} != {ref[k]}"
    digests = {}
    for a in dasp_b200.lib.ARRAYS:
        got = h.export(a)
        want = ref[a]
        if half and a == "irreg_val":  # the reference leaves the pad element uninitialised (src/dasp_f16.h:1368-1369)
            got, want = got[:st["nnz_irreg"]], want[:st["nnz_irreg"]]
        assert got.shape == want.shape, f"{name}: {a} length {got.shape} != {want.shape}"
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), f"{name}: {a} differs from the oracle"
        digests[a] = hashlib.sha256(got.tobytes()).hexdigest() if got.nbytes <= (1 << 31) else f"{got.nbytes} bytes, equal"
        del got, want
    order_h = ref["order_rid"].copy()
    del ref
    _record(name, dtype, digests)

    # ---- y on ALL rows, permuted and original order, against an independent evaluation ----
    order = torch.from_numpy(order_h).to(cuda_device).long()
    assert int(torch.bincount(order, minlength=m).max().item()) == 1 and order.numel() == m
    gen = torch.Generator(device=cuda_device)
    gen.manual_seed(3)
    tdt = torch.float16 if half else torch.float64
    x = (torch.rand(n, generator=gen, device=cuda_device, dtype=torch.float64) * 2 - 1).to(tdt)
    y = torch.full((m,), float("nan"), dtype=tdt, device=cuda_device)
    yo = torch.full((m,), float("nan"), dtype=tdt, device=cuda_device)
    h.spmv(x, y, s)
    h.spmv_unpermuted(x, yo, s)
    torch.cuda.synchronize()
    y_ref = reference_y(rp, ci, v, x, m)
    den = float(torch.linalg.vector_norm(y_ref).item())
    tol = 2e-3 if half else 1e-12
    for got, want, what in ((y.double(), y_ref[order], "permuted"), (yo.double(), y_ref, "original order")):
        assert bool(torch.isfinite(got).all()), f"{name}: unwritten y entries ({what})"
        err = float(torch.linalg.vector_norm(got - want).item()) / den
        assert err <= tol, f"{name}: y {what}: relative L2 {err}"
        if half:
            assert float((got - want).abs().max().item()) <= 1.0  # the reference's own threshold, src/main_f16.cu:10
    # the chunked (deterministic) long-row path: permuted and original-order outputs are the same numbers
    h.set_variant(0, dasp_b200.VARIANT_CUDA_CORE, 0)
    h.spmv(x, y, s)
    h.spmv_unpermuted(x, yo, s)
    torch.cuda.synchronize()
    assert bool(torch.equal(yo[order], y))
    err = float(torch.linalg.vector_norm(yo.double() - y_ref).item()) / den
    assert err <= tol, f"{name}: chunked long rows: relative L2 {err}"
    h.close()


def test_f2_against_the_unmodified_reference(cuda_device):
    """Fixture F2 of SURVEY Appendix A (27-point stencil 128^3) through the reference's own host code (oracle/_ref) and its
    closed-form scalars; the reference also runs its kernels here, so its y is compared as well."""
    import torch

    import dasp_b200
    from dasp_b200 import synth

    if not oracle.ref_available(oracle.F64):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    spec = synth.stencil27(128)
    m = int(spec.m)
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device)
    h = dasp_b200.Dasp(oracle.F64, m, m, rp, ci, v, nnz=nnz)
    st = h.stats()
    assert (m, nnz) == (2_097_152, 55_742_968)
    assert (st["row_block"], st["rowloop"], st["blocknum"]) == (m, 4, 262_144)
    assert st["fill0_nnz_reg"] == 28 * 126 ** 3 + 16 * 6 * 126 ** 2 + 12 * 12 * 126 + 8 * 8 == 57_552_832
    assert (st["nnz_irreg"], st["origin_nnz_reg"]) == (190_512, 55_552_456)
    x = np.random.default_rng(5).uniform(-1, 1, m)
    r = oracle.ref_spmv_all(oracle.F64, m, m, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), x=x)
    for a in ("order_rid", "blockPtr", "irreg_rpt", "irreg_val", "irreg_cid", "reg_val", "reg_cid"):
        assert np.array_equal(h.export(a).view(np.uint8), r[a].view(np.uint8)), f"F2: {a} differs from the reference build"
    dx = torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    h.spmv(dx, dy, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if r["ran_on_gpu"]:
        ours, theirs = dy.cpu().numpy(), r["y_perm"]
        assert np.linalg.norm(ours - theirs) / np.linalg.norm(theirs) <= 1e-12
    h.close()


def test_all_ones_gives_row_lengths_at_full_size(cuda_device):
    """The reference's shipped run (A := 1, x := 1, src/main_f64.cu:131-132) on C3: y_perm == row lengths, exactly."""
    import torch

    import dasp_b200
    from dasp_b200 import synth

    spec = synth.powerlaw_spec()
    m = int(spec.m)
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device)
    v.fill_(1.0)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, m, rp, ci, v, nnz=nnz)
    order = torch.from_numpy(h.export("order_rid")).to(cuda_device).long()
    y = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    h.spmv(torch.ones(m, dtype=torch.float64, device=cuda_device), y, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.equal(y, (rp[1:] - rp[:-1]).double()[order]))
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
