"""GPU: the bench generators (libdasp_synth.so) produce what include/dasp_synth.h says, slabs are consistent
with the global matrix, and the product on generated matrices matches the oracle."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _host(spec, r0, r1, dev, half=False):
    from dasp_b200 import synth

    rp, ci, v, nnz = synth.generate(spec, r0, r1, dev, half=half)
    return rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), nnz


def test_stencil_matches_numpy_twin_structure(cuda_device):
    import matrices
    from dasp_b200 import synth

    spec = synth.stencil27(11, 7, 5)
    rp, ci, v, nnz = _host(spec, 0, spec.m, cuda_device)
    m, n, rp2, ci2, _ = matrices.stencil27(11, 7, 5)
    assert spec.m == m and np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    assert np.all(np.abs(v) <= 1.0) and len(np.unique(v)) > nnz // 2


@pytest.mark.parametrize("kind", ["powerlaw", "skewed", "banded"])
def test_slabs_concatenate_to_the_global_matrix(cuda_device, kind):
    from dasp_b200 import synth

    spec = {"powerlaw": lambda: synth.powerlaw(m=30000, lmax=20000),
            "skewed": lambda: synth.skewed(n_long=5, long_len=6000, n_short=20000),
            "banded": lambda: synth.banded(m=5000, mean_len=22, window=256)}[kind]()
    rp, ci, v, nnz = _host(spec, 0, spec.m, cuda_device)
    cut = int(spec.m) // 3
    a = _host(spec, 0, cut, cuda_device)
    b = _host(spec, cut, spec.m, cuda_device)
    assert a[3] + b[3] == nnz
    assert np.array_equal(np.concatenate([a[1], b[1]]), ci) and np.array_equal(np.concatenate([a[2], b[2]]), v)
    assert np.array_equal(b[0] + a[3], rp[cut:])
    assert ci.min() >= 0 and ci.max() < spec.n
    lens = np.diff(rp)
    if kind == "skewed":
        assert np.all(lens[:5] == 6000) and set(np.unique(lens[5:])) <= {1, 2, 3, 4}
        for r in range(5):  # long rows: ascending, distinct, inside the common band
            c = ci[rp[r]:rp[r + 1]]
            assert np.all(np.diff(c) > 0) and c[0] >= spec.band_lo and c[-1] < spec.band_lo + spec.band
    if kind == "banded":  # structurally and numerically symmetric, diagonal present
        import scipy.sparse as sp

        A = sp.csr_matrix((v, ci, rp), shape=(spec.m, spec.n))
        assert (abs(A - A.T)).max() == 0.0 and np.all(A.diagonal() != 0.0)
        assert 15 < lens.mean() < 30
    if kind == "powerlaw":
        assert lens.max() > 1000 and np.median(lens) <= 2 and (lens == 1).mean() > 0.3


@pytest.mark.parametrize("kind,half", [("powerlaw", False), ("skewed", False), ("banded", True)])
def test_product_on_generated_matrices(cuda_device, kind, half):
    import torch

    import dasp_b200
    from dasp_b200 import synth

    spec = {"powerlaw": lambda: synth.powerlaw(m=40000, lmax=20000),
            "skewed": lambda: synth.skewed(n_long=9, long_len=9000, n_short=30000),
            "banded": lambda: synth.banded(m=8000, mean_len=22, window=512)}[kind]()
    m, n = int(spec.m), int(spec.n)
    rp, ci, v, nnz = synth.generate(spec, 0, m, cuda_device, half=half)
    dtype = dasp_b200.DASP_F16 if half else dasp_b200.DASP_F64
    h = dasp_b200.Dasp(dtype, m, n, rp, ci, v, nnz=nnz)
    ref = oracle.preprocess(dtype, m, n, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy())
    for a in dasp_b200.lib.ARRAYS:
        assert np.array_equal(h.export(a).view(np.uint8), ref[a].view(np.uint8)), a
    x = np.random.default_rng(1).uniform(-1, 1, n).astype(np.float16 if half else np.float64)
    dx = torch.from_numpy(x).to(cuda_device)
    dy = torch.zeros(m, dtype=dx.dtype, device=cuda_device)
    h.spmv(dx, dy, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    y = dy.cpu().numpy().astype(np.float64)
    f = oracle.csr_spmv_f16 if half else oracle.csr_spmv_f64
    y_ref = f(m, rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy(), x)[ref["order_rid"]]
    err = np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref)
    assert err <= (2e-3 if half else 1e-12)
    h.close()


@pytest.mark.parametrize("kind", ["powerlaw_spec", "skewed_spec"])
def test_spec_literal_generators(cuda_device, kind):
    """Kinds 4 / 5 (SURVEY.md §8(d) taken literally): slabs concatenate to the global matrix, columns are in range, DISTINCT
    within every row and NOT sorted; C3: 9 of 10 entries inside the row's window; C5: exactly n_long long rows at one seeded
    position per stride, their columns spread over all of n; and the product on them matches the oracle."""
    import torch

    import dasp_b200
    from dasp_b200 import synth

    spec = {"powerlaw_spec": lambda: synth.powerlaw_spec(m=60000, lmax=20000, window=512),
            "skewed_spec": lambda: synth.skewed_spec(n_long=7, long_len=9000, n_short=40000, window=512)}[kind]()
    m, n = int(spec.m), int(spec.n)
    rp, ci, v, nnz = _host(spec, 0, m, cuda_device)
    cut = m // 3
    a, b = _host(spec, 0, cut, cuda_device), _host(spec, cut, m, cuda_device)
    assert a[3] + b[3] == nnz and np.array_equal(np.concatenate([a[1], b[1]]), ci) and np.array_equal(np.concatenate([a[2], b[2]]), v)
    assert ci.min() >= 0 and ci.max() < n
    lens = np.diff(rp)
    rows = np.repeat(np.arange(m), lens)
    key = rows.astype(np.int64) * n + ci
    assert len(np.unique(key)) == nnz, "duplicate column inside a row"
    longest = int(np.argmax(lens))
    c = ci[rp[longest]:rp[longest + 1]]
    assert np.any(np.diff(c) < 0), "columns of the longest row are sorted"
    if kind == "skewed_spec":
        long_rows = np.flatnonzero(lens == 9000)
        stride = m // 7
        assert len(long_rows) == 7 and np.array_equal(long_rows // stride, np.arange(7))
        assert set(np.unique(lens[lens != 9000])) <= {1, 2, 3, 4}
        assert c.max() - c.min() > n // 2
        short = np.flatnonzero(lens <= 4)[:2000]
        for r in short[::97]:
            cc = ci[rp[r]:rp[r + 1]]
            assert np.all(np.abs(cc - r * n / m) <= 2 * 512)
    else:
        assert lens.max() > 1000 and np.median(lens) <= 2
        mid = (np.arange(m) > 5000) & (np.arange(m) < m - 5000)  # away from the border, where the window is clipped
        r = int(np.flatnonzero((lens >= 50) & (lens < 400) & mid)[0])
        cc = ci[rp[r]:rp[r + 1]]
        inside = np.abs(cc - r * n / m) <= 512 + 1
        assert 0.85 <= inside.mean() <= 0.95
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    ref = oracle.preprocess(oracle.F64, m, n, rp, ci, v)
    for arr in dasp_b200.lib.ARRAYS:
        assert np.array_equal(h.export(arr).view(np.uint8), ref[arr].view(np.uint8)), arr
    x = np.random.default_rng(1).uniform(-1, 1, n)
    dx = torch.from_numpy(x).to(cuda_device)
    y_ref = oracle.csr_spmv_f64(m, rp, ci, v, x)[ref["order_rid"]]
    for long_variant in (dasp_b200.VARIANT_CUDA_CORE, dasp_b200.VARIANT_BLOCKED):
        h.set_variant(0, long_variant, 0)
        for rep in range(2):
            dy = torch.full((m,), float("nan"), dtype=torch.float64, device=cuda_device)
            h.spmv(dx, dy, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            y = dy.cpu().numpy()
            assert np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref) <= 1e-12, (long_variant, rep)
    h.close()
