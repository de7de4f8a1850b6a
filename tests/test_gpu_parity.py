"""GPU parity tests: the C-ABI library against the oracle, on the same seeded inputs.

* every preprocessing output bit-exact against oracle.preprocess (the restatement pinned to the
  compiled reference, see tests/test_oracle.py);
* FP64 y within 1e-12 relative L2 of the serial CSR result, through order_rid (north_star);
* FP16 y within the reference's own threshold |dy| <= 1 (src/main_f16.cu:10) AND a relative-L2
  bound so the gate is not vacuous (inputs rounded to half, reference product in double).
"""
import numpy as np
import pytest

import oracle
from cases import CASES, get, x_for

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-12          # north_star: relative L2 vs the serial CSR result
FP16_ABS_TOL = 1.0        # the reference's threshold, src/main_f16.cu:10
FP16_REL_TOL = 2e-3       # fp32 accumulation + one rounding to half (2^-11 ~ 4.9e-4 per element)

SCALARS = ["row_long", "row_block", "row_zero", "short_row_1", "short_row_3", "short_row_2", "short_row_4",
           "common_13", "short_row_34", "rowloop", "blocknum", "warp_number", "BlockNum_long", "fill0_nnz_long",
           "fill0_nnz_reg", "nnz_irreg", "origin_nnz_reg", "fill0_nnz_short", "fill0_nnz_short13",
           "fill0_nnz_short34", "fill0_nnz_short22", "threadblock13", "threadblock34", "threadblock22",
           "nnz_short", "nnz_long", "BlockNum", "BlockNum_short_1", "BlockNum_all", "sumBlockNum", "fill0_nnz_irreg"]


def _rel_l2(a, b):
    d = np.linalg.norm(a.astype(np.float64) - b.astype(np.float64))
    nb = np.linalg.norm(b.astype(np.float64))
    return d / nb if nb > 0 else d


@pytest.fixture(scope="module")
def dasp(cuda_device):
    import dasp_b200

    dasp_b200.load()
    return dasp_b200


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("name", list(CASES))
def test_preprocessing_bit_exact_and_spmv(dasp, cuda_device, name, dtype):
    import torch

    m, n, rp, ci, v = get(name)
    npdt = np.float16 if dtype == oracle.F16 else np.float64
    v = v.astype(npdt)
    ref = oracle.preprocess(dtype, m, n, rp, ci, v)
    h = dasp.Dasp(dtype, m, n, rp, ci, v)
    try:
        st = h.stats()
        for s in SCALARS:
            assert st[s] == ref[s], f"{name}: scalar {s}: {st[s]} != {ref[s]}"
        for a in dasp.lib.ARRAYS:
            got = h.export(a)
            assert got.shape == ref[a].shape, f"{name}: {a} length {got.shape} != {ref[a].shape}"
            assert np.array_equal(got.view(np.uint8), ref[a].view(np.uint8)), f"{name}: {a} differs"

        x = x_for(n).astype(npdt)
        if dtype == oracle.F64:
            y_ref = oracle.csr_spmv_f64(m, rp, ci, v, x)
        else:
            y_ref = oracle.csr_spmv_f16(m, rp, ci, v, x)
        order = ref["order_rid"]
        tdt = torch.float16 if dtype == oracle.F16 else torch.float64
        dx = torch.from_numpy(x).to(cuda_device)
        C_, M_, S_, T_, B_, N_ = (dasp.VARIANT_CUDA_CORE, dasp.VARIANT_MMA, dasp.VARIANT_SPLIT, dasp.VARIANT_TMA, dasp.VARIANT_BLOCKED,
                                  dasp.VARIANT_BANDED)
        # (medium, long, short): medium cuda/mma/split/banded; long cuda/mma/tma/blocked; short cuda/mma(FP64 only)/banded.
        # FP16 MMA = HMMA m16n8k16 (medium and long rows); the last triple is also used for the original-order product below
        triples = [(C_, C_, C_), (M_, M_, M_ if dtype == oracle.F64 else C_), (S_, C_, C_), (C_, T_, C_), (N_, C_, N_), (N_, B_, N_)]
        for variant in triples:
            h.set_variant(*variant)
            for rep in range(2):  # second call checks the self-resetting long-row counters / zero rows
                dy = torch.full((max(m, 1),), float("nan"), dtype=tdt, device=cuda_device)
                h.spmv(dx, dy, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                y = dy.cpu().numpy()[:m]
                assert np.all(np.isfinite(y.astype(np.float64))), f"{name}: unwritten y entries (variant {variant})"
                if dtype == oracle.F64:
                    assert _rel_l2(y, y_ref[order]) <= FP64_TOL, f"{name} variant {variant}"
                else:
                    assert np.max(np.abs(y.astype(np.float64) - y_ref[order]), initial=0.0) <= FP16_ABS_TOL
                    assert _rel_l2(y, y_ref[order]) <= FP16_REL_TOL, f"{name}"
            # the same product with the kernels reading reg_cid instead of the compact 16-bit indices: bit-equal
            if variant == (C_, C_, C_):
                h.set_index_compression(False)
                dy2 = torch.full((max(m, 1),), float("nan"), dtype=tdt, device=cuda_device)
                h.spmv(dx, dy2, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                h.set_index_compression(True)
                assert bool(torch.equal(dy2, dy)), f"{name}: compact indices change the result"
            # original-order output (the last variant of the loop is still selected: blocked long rows, banded short rows)
            dy = torch.full((max(m, 1),), float("nan"), dtype=tdt, device=cuda_device)
            h.spmv_unpermuted(dx, dy, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            y = dy.cpu().numpy()[:m]
            tol = FP64_TOL if dtype == oracle.F64 else FP16_REL_TOL
            assert _rel_l2(y, y_ref) <= tol, f"{name}: unpermuted"
    finally:
        h.close()


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
def test_all_ones_gives_row_lengths(dasp, cuda_device, dtype):
    """The reference's shipped run sets A and x to 1 (src/main_f64.cu:131-132): y_perm[k] must equal the
    length of row order_rid[k]; exact in half up to 2048."""
    m, n, rp, ci, v = get("mixed_f1")
    npdt = np.float16 if dtype == oracle.F16 else np.float64
    ones = np.ones(len(v), dtype=npdt)
    y, order = dasp.spmv_all(dtype, ones, rp, ci, np.ones(n, dtype=npdt), m, n)
    lens = np.diff(rp)[order]
    assert np.array_equal(y.astype(np.float64), lens.astype(np.float64))
    assert sorted(order.tolist()) == list(range(m))


def test_spmv_host_matches_device(dasp, cuda_device):
    m, n, rp, ci, v = get("powerlaw_20k")
    x = x_for(n)
    h = dasp.Dasp(oracle.F64, m, n, rp, ci, v)
    y = h.spmv_host(x)
    ref = oracle.csr_spmv_f64(m, rp, ci, v, x)
    assert _rel_l2(y, ref[h.export("order_rid")]) <= FP64_TOL
    h.close()


def test_device_csr_input(dasp, cuda_device):
    import torch

    m, n, rp, ci, v = get("mixed_f1")
    d = [torch.from_numpy(a).to(cuda_device) for a in (rp, ci, v)]
    h = dasp.Dasp(oracle.F64, m, n, d[0], d[1], d[2], nnz=int(rp[m]))
    ref = oracle.preprocess(oracle.F64, m, n, rp, ci, v)
    for a in dasp.lib.ARRAYS:
        assert np.array_equal(h.export(a).view(np.uint8), ref[a].view(np.uint8)), a
    h.close()


def test_malformed_csr_is_rejected(dasp, cuda_device):
    """dasp_create validates the CSR on the device: rowptr[0] != 0, decreasing rowptr, rowptr[m] != nnz and columns outside
    [0, n) all fail with a status instead of indexing out of bounds."""
    m, n, rp, ci, v = get("mixed_f1")
    nnz = int(rp[m])
    bad = []
    r = rp.copy(); r[0] = 1; bad.append((r, ci, nnz))
    r = rp.copy(); r[5], r[6] = r[6] + 3, r[5]; bad.append((r, ci, nnz))
    bad.append((rp, ci, nnz - 1))
    c = ci.copy(); c[nnz // 2] = n; bad.append((rp, c, nnz))
    c = ci.copy(); c[7] = -1; bad.append((rp, c, nnz))
    for r, c, k in bad:
        with pytest.raises(dasp.DaspError, match="malformed CSR"):
            dasp.Dasp(oracle.F64, m, n, r, c, v, nnz=k)
    h = dasp.Dasp(oracle.F64, m, n, rp, ci, v)  # and the library is still usable afterwards
    st = h.stats()
    assert (st["col_min"], st["col_max"]) == (int(ci.min()), int(ci.max()))
    h.close()


def test_host_path_uploads_only_the_column_range(dasp, cuda_device):
    """A row slab whose columns lie in [col_min, col_max]: dasp_spmv_host must not read x outside that range (the rest of
    the host vector is poisoned with NaN) and still give the CSR result."""
    m, n, rp, ci, v = get("stencil27_12")
    r0, r1 = m // 3, 2 * m // 3
    rps = (rp[r0:r1 + 1] - rp[r0]).astype(np.int32)
    cis, vs = ci[rp[r0]:rp[r1]], v[rp[r0]:rp[r1]]
    h = dasp.Dasp(oracle.F64, r1 - r0, n, rps, cis, vs)
    st = h.stats()
    assert st["col_min"] == int(cis.min()) and st["col_max"] == int(cis.max()) and st["col_max"] - st["col_min"] + 1 < n
    x = x_for(n)
    xp = np.full(n, np.nan)
    xp[st["col_min"]:st["col_max"] + 1] = x[st["col_min"]:st["col_max"] + 1]
    y = h.spmv_host(xp)
    ref = oracle.csr_spmv_f64(r1 - r0, rps, cis, vs, x)
    assert _rel_l2(y, ref[h.export("order_rid")]) <= FP64_TOL
    h.close()


def test_corrupted_checkpoint_is_rejected(dasp, cuda_device, tmp_path):
    """dasp_load checks the file against more than itself: format version, consistent scalars, monotone offsets, indices in
    range, no trailing bytes."""
    import struct

    m, n, rp, ci, v = get("mixed_f1")
    h = dasp.Dasp(oracle.F64, m, n, rp, ci, v)
    path = str(tmp_path / "a.dasp")
    h.save(path)
    st = h.stats()
    h.close()
    good = open(path, "rb").read()

    def load(data):
        q = str(tmp_path / "b.dasp")
        open(q, "wb").write(data)
        return dasp.Dasp.load_file(q)

    load(good).close()
    with pytest.raises(dasp.DaspError):  # trailing bytes
        load(good + b"\0")
    with pytest.raises(dasp.DaspError):  # truncated
        load(good[:-16])
    with pytest.raises(dasp.DaspError):  # wrong version
        load(good[:8 + 16] + struct.pack("<i", 99) + good[8 + 20:])
    # first array of the file is order_rid: make it a non-permutation
    head = 8 + 24 + 8 + dasp.lib.stats_struct_size()
    assert struct.unpack_from("<q", good, head)[0] == 4 * m
    data = bytearray(good)
    struct.pack_into("<i", data, head + 8, struct.unpack_from("<i", good, head + 12)[0])
    with pytest.raises(dasp.DaspError):
        load(bytes(data))
    # a column index out of range in long_cid (4th array)
    off = head
    for k in range(3):
        off += 8 + struct.unpack_from("<q", good, off)[0]
    assert struct.unpack_from("<q", good, off)[0] == 4 * st["fill0_nnz_long"]
    data = bytearray(good)
    struct.pack_into("<i", data, off + 8, n + 5)
    with pytest.raises(dasp.DaspError):
        load(bytes(data))


def test_two_handles_keep_their_devices(dasp, cuda_device):
    """Every entry selects the handle's device itself and restores the caller's; with two GPUs visible a handle on device 1
    is created and used while device 0 stays current (skipped on a single-GPU box, where only the restore is checked)."""
    import torch

    m, n, rp, ci, v = get("mixed_f1")
    x = x_for(n)
    ref = oracle.csr_spmv_f64(m, rp, ci, v, x)
    ndev = torch.cuda.device_count()
    torch.cuda.set_device(0)
    handles = [dasp.Dasp(oracle.F64, m, n, rp, ci, v, device=d) for d in range(min(ndev, 2))]
    assert torch.cuda.current_device() == 0
    for d, h in enumerate(handles):
        dx = torch.from_numpy(x).to(f"cuda:{d}")
        dy = torch.zeros(m, dtype=torch.float64, device=f"cuda:{d}")
        h.spmv(dx, dy, 0)  # default stream of the handle's device, while device 0 is current
        torch.cuda.synchronize(d)
        assert torch.cuda.current_device() == 0
        assert _rel_l2(dy.cpu().numpy(), ref[h.export("order_rid")]) <= FP64_TOL
    for h in handles:
        h.close()


def test_bad_arguments_fail_loudly(dasp, cuda_device):
    m, n, rp, ci, v = get("only_5")
    with pytest.raises(dasp.DaspError):
        dasp.Dasp(oracle.F64, m, n, rp, ci, v, threshold=0.0)
    with pytest.raises(dasp.DaspError):
        dasp.Dasp(7, m, n, rp, ci, v)
    h = dasp.Dasp(oracle.F64, m, n, rp, ci, v)
    with pytest.raises(dasp.DaspError):
        h.export("no_such_array")
    h.close()


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("name", ["mixed_f1", "powerlaw_20k", "stencil27_12"])
def test_report_matches_reference_csv_record(dasp, cuda_device, name, dtype):
    """dasp_report reproduces the structure columns of the record the reference writes (18 structure columns,
    rate_fill0, block_longest, data_X), compared textually with the reference's own output (oracle/_ref)."""
    if not oracle.ref_available(dtype):
        pytest.skip("oracle/_ref not built")
    m, n, rp, ci, v = get(name)
    v = v.astype(np.float16 if dtype == oracle.F16 else np.float64)
    ref = oracle.ref_spmv_all(dtype, m, n, rp, ci, v)["csv"].split(",")
    h = dasp.Dasp(dtype, m, n, rp, ci, v)
    got = h.report("ref_wrap", 0.5).split(",")
    h.close()
    assert len(got) == len(ref)
    assert got[:21] == ref[:21], (got[:21], ref[:21])


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
def test_axpby_and_checkpoint_roundtrip(dasp, cuda_device, dtype, tmp_path):
    """y = alpha*A*x + beta*y in both output orders, and a handle rebuilt from its checkpoint file holds the same
    arrays bit for bit and computes the same y bit for bit."""
    import torch

    m, n, rp, ci, v = get("mixed_f1")
    npdt = np.float16 if dtype == oracle.F16 else np.float64
    tdt = torch.float16 if dtype == oracle.F16 else torch.float64
    v = v.astype(npdt)
    x = x_for(n).astype(npdt)
    y0 = x_for(m, seed=11).astype(npdt)
    f = oracle.csr_spmv_f16 if dtype == oracle.F16 else oracle.csr_spmv_f64
    ax = f(m, rp, ci, v, x)
    h = dasp.Dasp(dtype, m, n, rp, ci, v)
    h.set_variant(0, dasp.VARIANT_CUDA_CORE, 0)  # bit equality below: the chunked long rows merge deterministically
    order = h.export("order_rid")
    s = torch.cuda.current_stream().cuda_stream
    dx = torch.from_numpy(x).to(cuda_device)
    tol = 1e-12 if dtype == oracle.F64 else 3e-3
    for alpha, beta in ((1.0, 0.0), (2.5, -0.75), (0.0, 1.0), (-1.0, 0.0)):
        for permuted in (True, False):
            dy = torch.from_numpy(y0.copy()).to(cuda_device)
            h.spmv_axpby(alpha, dx, beta, dy, permuted, s)
            torch.cuda.synchronize()
            got = dy.cpu().numpy().astype(np.float64)
            want = alpha * (ax[order] if permuted else ax) + beta * y0.astype(np.float64)
            assert np.linalg.norm(got - want) <= tol * max(np.linalg.norm(want), 1.0), (alpha, beta, permuted)
    path = str(tmp_path / "layout.dasp")
    h.save(path)
    h2 = dasp.Dasp.load_file(path)
    h2.set_variant(0, dasp.VARIANT_CUDA_CORE, 0)
    assert (h2.m, h2.n, h2.nnz, h2.dtype) == (m, n, int(rp[m]), dtype)
    for a in dasp.lib.ARRAYS:
        assert np.array_equal(h.export(a).view(np.uint8), h2.export(a).view(np.uint8)), a
    st1, st2 = h.stats(), h2.stats()
    assert {k: st1[k] for k in st1 if k not in ("device_bytes",)} == {k: st2[k] for k in st2 if k not in ("device_bytes",)}
    ya = torch.zeros(m, dtype=tdt, device=cuda_device)
    yb = torch.zeros(m, dtype=tdt, device=cuda_device)
    for rep in range(2):
        h.spmv(dx, ya, s)
        h2.spmv(dx, yb, s)
        torch.cuda.synchronize()
        assert bool(torch.equal(ya, yb))
    h.close()
    h2.close()
    with pytest.raises(dasp.DaspError):
        dasp.Dasp.load_file(str(tmp_path / "missing.dasp"))


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("threshold,block_longest", [(0.5, 256), (1.0, 256), (0.75, 64), (0.3, 1000), (0.9, 17)])
def test_non_default_threshold_and_block_longest(dasp, cuda_device, threshold, block_longest, dtype):
    """The reference's two run-time constants away from 0.75 / 256: layout bit-exact vs the oracle, y correct."""
    import torch

    for name in ("mixed_f1", "powerlaw_20k"):
        m, n, rp, ci, v = get(name)
        npdt = np.float16 if dtype == oracle.F16 else np.float64
        v = v.astype(npdt)
        ref = oracle.preprocess(dtype, m, n, rp, ci, v, threshold, block_longest)
        h = dasp.Dasp(dtype, m, n, rp, ci, v, threshold=threshold, block_longest=block_longest)
        for a in dasp.lib.ARRAYS:
            assert np.array_equal(h.export(a).view(np.uint8), ref[a].view(np.uint8)), f"{name}: {a}"
        x = x_for(n).astype(npdt)
        dx = torch.from_numpy(x).to(cuda_device)
        dy = torch.zeros(m, dtype=dx.dtype, device=cuda_device)
        h.spmv_unpermuted(dx, dy, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        f = oracle.csr_spmv_f16 if dtype == oracle.F16 else oracle.csr_spmv_f64
        y_ref = f(m, rp, ci, v, x)
        assert _rel_l2(dy.cpu().numpy(), y_ref) <= (FP64_TOL if dtype == oracle.F64 else FP16_REL_TOL)
        h.close()


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
def test_host_batch_equals_individual_products(dasp, cuda_device, dtype):
    """dasp_spmv_host_batch (pipelined copies/kernels, double-buffered staging) returns, for every right-hand side,
    exactly what dasp_spmv_host returns, for batches shorter and longer than the pipeline depth."""
    import torch

    m, n, rp, ci, v = get("powerlaw_20k")
    tdt = torch.float16 if dtype == oracle.F16 else torch.float64
    h = dasp.Dasp(dtype, m, n, rp, ci, v.astype(np.float16 if dtype == oracle.F16 else np.float64))
    h.set_variant(0, dasp.VARIANT_CUDA_CORE, 0)  # bit equality below
    for count in (1, 2, 5):
        xs = [torch.from_numpy(x_for(n, seed=100 + j)).to(tdt).pin_memory() for j in range(count)]
        ys = [torch.full((m,), float("nan"), dtype=tdt).pin_memory() for _ in range(count)]
        h.spmv_host_batch(xs, ys)
        for j in range(count):
            want = torch.empty(m, dtype=tdt)
            h.spmv_host(xs[j], want)
            assert bool(torch.equal(ys[j], want)), (count, j)
    h.spmv_host_batch([], [])
    h.close()


def test_spmv_is_cuda_graph_capturable(dasp, cuda_device):
    """dasp_spmv only enqueues one kernel on the caller's stream, so a solver can capture its iteration in a CUDA
    graph (launch-bound small matrices); replaying the graph gives the same y as direct launches."""
    import torch

    m, n, rp, ci, v = get("mixed_f1")
    h = dasp.Dasp(oracle.F64, m, n, rp, ci, v)
    h.set_variant(0, dasp.VARIANT_CUDA_CORE, 0)  # bit equality below
    dx = torch.from_numpy(x_for(n)).to(cuda_device)
    y_direct = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    y_graph = torch.zeros(m, dtype=torch.float64, device=cuda_device)
    h.spmv(dx, y_direct, torch.cuda.current_stream().cuda_stream)  # also sets the kernel attributes before capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(3):
                h.spmv(dx, y_graph, torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    assert bool(torch.equal(y_graph, y_direct))
    h.close()


def test_c_example_runs_against_the_abi(dasp, cuda_device, tmp_path):
    """examples/spmv_mtx.c (the reference's command line rewritten on the C ABI) builds with gcc against include/dasp.h,
    reads a Matrix Market file, runs and verifies itself against the serial CSR loop."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "spmv_mtx")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "spmv_mtx.c"),
                           "-L", os.path.join(root, "dasp_b200"), "-ldasp_b200", "-Wl,-rpath," + os.path.join(root, "dasp_b200"),
                           "-lm", "-o", exe])
    m, n, rp, ci, v = get("mixed_f1")
    mtx = tmp_path / "f1.mtx"
    with open(mtx, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (m, n, int(rp[m])))
        for i in range(m):
            for j in range(rp[i], rp[i + 1]):
                f.write("%d %d %.17g\n" % (i + 1, ci[j] + 1, v[j]))
    for extra in ([], ["-ones"]):
        p = subprocess.run([exe, str(mtx), "20"] + extra, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr
        assert "PASS" in p.stdout and "SpMV_X" in p.stdout
        rec = [l for l in p.stdout.splitlines() if l.startswith("record: ")][0][len("record: "):].split(",")
        assert rec[1:4] == [str(m), str(n), str(int(rp[m]))]


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_layout_and_product(dasp, cuda_device, seed):
    """Random row-length mixes, duplicate / unsorted columns, random threshold and block_longest: GPU preprocessing
    bit-exact vs the oracle and y within tolerance, FP64 and FP16."""
    import torch
    from fuzz import random_case

    m, n, rp, ci, v, threshold, block_longest = random_case(seed)
    for dtype in (oracle.F64, oracle.F16):
        npdt = np.float16 if dtype == oracle.F16 else np.float64
        vv = v.astype(npdt)
        ref = oracle.preprocess(dtype, m, n, rp, ci, vv, threshold, block_longest)
        h = dasp.Dasp(dtype, m, n, rp, ci, vv, threshold=threshold, block_longest=block_longest)
        for a in dasp.lib.ARRAYS:
            assert np.array_equal(h.export(a).view(np.uint8), ref[a].view(np.uint8)), f"seed {seed}: {a}"
        x = x_for(n, seed=seed).astype(npdt)
        dx = torch.from_numpy(x).to(cuda_device)
        dy = torch.full((m,), float("nan"), dtype=dx.dtype, device=cuda_device)
        h.spmv_unpermuted(dx, dy, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        f = oracle.csr_spmv_f16 if dtype == oracle.F16 else oracle.csr_spmv_f64
        y_ref = f(m, rp, ci, vv, x)
        got = dy.cpu().numpy().astype(np.float64)
        assert np.all(np.isfinite(got))
        scale = max(np.linalg.norm(y_ref), 1e-300)
        assert np.linalg.norm(got - y_ref) / scale <= (FP64_TOL if dtype == oracle.F64 else 4e-3), f"seed {seed}"
        h.close()


@pytest.mark.parametrize("prec", ["double", "half"])
def test_reference_main_linked_against_the_shim(cuda_device, tmp_path, prec):
    """The reference's OWN main program (src/main_f64.cu / src/main_f16.cu, unmodified: its Matrix Market reader, its
    cuSPARSE comparator, its spmv_all call site, its CSV append) built against libdasp_b200.so through
    include/dasp_reference_shim.h (oracle/Makefile `ref`, symlink shadow directory), with the verification it left
    commented out re-enabled: verify_new(cuSPARSE y, our y, our order_rid, rows) (src/main_f64.cu:3-16,157) must succeed."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", f"spmv_{prec}_shim")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/spmv_*_shim not built (needs /root/reference at build time)")
    os.makedirs(tmp_path / "data")  # the reference appends its record to data/*.csv relative to the working directory
    files = [os.path.join(root, "tests", "golden", "mtx", f) for f in ("general_real.mtx", "symmetric_real.mtx", "pattern_symmetric.mtx")]
    m, n, rp, ci, v = get("mixed_f1")  # every row category, long rows included
    big = tmp_path / "f1.mtx"
    with open(big, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (m, n, int(rp[m])))
        for i in range(m):
            for j in range(rp[i], rp[i + 1]):
                f.write("%d %d %.17g\n" % (i + 1, ci[j] + 1, v[j]))
    for path in files + [str(big)]:
        p = subprocess.run([exe, path], capture_output=True, text=True, timeout=300, cwd=tmp_path)
        assert p.returncode == 0, (path, p.stdout[-800:], p.stderr[-800:])
        assert "compute succeed" in p.stdout and "VERIFY_NEW rc=0" in p.stdout, p.stdout[-800:]
        assert "cusparse:" in p.stdout
    rec = open(tmp_path / "data" / ("spmv_f64_record.csv" if prec == "double" else "spmv_f16_record.csv")).read().strip().splitlines()
    assert len(rec) == len(files) + 1


@pytest.mark.parametrize("name", ["mixed_f1", "powerlaw_20k", "stencil27_12"])
def test_fp16_in_fp32_out(dasp, cuda_device, name):
    """FP16 matrix and x, FP32 y: the unrounded fp32 accumulator, permuted and original order; much closer to
    the double-precision product of the half inputs than the half-rounded y, and rounding it to half reproduces dasp_spmv."""
    import torch

    m, n, rp, ci, v = get(name)
    v = v.astype(np.float16)
    x = x_for(n).astype(np.float16)
    y_ref = oracle.csr_spmv_f16(m, rp, ci, v, x)
    h = dasp.Dasp(oracle.F16, m, n, rp, ci, v)
    h.set_variant(0, dasp.VARIANT_CUDA_CORE, 0)
    order = h.export("order_rid")
    s = torch.cuda.current_stream().cuda_stream
    dx = torch.from_numpy(x).to(cuda_device)
    y16 = torch.zeros(m, dtype=torch.float16, device=cuda_device)
    y32 = torch.full((m,), float("nan"), dtype=torch.float32, device=cuda_device)
    y32o = torch.full((m,), float("nan"), dtype=torch.float32, device=cuda_device)
    h.spmv(dx, y16, s)
    h.spmv_f32out(dx, y32, True, s)
    h.spmv_f32out(dx, y32o, False, s)
    torch.cuda.synchronize()
    assert bool(torch.equal(y32.to(torch.float16), y16))
    assert bool(torch.equal(y32o[torch.from_numpy(order).to(cuda_device).long()], y32))
    err32 = _rel_l2(y32.cpu().numpy(), y_ref[order])
    err16 = _rel_l2(y16.cpu().numpy(), y_ref[order])
    assert err32 <= 2e-6 and err32 < err16
    h.close()
    with pytest.raises(dasp.DaspError):
        h64 = dasp.Dasp(oracle.F64, m, n, rp, ci, v.astype(np.float64))
        try:
            h64.spmv_f32out(dx, y32, True, s)
        finally:
            h64.close()


_SMQ_SNIPPET = r"""
import sys
import numpy as np, torch
sys.path.insert(0, {tests!r}); sys.path.insert(0, {root!r})
import dasp_b200, oracle
from cases import get, x_for
dasp_b200.load()
dev = torch.device("cuda:0")
for name in ("mixed_f1", "powerlaw_20k", "ragged_tail_blocks", "wide_span_mixed", "stencil27_12", "rowloop_59990"):
    m, n, rp, ci, v = get(name)
    for dtype, npdt, tdt, tol in ((oracle.F64, np.float64, torch.float64, 1e-12), (oracle.F16, np.float16, torch.float16, 2e-3)):
        vv = v.astype(npdt); x = x_for(n).astype(npdt)
        h = dasp_b200.Dasp(dtype, m, n, rp, ci, vv)
        order = h.export("order_rid")
        dx = torch.from_numpy(x).to(dev); dy = torch.zeros(m, dtype=tdt, device=dev)
        want = (oracle.csr_spmv_f64 if dtype == oracle.F64 else oracle.csr_spmv_f16)(m, rp, ci, vv, x)[order]
        for rep in range(3):  # the queue counters reset themselves: the second and third product must be right too
            dy.fill_(7)
            h.spmv(dx, dy, torch.cuda.current_stream().cuda_stream); torch.cuda.synchronize()
            got = dy.cpu().numpy().astype(np.float64)
            err = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)
            assert err <= tol, (name, dtype, rep, err)
        h.close()
print("smq ok")
"""


@pytest.mark.parametrize("lean", ["0", "1"], ids=["pipelined", "lean"])
def test_sm_affine_queue_variant(cuda_device, lean):
    """smq_kernel (spmv.cu) is chosen by the environment only (measured slower than the fused kernel, never AUTO): run it
    in a child process on small matrices of every category mix, three products in a row per handle."""
    import os
    import subprocess
    import sys

    tests = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, DASP_SMQ="1", DASP_SMQ_LEAN=lean)
    out = subprocess.run([sys.executable, "-c", _SMQ_SNIPPET.format(tests=tests, root=os.path.dirname(tests))],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "smq ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
