"""CPU suite, part 1: pin the oracle.

The reference ships no golden vectors, so the pins are
  (a) tests/golden/golden_digests.json + golden_small.npz — produced by tests/golden/make_golden.py from
      the UNMODIFIED reference compiled from /root/reference (oracle/_ref), and
  (b) a live comparison with oracle/_ref when that library is present (build container),
  (c) the known-answer property of the reference's shipped run (all-ones => row lengths) and the
      structural identities of src/dasp_f64.h:1091.
"""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from cases import CASES, CPU_CASES, get, x_for

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ARRAYS = ["order_rid"] + oracle._REF_ARRAYS


def _defined(d):
    """The reference leaves the FP16 pad element of irreg_val uninitialised (src/dasp_f16.h:1368-1369);
    compare only the nnz_irreg defined entries."""
    d = dict(d)
    d["irreg_val"] = d["irreg_val"][: d["irreg_cid"].size]
    return d

CSV_COLS = ["short_row_1", "common_13", "short_row_3", "short_row_4", "short_row_2", "row_long", "row_block",
            "nnz_short", "fill0_nnz_short", "nnz_long", "fill0_nnz_long", "origin_nnz_reg", "fill0_nnz_reg", "nnz_irreg"]


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "golden_digests.json")) as f:
        return json.load(f)


def _val(v, dtype):
    return v.astype(np.float16 if dtype == oracle.F16 else np.float64)


@pytest.mark.parametrize("tag,dtype", [("f64", oracle.F64), ("f16", oracle.F16)])
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_golden_digests(golden, name, tag, dtype):
    m, n, rp, ci, v = get(name)
    o = _defined(oracle.preprocess(dtype, m, n, rp, ci, _val(v, dtype)))
    g = golden[f"{name}/{tag}"]
    for c in CSV_COLS:
        assert o[c] == g[c], f"{name}/{tag}: {c}"
    for a in ARRAYS:
        assert o[a].size == g["arrays"][a]["len"], f"{name}/{tag}: len({a})"
        assert _digest(o[a]) == g["arrays"][a]["sha256"], f"{name}/{tag}: {a}"


@pytest.mark.parametrize("tag,dtype", [("f64", oracle.F64), ("f16", oracle.F16)])
@pytest.mark.parametrize("name", ["pairs_128", "len_255_256"])
def test_oracle_matches_golden_arrays(name, tag, dtype):
    z = np.load(os.path.join(GOLD, "golden_small.npz"))
    m, n, rp, ci, v = get(name)
    o = _defined(oracle.preprocess(dtype, m, n, rp, ci, _val(v, dtype)))
    for a in ARRAYS:
        ref = z[f"{name}/{tag}/{a}"]
        assert np.array_equal(o[a].view(np.uint8), ref.view(np.uint8)), f"{name}/{tag}: {a}"


@pytest.mark.skipif(not (oracle.ref_available(oracle.F64) and oracle.ref_available(oracle.F16)),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("name", CPU_CASES)
def test_oracle_matches_compiled_reference(name, dtype):
    m, n, rp, ci, v = get(name)
    vv = _val(v, dtype)
    o = _defined(oracle.preprocess(dtype, m, n, rp, ci, vv))
    r = _defined(oracle.ref_spmv_all(dtype, m, n, rp, ci, vv))
    for a in ARRAYS:
        assert np.array_equal(o[a].view(np.uint8), r[a].view(np.uint8)), f"{name}: {a}"


@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("name", CPU_CASES)
def test_layout_semantics_reproduce_csr_product(name, dtype):
    """Evaluating y from the packed arrays (what the kernels do, K1-K7/K11) gives the CSR product through
    order_rid; also the structural identities of the layout."""
    m, n, rp, ci, v = get(name)
    vv = _val(v, dtype)
    o = oracle.preprocess(dtype, m, n, rp, ci, vv)
    assert sorted(o["order_rid"].tolist()) == list(range(m))
    assert o["nnz_long"] + o["nnz_short"] + o["origin_nnz_reg"] + o["nnz_irreg"] == int(rp[m])
    assert o["row_long"] + o["row_block"] + o["short_row_1"] + 2 * o["common_13"] + o["short_row_34"] \
        + o["short_row_2"] + o["row_zero"] == m
    x = x_for(n)
    if dtype == oracle.F64:
        y = oracle.csr_spmv_f64(m, rp, ci, vv, x)
        yl = oracle.layout_spmv(o, x)
    else:
        xh = x.astype(np.float16)
        y = oracle.csr_spmv_f16(m, rp, ci, vv, xh)
        yl = oracle.layout_spmv(o, xh)
    scale = max(np.linalg.norm(y), 1e-300)
    assert np.linalg.norm(yl - y[o["order_rid"]]) / scale <= 1e-14
    if o["row_zero"]:
        assert np.all(yl[m - o["row_zero"]:] == 0.0)


def test_all_ones_known_answer():
    """What the reference's main effectively checks (src/main_f64.cu:131-132): A := 1, x := 1."""
    m, n, rp, ci, v = get("mixed_f1")
    for dtype in (oracle.F64, oracle.F16):
        o = oracle.preprocess(dtype, m, n, rp, ci, np.ones_like(v))
        y = oracle.layout_spmv(o, np.ones(n))
        assert np.array_equal(y, np.diff(rp)[o["order_rid"]].astype(np.float64))


def test_f1_scalars_of_the_survey():
    """Fixture F1 scalar table of SURVEY.md Appendix A (obtained there from a patched reference copy)."""
    m, n, rp, ci, v = get("mixed_f1")
    want = {
        oracle.F64: dict(row_long=63, row_block=2600, row_zero=37, common_13=296, short_row_1=4, short_row_3=154,
                         short_row_4=333, short_row_2=217, rowloop=1, blocknum=328, warp_number=944, BlockNum_long=236,
                         fill0_nnz_long=60416, fill0_nnz_reg=338560, nnz_irreg=2614, origin_nnz_reg=338111,
                         fill0_nnz_short=3844, fill0_nnz_short13=1280, fill0_nnz_short34=2048, fill0_nnz_short22=512,
                         threadblock13=5, threadblock34=4, threadblock22=2, nnz_short=3416, nnz_long=58419),
        oracle.F16: dict(common_13=288, short_row_1=12, short_row_3=162, warp_number=264, BlockNum_long=66,
                         fill0_nnz_long=67584, fill0_nnz_reg=354304, nnz_irreg=2614, fill0_nnz_short=4108,
                         fill0_nnz_short13=1536, threadblock13=3, threadblock34=4, threadblock22=1),
    }
    for dtype, exp in want.items():
        o = oracle.preprocess(dtype, m, n, rp, ci, _val(v, dtype))
        for k, val in exp.items():
            assert o[k] == val, (dtype, k, o[k], val)


def test_stencil_closed_form_counts():
    """27-point stencil: every row is medium; padded size has the closed form of SURVEY.md Appendix A (F2)."""
    import matrices

    g = 20
    m, n, rp, ci, v = matrices.stencil27(g)
    o = oracle.preprocess(oracle.F64, m, n, rp, ci, v)
    i = g - 2
    assert o["row_block"] == m and o["row_long"] == 0 and o["fill0_nnz_short"] == 0
    assert int(rp[m]) == (3 * g - 2) ** 3
    assert np.bincount(np.diff(rp))[[8, 12, 18, 27]].tolist() == [8, 12 * i, 6 * i * i, i ** 3]


def test_half_conversion_matches_numpy():
    L = oracle.lib()
    L.dasp_oracle_half_to_double.restype = __import__("ctypes").c_double
    L.dasp_oracle_half_to_double.argtypes = [__import__("ctypes").c_uint16]
    L.dasp_oracle_double_to_half.restype = __import__("ctypes").c_uint16
    L.dasp_oracle_double_to_half.argtypes = [__import__("ctypes").c_double]
    bits = np.arange(0, 65536, 7, dtype=np.uint16)
    for b in bits:
        h = np.array([b], dtype=np.uint16).view(np.float16)[0]
        d = L.dasp_oracle_half_to_double(int(b))
        if np.isnan(h):
            assert np.isnan(d)
        else:
            assert d == float(h)
    rng = np.random.default_rng(0)
    for d in np.concatenate([rng.normal(0, 100, 2000), rng.normal(0, 1e-6, 500), [0.0, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 2.0 ** -25, 2.0 ** -24]]):
        got = L.dasp_oracle_double_to_half(float(d))
        want = int(np.array([d], dtype=np.float64).astype(np.float16).view(np.uint16)[0])
        assert got == want, (d, got, want)


def test_serial_csr_definition_and_threads_agree():
    m, n, rp, ci, v = get("powerlaw_20k")
    x = x_for(n)
    y = oracle.csr_spmv_f64(m, rp, ci, v, x)
    # literal definition: sequential, one accumulator per row
    for i in (0, 17, m // 2, m - 1):
        s = 0.0
        for j in range(rp[i], rp[i + 1]):
            s += v[j] * x[ci[j]]
        assert y[i] == s
    y_mt = oracle.csr_spmv_f64(m, rp, ci, v, x, threads=4)
    assert np.array_equal(y, y_mt)


@pytest.mark.skipif(not (oracle.ref_available(oracle.F64) and oracle.ref_available(oracle.F16)),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("threshold,block_longest", [(0.5, 256), (1.0, 256), (0.75, 64), (0.3, 1000), (0.9, 17)])
def test_oracle_matches_compiled_reference_with_other_parameters(threshold, block_longest, dtype):
    """The two run-time constants of the reference (src/main_f64.cu:124-125) away from their defaults."""
    for name in ("mixed_f1", "powerlaw_20k", "ragged_tail_blocks"):
        m, n, rp, ci, v = get(name)
        vv = _val(v, dtype)
        o = _defined(oracle.preprocess(dtype, m, n, rp, ci, vv, threshold, block_longest))
        r = _defined(oracle.ref_spmv_all(dtype, m, n, rp, ci, vv, threshold=threshold, block_longest=block_longest))
        for a in ARRAYS:
            assert np.array_equal(o[a].view(np.uint8), r[a].view(np.uint8)), f"{name}: {a}"


@pytest.mark.skipif(not (oracle.ref_available(oracle.F64) and oracle.ref_available(oracle.F16)),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(40))
def test_oracle_fuzz_against_compiled_reference(seed):
    """Random row-length mixes, duplicate / unsorted columns, random threshold and block_longest: restatement ==
    compiled reference, bit for bit, both precisions."""
    from fuzz import random_case

    m, n, rp, ci, v, threshold, block_longest = random_case(seed)
    for dtype in (oracle.F64, oracle.F16):
        vv = _val(v, dtype)
        o = _defined(oracle.preprocess(dtype, m, n, rp, ci, vv, threshold, block_longest))
        r = _defined(oracle.ref_spmv_all(dtype, m, n, rp, ci, vv, threshold=threshold, block_longest=block_longest))
        for a in ARRAYS:
            assert np.array_equal(o[a].view(np.uint8), r[a].view(np.uint8)), f"seed {seed} dtype {dtype}: {a}"
