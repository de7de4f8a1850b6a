"""CPU suite: dasp_read_mtx (host-only entry of the C ABI) against the CSR the reference's own reader produced
for the committed fixtures (tests/golden/mtx_golden.json, generated from oracle/_ref), and live against oracle/_ref
when it is present.  Also the full file -> layout chain against the oracle."""
import json
import os

import numpy as np
import pytest

import dasp_b200
import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(f for f in os.listdir(os.path.join(GOLD, "mtx")) if f.endswith(".mtx"))


def _bits(v):
    return v.view(np.uint16) if v.dtype == np.float16 else v.view(np.uint64)


@pytest.mark.parametrize("tag,dtype", [("f64", dasp_b200.DASP_F64), ("f16", dasp_b200.DASP_F16)])
@pytest.mark.parametrize("name", FILES)
def test_reader_matches_reference_golden(name, tag, dtype):
    gold = json.load(open(os.path.join(GOLD, "mtx_golden.json")))[f"{name}/{tag}"]
    m, n, rp, ci, v, sym = dasp_b200.read_mtx(os.path.join(GOLD, "mtx", name), dtype)
    assert (m, n, sym) == (gold["m"], gold["n"], gold["is_symmetric"])
    assert rp.tolist() == gold["rowptr"] and ci.tolist() == gold["colidx"]
    assert _bits(v).tolist() == gold["val_bits"]


@pytest.mark.skipif(not oracle.ref_available(oracle.F64), reason="oracle/_ref not built")
@pytest.mark.parametrize("dtype", [oracle.F64, oracle.F16], ids=["f64", "f16"])
@pytest.mark.parametrize("name", FILES)
def test_reader_matches_compiled_reference(name, dtype):
    path = os.path.join(GOLD, "mtx", name)
    rc, ref = oracle.ref_read_mtx(dtype, path)
    assert rc == 0
    got = dasp_b200.read_mtx(path, dtype)
    assert got[:2] == ref[:2] and got[5] == ref[5]
    assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
    assert np.array_equal(_bits(got[4]), _bits(ref[4]))


def test_reader_semantics_spelled_out(tmp_path):
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n3 3 3\n3 1 2.5\n2 2 7\n2 1 -1\n")
    m, n, rp, ci, v, sym = dasp_b200.read_mtx(str(p))
    assert sym and (m, n) == (3, 3)
    # row 0 receives the mirrors of (3,1) then (2,1) in file order; row 1: (2,2) then (2,1); row 2: (3,1)
    assert rp.tolist() == [0, 2, 4, 5]
    assert ci.tolist() == [2, 1, 1, 0, 0] and v.tolist() == [2.5, -1.0, 7.0, -1.0, 2.5]
    p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 4\n")
    m, n, rp, ci, v, sym = dasp_b200.read_mtx(str(p))
    assert not sym and ci.tolist() == [0] and rp.tolist() == [0, 0, 1]  # not expanded (src/mmio_highlevel.h:642)


def test_reader_errors(tmp_path):
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.read_mtx(str(tmp_path / "missing.mtx"))
    p = tmp_path / "bad.mtx"
    p.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.read_mtx(str(p))
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n3 1 2.0\n")
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.read_mtx(str(p))
    p.write_text("not a banner\n")
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.read_mtx(str(p))


def test_file_to_layout_chain_against_oracle(tmp_path):
    """A file read by dasp_read_mtx gives the CSR the oracle preprocesses identically to one read by the
    reference reader (identical CSR => identical layout)."""
    from cases import get

    m, n, rp, ci, v = get("mixed_f1")
    p = tmp_path / "f1.mtx"
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (m, n, int(rp[m])))
        for i in range(m):
            for j in range(rp[i], rp[i + 1]):
                f.write("%d %d %.17g\n" % (i + 1, ci[j] + 1, v[j]))
    m2, n2, rp2, ci2, v2, _ = dasp_b200.read_mtx(str(p))
    assert (m2, n2) == (m, n) and np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(v2, v)
