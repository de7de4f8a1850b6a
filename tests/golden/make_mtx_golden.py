"""Writes the small Matrix Market fixtures under tests/golden/mtx/ and records, in mtx_golden.json, the CSR the
UNMODIFIED reference reader (mmio_allinone, src/mmio_highlevel.h:608, via oracle/_ref) produces for each of them in
both precisions.  Run in the build container:  make -C oracle ref && python tests/golden/make_mtx_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT]
import oracle  # noqa: E402

MTX = os.path.join(HERE, "mtx")
rng = np.random.default_rng(42)


def coo(m, n, k, sym=False):
    ent = set()
    while len(ent) < k:
        i, j = int(rng.integers(1, m + 1)), int(rng.integers(1, n + 1))
        if sym and j > i:
            i, j = j, i
        ent.add((i, j))
    ent = list(ent)
    rng.shuffle(ent)
    return ent


FILES = {}
e = coo(9, 7, 25)
FILES["general_real.mtx"] = ("%%MatrixMarket matrix coordinate real general\n% a comment\n%another\n9 7 27\n"
                             + "".join(f"{i} {j} {rng.uniform(-3, 3):.17g}\n" for i, j in e)
                             + f"{e[0][0]} {e[0][1]} 0.5\n{e[3][0]} {e[3][1]} -1e-3\n")  # duplicates are kept
e = coo(12, 12, 30, sym=True)
FILES["symmetric_real.mtx"] = ("%%MatrixMarket matrix coordinate real symmetric\n12 12 30\n"
                               + "".join(f"{i}  {j}\t{rng.uniform(-70000, 70000):.9e}\n" for i, j in e))
e = coo(10, 10, 22, sym=True)
FILES["pattern_symmetric.mtx"] = ("%%MatrixMarket matrix coordinate pattern symmetric\n%\n10 10 22\n"
                                  + "".join(f"{i} {j}\n" for i, j in e))
e = coo(6, 11, 18)
FILES["integer_general.mtx"] = ("%%MatrixMarket matrix coordinate integer general\n6 11 18\n"
                                + "".join(f"{i} {j} {int(rng.integers(-2000, 2000))}\n" for i, j in e))
e = coo(8, 8, 16, sym=True)
FILES["complex_hermitian.mtx"] = ("%%MatrixMarket matrix coordinate complex hermitian\n8 8 16\n"
                                  + "".join(f"{i} {j} {rng.uniform(-1, 1):.12g} {0.0 if i == j else rng.uniform(-1, 1):.12g}\n" for i, j in e))
e = [(i, j) for i, j in coo(9, 9, 14, sym=True) if i != j]
FILES["skew_symmetric.mtx"] = ("%%MatrixMarket matrix coordinate real skew-symmetric\n" + "9 9 %d\n" % len(e)
                               + "".join(f"{i} {j} {rng.uniform(-1, 1):.10g}\n" for i, j in e))
e = coo(5, 5, 9)
FILES["upper_case_crlf.mtx"] = ("%%MatrixMarket MATRIX Coordinate Real General\r\n%c\r\n5 5 9\r\n"
                                + "".join(f"{i} {j} {rng.uniform(-1, 1):.8f}\r\n" for i, j in e))
FILES["empty_rows.mtx"] = "%%MatrixMarket matrix coordinate real general\n6 6 3\n2 5 1.5\n2 1 -2.5\n6 6 4\n"


def main():
    os.makedirs(MTX, exist_ok=True)
    gold = {"_about": "CSR produced by the reference's mmio_allinone (oracle/_ref) for tests/golden/mtx/*.mtx"}
    for name, text in FILES.items():
        path = os.path.join(MTX, name)
        with open(path, "w", newline="") as f:
            f.write(text)
        for dtype, tag in ((oracle.F64, "f64"), (oracle.F16, "f16")):
            rc, r = oracle.ref_read_mtx(dtype, path)
            assert rc == 0, (name, rc)
            m, n, rp, ci, v, sym = r
            gold[f"{name}/{tag}"] = {"m": m, "n": n, "is_symmetric": sym, "rowptr": rp.tolist(), "colidx": ci.tolist(),
                                     "val_bits": (v.view(np.uint16) if dtype == oracle.F16 else v.view(np.uint64)).tolist()}
            print(name, tag, m, n, len(ci), sym)
    with open(os.path.join(HERE, "mtx_golden.json"), "w") as f:
        json.dump(gold, f, sort_keys=True)


if __name__ == "__main__":
    main()
