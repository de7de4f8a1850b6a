"""CPU suite, part 2: the C-ABI boundary without a GPU — the library loads, exports every symbol
include/dasp.h declares, its host-only logic works, and compute entries fail loudly (no fallback)."""
import ctypes as C

import numpy as np
import pytest

import dasp_b200
from cases import get


def test_library_loads_and_exports_every_declared_symbol():
    dasp_b200.load()
    declared = set(dasp_b200.lib.declared_symbols())
    exported = set(dasp_b200.exported_symbols())
    assert declared, "no declarations parsed from include/dasp.h"
    assert declared <= exported, f"missing exports: {sorted(declared - exported)}"
    # no torch types / C++ names on the boundary: every dasp_* export is an unmangled C symbol
    assert not [s for s in exported if s.startswith("_Z") and "dasp_" in s and "dasp4" not in s and "N4dasp" not in s]


def test_synth_library_loads():
    from dasp_b200 import synth

    L = synth.load()
    for name in ("dasp_synth_rowlen", "dasp_synth_fill", "dasp_synth_to_half", "dasp_synth_flush_l2"):
        assert hasattr(L, name)


def test_strerror_table():
    L = dasp_b200.load()
    assert L.dasp_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert L.dasp_strerror(code) not in (b"ok", b"unknown status")
    assert L.dasp_strerror(-99) == b"unknown status"


def test_partition_rows_is_nnz_balanced():
    m, n, rp, ci, v = get("powerlaw_20k")
    for parts in (1, 2, 3, 8):
        cuts = dasp_b200.partition_rows(rp, parts)
        assert cuts[0] == 0 and cuts[-1] == m and np.all(np.diff(cuts) >= 0)
        nnz = int(rp[m])
        for p in range(1, parts):
            # cut p is the smallest i with rowptr[i] >= p*nnz/parts
            t = nnz * p // parts
            i = cuts[p]
            assert rp[i] >= t and (i == 0 or rp[i - 1] < t)


def test_partition_rows_never_splits_a_row_and_handles_degenerate_input():
    rp = np.array([0, 0, 0, 1000000, 1000000, 1000003], dtype=np.int32)
    cuts = dasp_b200.partition_rows(rp, 4)
    assert cuts.tolist()[0] == 0 and cuts.tolist()[-1] == 5
    assert np.all(np.diff(cuts) >= 0)
    empty = np.zeros(1, dtype=np.int32)
    assert dasp_b200.partition_rows(empty, 3).tolist() == [0, 0, 0, 0]
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.partition_rows(rp, 0)


def test_invalid_arguments_are_rejected_before_touching_cuda():
    L = dasp_b200.load()
    h = C.c_void_p(None)
    rp = np.zeros(2, dtype=np.int32)
    assert L.dasp_create(None, 0, 0, 1, 1, 0, rp.ctypes.data, None, None, 0.75, 256) == -1
    assert L.dasp_create(C.byref(h), 0, 0, -1, 1, 0, rp.ctypes.data, None, None, 0.75, 256) == -1
    assert L.dasp_create(C.byref(h), 0, 0, 1, 1, 0, None, None, None, 0.75, 256) == -1
    assert L.dasp_create(C.byref(h), 0, 0, 1, 1, 0, rp.ctypes.data, None, None, 0.0, 256) == -1
    assert L.dasp_create(C.byref(h), 5, 0, 1, 1, 0, rp.ctypes.data, None, None, 0.75, 256) == -1
    assert L.dasp_create(C.byref(h), 0, 0, 1, 1, 2 ** 31, rp.ctypes.data, rp.ctypes.data, rp.ctypes.data, 0.75, 256) == -4
    assert h.value is None
    assert L.dasp_spmv(None, None, None, None) == -1
    assert L.dasp_destroy(None) == 0
    assert b"dasp_spmv" in L.dasp_last_error()


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    m, n, rp, ci, v = get("only_5")
    with pytest.raises(dasp_b200.DaspError) as e:
        dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    assert "CUDA" in str(e.value)
    with pytest.raises(dasp_b200.DaspError):
        dasp_b200.spmv_all(dasp_b200.DASP_F64, v, rp, ci, np.ones(n), m, n)


def test_product_never_imports_the_oracle():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "dasp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle|#include\s+\"[^\"]*oracle", text, re.M), f
