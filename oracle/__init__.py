"""oracle — CPU checker for the DASP hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``dasp_b200``) never does.

Two things live here:

* ``libdasp_oracle.so``  — plain-C restatement of the reference's host preprocessing and of the
  serial CSR SpMV (``dasp_oracle.c``; citations inside).
* ``_ref/libdasp_ref_f64.so`` / ``_ref/libdasp_ref_f16.so`` — the UNMODIFIED reference compiled
  from ``/root/reference/src`` by ``oracle/Makefile`` (``ref_wrap.cu`` only ``#include``s it and
  records what ``spmv_all`` uploads).  Built in the build container, git-ignored, shipped to the
  GPU box as a binary.  Used to pin the restatement bit-for-bit and, with a GPU, to run the
  reference's own kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
F64, F16 = 0, 1

_SCALARS = [
    "dtype", "m", "n", "nnz", "row_long", "row_block", "row_zero",
    "short_row_1", "short_row_3", "short_row_2", "short_row_4", "common_13", "short_row_34",
    "rowloop", "blocknum", "warp_number", "BlockNum_long", "fill0_nnz_long",
    "fill0_nnz_reg", "nnz_irreg", "origin_nnz_reg",
    "fill0_nnz_short", "fill0_nnz_short13", "fill0_nnz_short34", "fill0_nnz_short22",
    "threadblock13", "threadblock34", "threadblock22", "nnz_short", "nnz_long",
    "BlockNum", "BlockNum_short_1", "BlockNum_all", "sumBlockNum", "fill0_nnz_irreg",
]
_ARRAYS = ["order_rid", "long_rpt_new", "long_val", "long_cid", "blockPtr", "irreg_rpt",
           "irreg_val", "irreg_cid", "reg_val", "reg_cid", "short_val", "short_cid"]


class _Layout(C.Structure):
    _fields_ = [(s, C.c_int) for s in _SCALARS] + [(a, C.c_void_p) for a in _ARRAYS]


def build(ref: bool = False) -> None:
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "libdasp_oracle.so")
        src = os.path.join(_HERE, "dasp_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build()
        L = C.CDLL(so)
        L.dasp_oracle_preprocess.restype = C.c_int
        L.dasp_oracle_preprocess.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_double, C.c_int, C.POINTER(_Layout)]
        L.dasp_oracle_free.argtypes = [C.POINTER(_Layout)]
        L.dasp_oracle_csr_spmv_f64.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.dasp_oracle_csr_spmv_f64_mt.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int]
        L.dasp_oracle_csr_spmv_f16.argtypes = [C.c_int] + [C.c_void_p] * 5
        L.dasp_oracle_layout_spmv.argtypes = [C.POINTER(_Layout), C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _val_np(dtype: int):
    return np.float16 if dtype == F16 else np.float64


def _array_len(name: str, s: dict) -> int:
    return {
        "order_rid": s["m"], "long_rpt_new": s["row_long"] + 1,
        "long_val": s["fill0_nnz_long"], "long_cid": s["fill0_nnz_long"],
        "blockPtr": s["blocknum"] + 1, "irreg_rpt": s["row_block"] + 1,
        "irreg_val": s["fill0_nnz_irreg"], "irreg_cid": s["nnz_irreg"],
        "reg_val": s["fill0_nnz_reg"], "reg_cid": s["fill0_nnz_reg"],
        "short_val": s["fill0_nnz_short"], "short_cid": s["fill0_nnz_short"],
    }[name]


def _canon(dtype, rowptr, colidx, val):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=_val_np(dtype))
    return rowptr, colidx, val


class _Owned:
    """Keeps the C layout alive while views into its arrays exist (preprocess(copy=False))."""

    def __init__(self, lay):
        self.lay = lay

    def __del__(self):
        try:
            lib().dasp_oracle_free(C.byref(self.lay))
        except Exception:
            pass


def preprocess(dtype: int, m: int, n: int, rowptr, colidx, val, threshold: float = 0.75,
               block_longest: int = 256, copy: bool = True) -> dict:
    """Run the C restatement; returns {scalar: int, array: np.ndarray}.  copy=False returns views into the C buffers
    (full-size configurations: no second copy of a 14 GB layout); they stay valid as long as the dict lives."""
    rowptr, colidx, val = _canon(dtype, rowptr, colidx, val)
    nnz = int(rowptr[m])
    lay = _Layout()
    rc = lib().dasp_oracle_preprocess(dtype, m, n, nnz, _p(rowptr), _p(colidx), _p(val), threshold,
                                      block_longest, C.byref(lay))
    assert rc == 0
    out = {s: int(getattr(lay, s)) for s in _SCALARS}
    for a in _ARRAYS:
        cnt = _array_len(a, out)
        npdt = _val_np(dtype) if a.endswith("_val") else np.int32
        ptr = getattr(lay, a)
        if cnt == 0 or not ptr:
            out[a] = np.zeros(0, dtype=npdt)
        else:
            buf = (C.c_char * (cnt * np.dtype(npdt).itemsize)).from_address(ptr)
            view = np.frombuffer(buf, dtype=npdt)
            out[a] = view.copy() if copy else view
    if copy:
        lib().dasp_oracle_free(C.byref(lay))
    else:
        out["_owner"] = _Owned(lay)
    return out


def layout_spmv(layout: dict, x) -> np.ndarray:
    """y (permuted order, float64) evaluated from the packed arrays of ``preprocess``."""
    lay = _Layout()
    keep = []
    for s in _SCALARS:
        setattr(lay, s, layout[s])
    for a in _ARRAYS:
        arr = np.ascontiguousarray(layout[a])
        keep.append(arr)
        setattr(lay, a, arr.ctypes.data if arr.size else None)
    x = np.ascontiguousarray(x, dtype=_val_np(layout["dtype"]))
    y = np.zeros(layout["m"], dtype=np.float64)
    lib().dasp_oracle_layout_spmv(C.byref(lay), _p(x), _p(y))
    return y


def csr_spmv_f64(m, rowptr, colidx, val, x, threads: int = 1) -> np.ndarray:
    rowptr, colidx, val = _canon(F64, rowptr, colidx, val)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(m, dtype=np.float64)
    if threads == 1:
        lib().dasp_oracle_csr_spmv_f64(m, _p(rowptr), _p(colidx), _p(val), _p(x), _p(y))
    else:
        lib().dasp_oracle_csr_spmv_f64_mt(m, _p(rowptr), _p(colidx), _p(val), _p(x), _p(y), threads)
    return y


def csr_spmv_f16(m, rowptr, colidx, val, x) -> np.ndarray:
    """half inputs, double accumulation; returns float64 (caller rounds)."""
    rowptr, colidx, val = _canon(F16, rowptr, colidx, val)
    x = np.ascontiguousarray(x, dtype=np.float16)
    y = np.empty(m, dtype=np.float64)
    lib().dasp_oracle_csr_spmv_f16(m, _p(rowptr), _p(colidx), _p(val), _p(x), _p(y))
    return y


# ------------------------------------------------------------------------------------------------
# oracle/_ref: the unmodified reference

_REF_ARRAYS = ["long_val", "long_cid", "long_rpt_new", "short_val", "short_cid", "reg_val", "reg_cid",
               "blockPtr", "irreg_val", "irreg_rpt", "irreg_cid"]
_ref_libs: dict = {}


def ref_path(dtype: int) -> str:
    return os.path.join(_HERE, "_ref", "libdasp_ref_%s.so" % ("f16" if dtype == F16 else "f64"))


def ref_available(dtype: int) -> bool:
    return os.path.exists(ref_path(dtype))


def _ref(dtype: int) -> C.CDLL:
    if dtype not in _ref_libs:
        L = C.CDLL(ref_path(dtype))
        L.dasp_ref_spmv_all.restype = C.c_int
        L.dasp_ref_spmv_all.argtypes = [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_double, C.c_int]
        L.dasp_ref_get.restype = C.c_long
        L.dasp_ref_get.argtypes = [C.c_char_p, C.c_void_p, C.c_long]
        L.dasp_ref_csv.restype = C.c_char_p
        L.dasp_ref_has_gpu.restype = C.c_int
        _ref_libs[dtype] = L
    return _ref_libs[dtype]


def ref_spmv_all(dtype: int, m: int, n: int, rowptr, colidx, val, x=None, threshold: float = 0.75,
                 block_longest: int = 256) -> dict:
    """Call the reference's ``spmv_all``.  Returns its uploaded arrays, ``order_rid``, the CSV record
    it wrote, and — only when a GPU is present (``ran_on_gpu``) — its y in permuted order."""
    rowptr, colidx, val = _canon(dtype, rowptr, colidx, val)
    nnz = int(rowptr[m])
    npdt = _val_np(dtype)
    # the FP16 upload reads 2*ceil(n/2) halves (src/dasp_f16.h:1501): pad x by one element
    xs = np.ones(n + 2, dtype=npdt) if x is None else np.concatenate(
        [np.asarray(x, dtype=npdt), np.zeros(2, dtype=npdt)])
    y = np.zeros(m + 2, dtype=npdt)
    order = np.zeros(max(m, 1), dtype=np.int32)
    L = _ref(dtype)
    gpu = L.dasp_ref_spmv_all(_p(val), _p(rowptr), _p(colidx), _p(xs), _p(y), _p(order), m, n, nnz,
                              threshold, block_longest)
    out = {"order_rid": order[:m].copy(), "ran_on_gpu": bool(gpu), "y_perm": y[:m].copy(),
           "csv": L.dasp_ref_csv().decode()}
    for a in _REF_ARRAYS:
        nbytes = L.dasp_ref_get(a.encode(), None, 0)
        dt = npdt if a.endswith("_val") else np.int32
        if nbytes <= 0:
            out[a] = np.zeros(0, dtype=dt)
            continue
        buf = np.empty(nbytes, dtype=np.uint8)
        L.dasp_ref_get(a.encode(), _p(buf), nbytes)
        out[a] = buf.view(dt).copy()
    return out


def ref_read_mtx(dtype: int, path: str):
    """The reference's own mmio_allinone (src/mmio_highlevel.h:608) from oracle/_ref."""
    L = _ref(dtype)
    L.dasp_ref_mmio_allinone.restype = C.c_int
    L.dasp_ref_mmio_allinone.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 4 + [C.POINTER(C.c_void_p)] * 3
    L.dasp_ref_free.argtypes = [C.c_void_p]
    L.dasp_ref_free.restype = None
    m, n, nnz, sym = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rp, ci, va = C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = L.dasp_ref_mmio_allinone(os.fsencode(path), C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym), C.byref(rp),
                                  C.byref(ci), C.byref(va))
    if rc != 0:
        return rc, None
    k = nnz.value
    rowptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int32)), shape=(m.value + 1,)).copy()
    colidx = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_int32)), shape=(max(k, 1),))[:k].copy()
    vt = C.c_uint16 if dtype == F16 else C.c_double
    val = np.ctypeslib.as_array(C.cast(va, C.POINTER(vt)), shape=(max(k, 1),))[:k].copy()
    if dtype == F16:
        val = val.view(np.float16)
    for p in (rp, ci, va):
        L.dasp_ref_free(p)
    return 0, (m.value, n.value, rowptr, colidx, val, bool(sym.value))
