"""ctypes binding of libdasp_b200.so (include/dasp.h).  No torch, no numpy compute: plumbing only."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

DASP_F64, DASP_F16 = 0, 1
VARIANT_AUTO, VARIANT_CUDA_CORE, VARIANT_MMA, VARIANT_SPLIT, VARIANT_TMA, VARIANT_BLOCKED, VARIANT_BANDED = 0, 1, 2, 3, 4, 5, 6

_STATS_INT = [
    "dtype", "m", "n", "nnz",  # nnz is int64, handled below
    "row_long", "row_block", "row_zero", "short_row_1", "short_row_3", "short_row_2", "short_row_4",
    "common_13", "short_row_34", "rowloop", "blocknum", "warp_number", "BlockNum_long", "fill0_nnz_long",
    "fill0_nnz_reg", "nnz_irreg", "origin_nnz_reg", "fill0_nnz_short", "fill0_nnz_short13",
    "fill0_nnz_short34", "fill0_nnz_short22", "threadblock13", "threadblock34", "threadblock22",
    "nnz_short", "nnz_long", "BlockNum", "BlockNum_short_1", "BlockNum_all", "sumBlockNum", "fill0_nnz_irreg",
]


class _Stats(C.Structure):
    _fields_ = ([(n, C.c_int64 if n == "nnz" else C.c_int) for n in _STATS_INT]
                + [("rate_fill0", C.c_double), ("data_X", C.c_int64), ("data_X2", C.c_int64),
                   ("data_origin1", C.c_int64), ("preprocess_ms", C.c_double), ("device_bytes", C.c_int64),
                   ("col_min", C.c_int), ("col_max", C.c_int), ("long_gather_lines", C.c_double),
                   ("long_blocked", C.c_int), ("short_banded", C.c_int), ("short_band_hit_rate", C.c_double),
                   ("medium_gather_lines", C.c_double), ("medium_band_hit_rate", C.c_double), ("medium_banded", C.c_int),
                   ("reserved_", C.c_int)])


def stats_struct_size() -> int:
    return C.sizeof(_Stats)


ARRAYS = ["order_rid", "long_rpt_new", "long_val", "long_cid", "blockPtr", "irreg_rpt", "irreg_val",
          "irreg_cid", "reg_val", "reg_cid", "short_val", "short_cid"]


class DaspError(RuntimeError):
    pass


def library_path() -> str:
    # DASP_B200_LIB: A/B aid of the tuning sweeps (an alternative build of the same sources); never set in tests or the bench
    return os.environ.get("DASP_B200_LIB") or os.path.join(_HERE, "libdasp_b200.so")


def build(verbose: bool = False) -> str:
    """Compile dasp_b200/csrc for sm_100a into dasp_b200/libdasp_b200.so (nvcc cross-compiles)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return library_path()


def declared_symbols() -> list:
    """Every function include/dasp.h declares."""
    text = open(os.path.join(_ROOT, "include", "dasp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dasp_[a-z0-9_]+)\s*\(", text)))


def exported_symbols() -> list:
    out = subprocess.check_output(["nm", "-D", "--defined-only", library_path()], text=True)
    return sorted(l.split()[-1] for l in out.splitlines() if " T " in l)


_lib = None


def load() -> C.CDLL:
    """dlopen the product library; raises if it was not built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise DaspError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        vp, ip = C.c_void_p, C.c_int
        L.dasp_create.argtypes = [C.POINTER(vp), ip, ip, ip, ip, C.c_int64, vp, vp, vp, C.c_double, ip]
        L.dasp_spmv.argtypes = [vp, vp, vp, vp]
        L.dasp_spmv_unpermuted.argtypes = [vp, vp, vp, vp]
        L.dasp_spmv_host.argtypes = [vp, vp, vp]
        L.dasp_spmv_scatter_to.argtypes = [vp, vp, C.POINTER(vp), ip, C.c_int64, vp, vp]
        L.dasp_unpermute_to.argtypes = [vp, vp, C.POINTER(vp), ip, C.c_int64, vp, vp]
        L.dasp_spmv_permuted_to.argtypes = [vp, vp, C.POINTER(vp), ip, C.c_int64, vp, vp]
        L.dasp_relabel_columns.argtypes = [vp, vp, ip]
        L.dasp_inverse_order.argtypes = [vp, C.POINTER(vp)]
        L.dasp_spmv_host_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), ip]
        L.dasp_spmv_axpby.argtypes = [vp, C.c_double, vp, C.c_double, vp, ip, vp]
        L.dasp_spmv_f16_f32out.argtypes = [vp, vp, vp, ip, vp]
        L.dasp_save.argtypes = [vp, C.c_char_p]
        L.dasp_load.argtypes = [C.POINTER(vp), C.c_char_p, ip]
        L.dasp_spmv_timed.argtypes = [vp, vp, vp, vp, ip, ip, C.POINTER(C.c_float)]
        L.dasp_order.argtypes = [vp, C.POINTER(vp)]
        L.dasp_stats.argtypes = [vp, C.POINTER(_Stats)]
        L.dasp_export.argtypes = [vp, C.c_char_p, vp, C.c_int64, C.POINTER(C.c_int64)]
        L.dasp_set_variant.argtypes = [vp, ip, ip, ip]
        L.dasp_report.argtypes = [vp, C.c_char_p, C.c_double, C.c_char_p, C.c_int64]
        L.dasp_launches_per_spmv.argtypes = [vp]
        L.dasp_set_category_mask.argtypes = [vp, ip]
        L.dasp_set_index_compression.argtypes = [vp, ip]
        L.dasp_destroy.argtypes = [vp]
        L.dasp_strerror.restype = C.c_char_p
        L.dasp_strerror.argtypes = [ip]
        L.dasp_last_error.restype = C.c_char_p
        L.dasp_spmv_all_f64.argtypes = [C.c_char_p] + [vp] * 6 + [ip] * 4 + [C.c_double, ip]
        L.dasp_spmv_all_f16.argtypes = [C.c_char_p] + [vp] * 6 + [ip] * 4 + [C.c_double, ip]
        L.dasp_partition_rows.argtypes = [ip, vp, ip, vp]
        L.dasp_read_mtx.argtypes = [C.c_char_p, ip, C.POINTER(ip), C.POINTER(ip), C.POINTER(C.c_int64), C.POINTER(ip),
                                    C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.dasp_free_host.argtypes = [vp]
        L.dasp_free_host.restype = None
        L.dasp_sumsq.argtypes = [vp, C.c_int64, vp, vp]
        L.dasp_scale_rsqrt.argtypes = [vp, C.c_int64, vp, vp]
        L.dasp_scale_copy_to.argtypes = [vp, C.c_int64, C.POINTER(vp), ip, C.c_int64, vp, vp]
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        L = load()
        raise DaspError(f"{what}: {L.dasp_strerror(rc).decode()} ({rc}): {L.dasp_last_error().decode()}")


def _ptr(a) -> C.c_void_p:
    """host numpy array, torch tensor (host or device) or raw integer address -> void*"""
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())  # torch tensor


def _np_val(dtype: int):
    return np.float16 if dtype == DASP_F16 else np.float64


class Dasp:
    """Analyse once (GPU preprocessing), multiply many times.  Mirrors the two halves of the
    reference's ``spmv_all``: lines 499-1157 (analyse) and 1285-1402 (execute) of src/dasp_f64.h."""

    def __init__(self, dtype: int, m: int, n: int, rowptr, colidx, val, device: int = 0,
                 threshold: float = 0.75, block_longest: int = 256, nnz: int | None = None):
        self._h = C.c_void_p(None)
        self.dtype, self.m, self.n = dtype, m, n
        if isinstance(rowptr, np.ndarray):
            rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
            colidx = np.ascontiguousarray(colidx, dtype=np.int32)
            val = np.ascontiguousarray(val, dtype=_np_val(dtype))
            if nnz is None:
                nnz = int(rowptr[m])
        if nnz is None:
            raise ValueError("nnz is required for device CSR input")
        self._keep = (rowptr, colidx, val)
        _check(load().dasp_create(C.byref(self._h), dtype, device, m, n, nnz, _ptr(rowptr), _ptr(colidx),
                                  _ptr(val), threshold, block_longest), "dasp_create")
        self._keep = None
        self.nnz = nnz

    # -- execute ---------------------------------------------------------------------------------
    def spmv(self, d_x, d_y, stream: int = 0) -> None:
        """y (permuted order) = A x; d_x/d_y: device tensors or raw device addresses."""
        _check(load().dasp_spmv(self._h, _ptr(d_x), _ptr(d_y), C.c_void_p(stream)), "dasp_spmv")

    def spmv_unpermuted(self, d_x, d_y, stream: int = 0) -> None:
        _check(load().dasp_spmv_unpermuted(self._h, _ptr(d_x), _ptr(d_y), C.c_void_p(stream)), "dasp_spmv_unpermuted")

    def spmv_axpby(self, alpha: float, d_x, beta: float, d_y, permuted: bool = True, stream: int = 0) -> None:
        """y = alpha*A*x + beta*y."""
        _check(load().dasp_spmv_axpby(self._h, alpha, _ptr(d_x), beta, _ptr(d_y), 1 if permuted else 0, C.c_void_p(stream)),
               "dasp_spmv_axpby")

    def spmv_f32out(self, d_x, d_y_f32, permuted: bool = True, stream: int = 0) -> None:
        """FP16 matrix and x, y stored as float32 (the fp32 accumulator, unrounded)."""
        _check(load().dasp_spmv_f16_f32out(self._h, _ptr(d_x), _ptr(d_y_f32), 1 if permuted else 0, C.c_void_p(stream)),
               "dasp_spmv_f16_f32out")

    def save(self, path: str) -> None:
        _check(load().dasp_save(self._h, os.fsencode(path)), "dasp_save")

    @classmethod
    def load_file(cls, path: str, device: int = 0) -> "Dasp":
        """Rebuild a handle from a file written by save(): no CSR, no preprocessing."""
        self = cls.__new__(cls)
        self._h = C.c_void_p(None)
        self._keep = None
        _check(load().dasp_load(C.byref(self._h), os.fsencode(path), device), "dasp_load")
        st = self.stats()
        self.dtype, self.m, self.n, self.nnz = st["dtype"], st["m"], st["n"], st["nnz"]
        return self

    def spmv_scatter_to(self, d_x, dests, row_offset: int, d_norm2=None, stream: int = 0) -> None:
        """Slab product in original row order, scaled by 1/sqrt(*d_norm2), stored at row_offset of every vector in
        `dests` (device addresses: local, peer-mapped or multicast)."""
        arr = (C.c_void_p * len(dests))(*[_ptr(d) for d in dests])
        _check(load().dasp_spmv_scatter_to(self._h, _ptr(d_x), arr, len(dests), row_offset, _ptr(d_norm2), C.c_void_p(stream)),
               "dasp_spmv_scatter_to")

    def spmv_permuted_to(self, d_x, dests, row_offset: int, d_norm2=None, stream: int = 0) -> None:
        """Slab product in PERMUTED order, scaled by 1/sqrt(*d_norm2), stored at row_offset of every vector in `dests`."""
        arr = (C.c_void_p * len(dests))(*[_ptr(d) for d in dests])
        _check(load().dasp_spmv_permuted_to(self._h, _ptr(d_x), arr, len(dests), row_offset, _ptr(d_norm2), C.c_void_p(stream)),
               "dasp_spmv_permuted_to")

    def relabel_columns(self, d_new_index, n_new: int) -> None:
        """Kernel-facing column indices := new_index[column] (device int32 array); x then lives in the new index space."""
        _check(load().dasp_relabel_columns(self._h, _ptr(d_new_index), n_new), "dasp_relabel_columns")

    def inverse_order_ptr(self) -> int:
        p = C.c_void_p(None)
        _check(load().dasp_inverse_order(self._h, C.byref(p)), "dasp_inverse_order")
        return p.value or 0

    def unpermute_to(self, d_y_perm, dests, row_offset: int, d_norm2=None, stream: int = 0) -> None:
        """Permuted slab product -> original order, scaled by 1/sqrt(*d_norm2), coalesced stores into every vector of
        `dests` (local, peer-mapped or multicast device addresses) at row_offset."""
        arr = (C.c_void_p * len(dests))(*[_ptr(d) for d in dests])
        _check(load().dasp_unpermute_to(self._h, _ptr(d_y_perm), arr, len(dests), row_offset, _ptr(d_norm2), C.c_void_p(stream)),
               "dasp_unpermute_to")

    def spmv_timed(self, d_x, d_y, stream: int = 0, warmup: int = 0, reps: int = 1) -> float:
        """`reps` back-to-back launches issued from C; returns their total device time in ms."""
        ms = C.c_float(0.0)
        _check(load().dasp_spmv_timed(self._h, _ptr(d_x), _ptr(d_y), C.c_void_p(stream), warmup, reps, C.byref(ms)),
               "dasp_spmv_timed")
        return float(ms.value)

    def spmv_host(self, x_host, y_host=None):
        """Host buffers in, host buffers out (H2D x, kernel, D2H y); y in permuted order."""
        if isinstance(x_host, np.ndarray):
            x_host = np.ascontiguousarray(x_host, dtype=_np_val(self.dtype))
        if y_host is None:
            y_host = np.empty(self.m, dtype=_np_val(self.dtype))
        _check(load().dasp_spmv_host(self._h, _ptr(x_host), _ptr(y_host)), "dasp_spmv_host")
        return y_host

    def spmv_host_batch(self, x_hosts, y_hosts) -> None:
        """Independent products y_j = A x_j on (preferably pinned) host buffers, copies and kernels pipelined."""
        k = len(x_hosts)
        assert len(y_hosts) == k
        xs = (C.c_void_p * k)(*[_ptr(x) for x in x_hosts])
        ys = (C.c_void_p * k)(*[_ptr(y) for y in y_hosts])
        _check(load().dasp_spmv_host_batch(self._h, xs, ys, k), "dasp_spmv_host_batch")

    # -- inspect ---------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = _Stats()
        _check(load().dasp_stats(self._h, C.byref(s)), "dasp_stats")
        return {n: getattr(s, n) for n, _ in _Stats._fields_}

    def export(self, name: str) -> np.ndarray:
        nbytes = C.c_int64(0)
        _check(load().dasp_export(self._h, name.encode(), None, 0, C.byref(nbytes)), "dasp_export")
        dt = _np_val(self.dtype) if name.endswith("_val") else np.int32
        out = np.empty(nbytes.value // np.dtype(dt).itemsize, dtype=dt)
        _check(load().dasp_export(self._h, name.encode(), _ptr(out), nbytes.value, None), "dasp_export")
        return out

    def report(self, label: str, spmv_ms: float) -> str:
        """The reference's CSV record (src/dasp_f64.h:1440-1441) for a measured time per SpMV."""
        buf = C.create_string_buffer(1024)
        k = load().dasp_report(self._h, label.encode(), spmv_ms, buf, len(buf))
        if k < 0:
            _check(k, "dasp_report")
        return buf.value.decode()

    def order_ptr(self) -> int:
        p = C.c_void_p(None)
        _check(load().dasp_order(self._h, C.byref(p)), "dasp_order")
        return p.value or 0

    def set_variant(self, medium: int = 0, long_rows: int = 0, short_rows: int = 0) -> None:
        _check(load().dasp_set_variant(self._h, medium, long_rows, short_rows), "dasp_set_variant")

    def set_index_compression(self, on: bool) -> None:
        _check(load().dasp_set_index_compression(self._h, 1 if on else 0), "dasp_set_index_compression")

    def set_category_mask(self, mask: int) -> None:
        """Profiling aid: bit 0 long, 1 medium, 2 short, 3 empty rows."""
        _check(load().dasp_set_category_mask(self._h, mask), "dasp_set_category_mask")

    def launches_per_spmv(self) -> int:
        return load().dasp_launches_per_spmv(self._h)

    def close(self) -> None:
        if self._h:
            load().dasp_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmv_all(dtype: int, csrValA, csrRowPtrA, csrColIdxA, X_val, rowA: int, colA: int, threshold: float = 0.75,
             block_longest: int = 256, filename: str = "mem"):
    """The reference's one-shot entry (src/dasp_f64.h:486): returns (Y_val permuted, order_rid)."""
    npdt = _np_val(dtype)
    val = np.ascontiguousarray(csrValA, dtype=npdt)
    rp = np.ascontiguousarray(csrRowPtrA, dtype=np.int32)
    ci = np.ascontiguousarray(csrColIdxA, dtype=np.int32)
    x = np.ascontiguousarray(X_val, dtype=npdt)
    y = np.zeros(rowA, dtype=npdt)
    order = np.zeros(rowA, dtype=np.int32)
    fn = load().dasp_spmv_all_f16 if dtype == DASP_F16 else load().dasp_spmv_all_f64
    _check(fn(filename.encode(), _ptr(val), _ptr(rp), _ptr(ci), _ptr(x), _ptr(y), _ptr(order), rowA, colA,
              int(rp[rowA]), 4, threshold, block_longest), "dasp_spmv_all")
    return y, order


def partition_rows(rowptr, parts: int) -> np.ndarray:
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    cuts = np.zeros(parts + 1, dtype=np.int32)
    _check(load().dasp_partition_rows(len(rp) - 1, _ptr(rp), parts, _ptr(cuts)), "dasp_partition_rows")
    return cuts


def sumsq(d_v, count: int, d_out, stream: int = 0) -> None:
    """*d_out = sum v[i]^2 over `count` device doubles."""
    _check(load().dasp_sumsq(_ptr(d_v), count, _ptr(d_out), C.c_void_p(stream)), "dasp_sumsq")


def scale_rsqrt(d_v, count: int, d_norm2, stream: int = 0) -> None:
    """v *= 1/sqrt(*d_norm2), all on the device."""
    _check(load().dasp_scale_rsqrt(_ptr(d_v), count, _ptr(d_norm2), C.c_void_p(stream)), "dasp_scale_rsqrt")


def read_mtx(path: str, dtype: int = DASP_F64):
    """Matrix Market -> (m, n, rowptr, colidx, val, is_symmetric) with the reference reader's semantics."""
    L = load()
    m, n, sym, nnz = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
    rp, ci, va = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(L.dasp_read_mtx(os.fsencode(path), dtype, C.byref(m), C.byref(n), C.byref(nnz), C.byref(sym), C.byref(rp),
                           C.byref(ci), C.byref(va)), "dasp_read_mtx")
    try:
        k = nnz.value
        rowptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int32)), shape=(m.value + 1,)).copy()
        colidx = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_int32)), shape=(max(k, 1),))[:k].copy()
        vt = C.c_uint16 if dtype == DASP_F16 else C.c_double
        val = np.ctypeslib.as_array(C.cast(va, C.POINTER(vt)), shape=(max(k, 1),))[:k].copy()
        if dtype == DASP_F16:
            val = val.view(np.float16)
    finally:
        for p in (rp, ci, va):
            L.dasp_free_host(p)
    return m.value, n.value, rowptr, colidx, val, bool(sym.value)


def scale_copy_to(d_v, count: int, dests, offset: int, d_norm2=None, stream: int = 0) -> None:
    """dest_p[offset + i] = v[i] / sqrt(*d_norm2) for every destination (local / peer / multicast addresses)."""
    arr = (C.c_void_p * len(dests))(*[_ptr(d) for d in dests])
    _check(load().dasp_scale_copy_to(_ptr(d_v), count, arr, len(dests), offset, _ptr(d_norm2), C.c_void_p(stream)),
           "dasp_scale_copy_to")
