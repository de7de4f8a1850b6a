"""ctypes binding of libdasp_synth.so (include/dasp_synth.h): benchmark matrices generated on the
GPU.  Bench/test tooling; torch is used only to own the device buffers and for the prefix sum."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

STENCIL27, POWERLAW, SKEWED, BANDED, POWERLAW_SPEC, SKEWED_SPEC = 0, 1, 2, 3, 4, 5


class Spec(C.Structure):
    _fields_ = [("kind", C.c_int), ("m", C.c_int64), ("n", C.c_int64), ("seed", C.c_uint64),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("alpha", C.c_double), ("lmax", C.c_int), ("window", C.c_int),
                ("n_long", C.c_int), ("long_len", C.c_int), ("band_lo", C.c_int64), ("band", C.c_int64),
                ("mean_len", C.c_int)]

    def describe(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


def stencil27(nx, ny=None, nz=None, seed=20240004) -> Spec:
    ny, nz = ny or nx, nz or nx
    return Spec(kind=STENCIL27, m=nx * ny * nz, n=nx * ny * nz, seed=seed, nx=nx, ny=ny, nz=nz)


def powerlaw(m=10_000_000, alpha=0.95, lmax=1_000_000, window=4096, seed=20240001) -> Spec:
    return Spec(kind=POWERLAW, m=m, n=m, seed=seed, alpha=alpha, lmax=min(lmax, m // 2), window=window)


def skewed(n_long=1000, long_len=1_000_000, n_short=50_000_000, window=4096, seed=20240005) -> Spec:
    m = n_long + n_short
    band = 1
    while band < 2 * long_len:
        band <<= 1
    band = min(band, 1 << (m.bit_length() - 1))
    return Spec(kind=SKEWED, m=m, n=m, seed=seed, n_long=n_long, long_len=long_len, window=window,
                band_lo=(m - band) // 2, band=band)


def powerlaw_spec(m=10_000_000, alpha=0.95, lmax=1_000_000, window=4096, seed=20240001) -> Spec:
    """C3 exactly as SURVEY.md §8(d) words it: unsorted distinct columns, 90 % in the +-4096 window, 10 % global."""
    return Spec(kind=POWERLAW_SPEC, m=m, n=m, seed=seed, alpha=alpha, lmax=min(lmax, m // 2), window=window)


def skewed_spec(n_long=1000, long_len=1_000_000, n_short=50_000_000, window=4096, seed=20240005) -> Spec:
    """C5 exactly as SURVEY.md §8(d) words it: long rows at seeded positions with columns uniform over n without
    replacement, short rows with distinct random columns of the +-4096 window."""
    m = n_long + n_short
    return Spec(kind=SKEWED_SPEC, m=m, n=m, seed=seed, n_long=n_long, long_len=min(long_len, m // 2), window=window)


def banded(m=121_192, mean_len=22, window=2048, seed=7) -> Spec:
    """cop20k_A stand-in: the real file is a missing blob (.MISSING_LARGE_BLOBS)."""
    return Spec(kind=BANDED, m=m, n=m, seed=seed, mean_len=mean_len, window=window)


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libdasp_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.dasp_synth_rowlen.argtypes = [C.POINTER(Spec), C.c_int64, C.c_int64, vp, vp]
        L.dasp_synth_fill.argtypes = [C.POINTER(Spec), C.c_int64, C.c_int64, vp, vp, vp, vp]
        L.dasp_synth_to_half.argtypes = [vp, vp, C.c_int64, vp]
        L.dasp_synth_flush_l2.argtypes = [vp, C.c_int64, vp]
        L.dasp_synth_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("dasp_synth: " + load().dasp_synth_last_error().decode())


def row_lengths(spec: Spec, row0: int, row1: int, device):
    import torch

    L = load()
    ln = torch.empty(max(row1 - row0, 1), dtype=torch.int32, device=device)
    st = torch.cuda.current_stream(device).cuda_stream
    _check(L.dasp_synth_rowlen(C.byref(spec), row0, row1, ln.data_ptr(), st))
    return ln[: row1 - row0]


def generate(spec: Spec, row0: int, row1: int, device, half: bool = False):
    """Rows [row0,row1) of the global matrix as device CSR: (rowptr int32, colidx int32, val f64|f16)."""
    import torch

    L = load()
    st = torch.cuda.current_stream(device).cuda_stream
    rows = row1 - row0
    ln = row_lengths(spec, row0, row1, device)
    rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    torch.cumsum(ln, 0, out=rowptr[1:])
    nnz = int(rowptr[-1].item())
    if nnz >= 2 ** 31:
        raise RuntimeError(f"slab nnz {nnz} exceeds 32-bit row pointers")
    rowptr = rowptr.to(torch.int32)
    colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)
    val = torch.empty(max(nnz, 1), dtype=torch.float64, device=device)
    _check(L.dasp_synth_fill(C.byref(spec), row0, row1, rowptr.data_ptr(), colidx.data_ptr(), val.data_ptr(), st))
    if half:
        hv = torch.empty(max(nnz, 1), dtype=torch.float16, device=device)
        _check(L.dasp_synth_to_half(val.data_ptr(), hv.data_ptr(), nnz, st))
        val = hv
    torch.cuda.synchronize(device)
    return rowptr, colidx[:nnz], val[:nnz], nnz


def flush_l2(scratch) -> None:
    import torch

    _check(load().dasp_synth_flush_l2(scratch.data_ptr(), scratch.numel() * scratch.element_size(),
                                      torch.cuda.current_stream(scratch.device).cuda_stream))
