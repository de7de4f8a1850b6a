"""dasp_b200 — B200-native DASP SpMV.

The product is the C-ABI shared library ``dasp_b200/libdasp_b200.so`` (``include/dasp.h``), built
from ``dasp_b200/csrc`` for sm_100a.  This package is only a thin ctypes binding used by the
tests and the benchmark driver; it mirrors the reference's single entry point ``spmv_all``
(``/root/reference/src/dasp_f64.h:486``) as :func:`spmv_all` and the analyse/execute split as
:class:`Dasp`.  There is no CPU fallback: importing works without a GPU (so the symbol table can
be checked), every compute call fails loudly without one.
"""
from .lib import (DASP_F16, DASP_F64, VARIANT_AUTO, VARIANT_CUDA_CORE, VARIANT_MMA, VARIANT_SPLIT, VARIANT_TMA, VARIANT_BLOCKED, VARIANT_BANDED, Dasp, DaspError,
                  build, exported_symbols, library_path, load, partition_rows, read_mtx, scale_copy_to, scale_rsqrt, spmv_all, sumsq)

__all__ = ["DASP_F16", "DASP_F64", "VARIANT_AUTO", "VARIANT_CUDA_CORE", "VARIANT_MMA", "VARIANT_SPLIT", "VARIANT_TMA", "VARIANT_BLOCKED", "VARIANT_BANDED", "Dasp",
           "DaspError", "build", "exported_symbols", "library_path", "load", "partition_rows",
           "read_mtx", "scale_copy_to", "scale_rsqrt", "spmv_all", "sumsq"]
