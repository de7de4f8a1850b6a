"""CPU suite: the public headers are plain C (the boundary must be bindable from cgo/JNI/ctypes-style FFI)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile_c(src: str) -> None:
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.c")
        with open(path, "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), path])


def test_dasp_h_is_c99():
    _compile_c('#include "dasp.h"\n#include "dasp_synth.h"\n'
               "int use(dasp_handle *h, const void *x, void *y) { dasp_stats_t s; (void)dasp_stats(h, &s);"
               " return dasp_spmv(h, x, y, 0) + (int)sizeof(dasp_synth_spec); }\n")


def test_reference_shim_header_compiles_for_both_precisions():
    """include/dasp_reference_shim.h (the reference-side binding INTEGRATION.md describes): the reference's spmv_all argument
    list on the C ABI, compiled as plain C with the two macro settings of the reference's common.h (-D f64 / half)."""
    body = ('#define MAT_PTR_TYPE int\n#include "dasp_reference_shim.h"\n'
            "int main(void) { char name[] = \"m\"; MAT_VAL_TYPE v[1]; int rp[2] = {0, 0}, ci[1], o[1];\n"
            "  spmv_all(name, v, rp, ci, v, v, o, 0, 0, 0, 4, 0.75, 256); return 0; }\n")
    _compile_c("#define f64\n#define MAT_VAL_TYPE double\n" + body)
    _compile_c("#define MAT_VAL_TYPE unsigned short\n" + body)


def test_shim_requires_the_reference_types():
    import pytest

    with pytest.raises(subprocess.CalledProcessError):
        _compile_c('#include "dasp_reference_shim.h"\nint main(void) { return 0; }\n')


def test_oracle_header_is_c99():
    _compile_c('#include "dasp_oracle.h"\nint f(void) { return (int)sizeof(dasp_oracle_layout); }\n')
