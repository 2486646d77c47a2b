"""CPU tier: libb200z.so loads, exports every symbol include/b200z.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200z.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200z_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import sharpziplib_b200 as z
    L = z.lib()
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(L, s), "libb200z.so does not export %s" % s
    assert sorted(z.EXPORTS) == syms, "python binding and header disagree"


def test_static_tables_roundtrip_without_gpu():
    import sharpziplib_b200 as z
    L = z.lib()
    n = L.b200z_static_tables_size()
    buf = (C.c_uint8 * n)()
    assert L.b200z_static_tables_export(buf, n) == 0
    assert L.b200z_static_tables_import(buf, n) == 0
    buf[5] ^= 1
    assert L.b200z_static_tables_import(buf, n) == 3  # B200Z_E_DATA


def test_argument_errors_need_no_gpu():
    import sharpziplib_b200 as z
    with pytest.raises(ValueError):
        z.Deflater(10)          # ArgumentOutOfRangeException (Deflater.cs:184-187)
    with pytest.raises(ValueError):
        z.Deflater(-2)
    d = z.Deflater(-1, True)    # DEFAULT_COMPRESSION -> 6
    assert d.GetLevel() == 6
    d.Finish()
    with pytest.raises(z.InvalidOperationException):
        d.SetInput(b"x")        # "Finish() already called" (Deflater.cs:335)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import sharpziplib_b200 as z
    with pytest.raises(z.B200zCudaError):
        z.deflate_batch([b"hello hello hello"], level=6)
    with pytest.raises(z.B200zCudaError):
        c = z.Crc32()
        c.Update(b"123456789")


def test_header_is_plain_c(tmp_path):
    """include/b200z.h is what a cgo / P/Invoke / ctypes binding declares: it must compile as C99, no C++ in the signatures"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_check.c"
    src.write_text('#include "b200z.h"\nint main(void) { b200z_history h; (void)h; return b200z_version() > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only",
                           "-I", os.path.join(root, "include"), str(src)])
