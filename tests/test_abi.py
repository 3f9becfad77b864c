"""CPU: libdm_b200.so builds for sm_100a, loads, and exports every symbol include/dm_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from disentangle_mlp_b200 import _lib, build

    build.build()
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "dm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = header_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header():
    from disentangle_mlp_b200 import _lib

    assert sorted(_lib.EXPORTED) == header_functions()


def test_version_and_error_string(lib):
    assert lib.dm_version() == 100
    assert isinstance(lib.dm_last_error(), bytes)
    assert lib.dm_launch_count() >= 0


def test_struct_layouts_match_c():
    from disentangle_mlp_b200 import _lib

    # dm_bn_fuse: ptr, int (+pad), ll, 5 ptr, 2 float, 2 ptr; dm_gemm_desc: 4 ints, (ptr, ll) x2, ptr, 2 ll, 2 ints, ptr,
    # 4 ints, dm_bn_fuse -> natural alignment (the numbers are what gcc's sizeof gives for include/dm_b200.h)
    assert ctypes.sizeof(_lib.BnFuse) == 88
    assert ctypes.sizeof(_lib.GemmDesc) == 16 + 16 + 16 + 24 + 8 + 8 + 16 + 88
    assert ctypes.sizeof(_lib.ConvGeom) == 32


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof as gcc sees include/dm_b200.h == the ctypes mirrors in _lib.py"""
    from disentangle_mlp_b200 import _lib

    src = tmp_path / "sz.c"
    fields = [("dm_bn_fuse", f[0], _lib.BnFuse) for f in _lib.BnFuse._fields_] + \
             [("dm_gemm_desc", f[0], _lib.GemmDesc) for f in _lib.GemmDesc._fields_] + \
             [("dm_conv_geom", f[0], _lib.ConvGeom) for f in _lib.ConvGeom._fields_]
    body = "".join(f'printf("%zu\\n", offsetof({t}, {n}));\n' for t, n, _ in fields)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{ROOT}/include/dm_b200.h"\n'
                   f'int main(void){{{body}printf("%zu %zu %zu\\n", sizeof(dm_bn_fuse), sizeof(dm_gemm_desc), '
                   f'sizeof(dm_conv_geom));return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    offs = [int(v) for v in out[:len(fields)]]
    assert offs == [getattr(cls, n).offset for _, n, cls in fields]
    assert [int(v) for v in out[len(fields):]] == [ctypes.sizeof(_lib.BnFuse), ctypes.sizeof(_lib.GemmDesc),
                                                   ctypes.sizeof(_lib.ConvGeom)]


def test_workspace_query(lib):
    def ws(op, *dims):
        arr = (ctypes.c_longlong * len(dims))(*dims)
        return lib.dm_workspace_bytes(op, arr, len(dims))

    assert ws(0, 64, 2048, 16384, 18) == 0 and ws(1, 64, 8, 8, 256, 256) == 0  # GEMM, implicit-GEMM convolutions
    assert ws(3, 256, 128) == 25 * 256 * 128 * 4
    assert ws(4, 32, 1) == 5 * 64 * 64 * 4 and ws(4, 64, 2) == 5 * 64 * 64 * 4
    assert ws(5, 256, 3) == 4 * lib.dm_bn_scratch_floats(256, 3)
    assert ws(6, 192) == 2 * lib.dm_pim_elems(192)
    assert ws(7, 64, 2048) == lib.dm_bn_parts(64, 2048) * 2048 * 4
    assert ws(99) == -1 and b"unknown op" in lib.dm_last_error()


def test_argument_errors_do_not_need_a_gpu(lib):
    from disentangle_mlp_b200 import _lib

    g = _lib.ConvGeom(4, 8, 8, 256, 17, 16, 256, 2)  # hb != hs*stride
    rc = lib.dm_conv_down(ctypes.byref(g), None, None, None, None, None, None)
    assert rc != 0 and b"stride x small" in lib.dm_last_error()
    d = _lib.GemmDesc(layout=7, m=1, n=1, k=1, lda=8, ldb=8)
    assert lib.dm_gemm_bf16(ctypes.byref(d), None) != 0


def test_sass_is_blackwell_native():
    from disentangle_mlp_b200.build import LIB_PATH

    out = subprocess.run(["cuobjdump", "-sass", str(LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in out, "TMA loads missing from SASS"
    assert "LDTM" in out, "tcgen05.ld missing from SASS"
    assert "sm_100a" in out or "sm_100" in out
