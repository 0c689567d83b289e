"""CPU tests: libtbns.so loads and exports every symbol include/tbns.h declares; ctypes mirror of the
descriptor struct matches the C layout; host-side argument validation fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from transformerbasednavierstokesolver_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tbns.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tbns_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tbns.h but not exported by libtbns.so"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in _lib.py"


def test_gemm_desc_layout_matches_c(tmp_path):
    """compile a 10-line C program against the real header and compare sizeof/offsetof with the ctypes mirror"""
    fields = [f[0] for f in _lib.GemmDesc._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "tbns.h"\nint main(){printf("%zu\\n", sizeof(tbns_gemm_desc));\n'
    for f in fields:
        prog += f'printf("%zu\\n", offsetof(tbns_gemm_desc, {f}));\n'
    prog += "return 0;}\n"
    c = tmp_path / "l.c"
    c.write_text(prog)
    exe = tmp_path / "l"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.GemmDesc)
    for f, off in zip(fields, out[1:]):
        assert getattr(_lib.GemmDesc, f).offset == int(off), f


def test_version_and_error_string(lib):
    assert lib.tbns_version() >= 100
    assert isinstance(lib.tbns_last_error(), bytes)


def test_invalid_arguments_fail_loudly(lib):
    assert lib.tbns_gemm(None, None) == -1
    assert b"null descriptor" in lib.tbns_last_error()
    d = _lib.GemmDesc()
    d.M, d.N, d.K = 4, 4, 4
    assert lib.tbns_gemm(ctypes.byref(d), None) == -1  # null operands
    with pytest.raises(_lib.TbnsError):
        _lib.check(lib.tbns_layernorm_fwd(None, None, None, None, None, None, None, 4, 4, 1e-5, None), "ln")


def test_no_cpu_path():
    import torch
    from transformerbasednavierstokesolver_b200.model.Physics_Attention import Physics_Attention_Irregular_Mesh
    m = Physics_Attention_Irregular_Mesh(16, heads=2, dim_head=8, slice_num=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 5, 16))
