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


def test_host_side_schedulers_fill_whole_waves(lib):
    """pure host logic (no GPU): CTAs per (batch, head) of the slice kernels and the split-K factor of the token-contraction
    GEMM are chosen against the 148-SM wave size; shape predicates of the tensor-core kernels."""
    from transformerbasednavierstokesolver_b200 import ops
    g = lib.tbns_slice_groups(20, 4096, 8)              # bench shape: 160 (b,h) pairs, 32 chunks of 128 tokens each
    assert 1 <= g <= 32
    ctas = 160 * g
    assert ctas / (-(-ctas // 296) * 296) > 0.9         # two-CTA-per-SM backward kernel: >= 90 % of whole waves
    assert lib.tbns_slice_groups(1, 100, 1) == 1        # one chunk: nothing to split
    for tiles, kblocks in ((2, 1280), (36, 1280), (40, 64), (1, 16)):
        s = ops._wgrad_split(tiles, kblocks)
        assert 1 <= s <= max(1, kblocks // 4)            # at least four 64-token k-blocks per CTA
    assert ops._wgrad_split(36, 1280) * 36 <= 148        # conv wgrad at the bench shape: a single wave
    assert lib.tbns_gemm_tc_supported(256, 512, 9) == 1 and lib.tbns_gemm_tc_supported(74, 512, 1) == 0
    assert lib.tbns_gemm_tc_wgrad_supported(256, 512, 9) == 1 and lib.tbns_gemm_tc_wgrad_supported(100, 512, 1) == 0
    assert lib.tbns_pa_slice_tc_supported(32, 32) == 1 and lib.tbns_pa_slice_tc_supported(8, 32) == 0
    assert lib.tbns_ln_linear1_supported(256) == 1 and lib.tbns_ln_linear1_supported(96) == 0
