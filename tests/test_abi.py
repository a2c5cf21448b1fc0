"""The C-ABI library builds, loads on a CPU-only box and exports every symbol include/vl2.h declares
(no compute calls here: without a GPU every compute entry point must fail with VL2_E_CUDA, never fall back)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from videollama2_b200 import _lib, build
    build.build()
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vl2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vl2_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_binding_list():
    from videollama2_b200 import _lib
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for sym in header_symbols():
        assert hasattr(lib, sym), f"libvl2.so does not export {sym}"
    assert lib.vl2_version() == 100
    assert isinstance(lib.vl2_last_error(), bytes)


def test_struct_layouts_match_header():
    from videollama2_b200._lib import AttnArgs, GemmArgs
    assert ctypes.sizeof(GemmArgs) == 6 * 8 + 4 * 8 + 6 * 4 + 8 * 8 + 8 + 2 * 4 + 2 * 8 + 2 * 4 + 2 * 4 + 2 * 8
    assert ctypes.sizeof(AttnArgs) == 4 * 8 + 4 * 8 + 8 * 4


def test_argument_validation_without_gpu(lib):
    """Shape/alignment checks run before any CUDA call, so they are testable on CPU."""
    from videollama2_b200._lib import GemmArgs
    a = GemmArgs(A=16, W=16, C=16, lda=8, ldw=8, ldc=8, ldr=0, M=4, N=8, K=12, act=0, out_f32=0)
    assert lib.vl2_gemm_bf16(ctypes.byref(a), None) == -1 and b"multiples of 8" in lib.vl2_last_error()
    a = GemmArgs(A=17, W=16, C=16, lda=8, ldw=8, ldc=8, ldr=0, M=4, N=8, K=8, act=0, out_f32=0)
    assert lib.vl2_gemm_bf16(ctypes.byref(a), None) == -3
    a = GemmArgs(A=16, W=16, C=16, lda=8, ldw=8, ldc=8, ldr=0, M=4, N=8, K=8, act=9, out_f32=0)
    assert lib.vl2_gemm_bf16(ctypes.byref(a), None) == -6


def test_ops_refuse_cpu_tensors():
    import torch
    from videollama2_b200 import ops
    from videollama2_b200._lib import Vl2Error
    x = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(Vl2Error):
        ops.gemm(x, x)
    with pytest.raises(Vl2Error):
        ops.layernorm(x, x[0], x[0], 1e-5)
