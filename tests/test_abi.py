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
    assert ctypes.sizeof(GemmArgs) == 6 * 8 + 4 * 8 + 6 * 4 + 8 * 8 + 8 + 2 * 4 + 2 * 8 + 2 * 4 + 2 * 4 + 2 * 8 + 3 * 8 + 6 * 4 + 8 + 4 * 4 
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


C_CONSUMER = r"""
#include <stdio.h>
#include <string.h>
#include "vl2.h"
/* A plain C99 consumer of the C-ABI: no torch, no C++.  Prints the struct sizes the Python binding must agree with and
 * checks that argument validation happens before any CUDA call (so it also runs on a box without a GPU). */
int main(void) {
  vl2_gemm_args g;
  vl2_attn_args a;
  vl2_preprocess_args pp;
  memset(&g, 0, sizeof g); memset(&a, 0, sizeof a); memset(&pp, 0, sizeof pp);
  printf("%d %zu %zu %zu\n", vl2_version(), sizeof g, sizeof a, sizeof pp);
  g.M = 4; g.N = 8; g.K = 12;                       /* K %% 8 != 0 */
  int rc = vl2_gemm_bf16(&g, NULL);
  printf("%d %s\n", rc, vl2_last_error());
  rc = vl2_attention(NULL, NULL);
  printf("%d\n", rc);
  rc = vl2_preprocess_frames(&pp, NULL);
  printf("%d\n", rc);
  printf("%zu %zu\n", vl2_attention_decode_workspace(32, 8, 128), vl2_preprocess_workspace(&pp));
  return 0;
}
"""


def test_plain_c_consumer_links_and_validates(lib, tmp_path):
    """include/vl2.h is valid C99, a C program links against libvl2.so with nothing but the header, the struct sizes
    match the ctypes mirrors, and bad arguments are rejected with VL2_E_* codes before any device work."""
    import shutil
    import subprocess
    from videollama2_b200 import _lib, preprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "consumer.c"
    src.write_text(C_CONSUMER)
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-l:libvl2.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    ver, sz_gemm, sz_attn, sz_pp = out[0].split()
    assert int(ver) == 100
    assert int(sz_gemm) == ctypes.sizeof(_lib.GemmArgs) and int(sz_attn) == ctypes.sizeof(_lib.AttnArgs)
    assert int(sz_pp) == ctypes.sizeof(preprocess.PreprocessArgs)
    assert out[1].startswith("-") and "multiples of 8" in out[1]          # VL2_E_BADSHAPE + message
    assert int(out[2]) < 0 and int(out[3]) < 0
    ws_attn, ws_pp = out[4].split()
    assert int(ws_attn) == 32 * (296 // 8) * 130 * 4 and int(ws_pp) == 0


def test_gemm_plan_for_the_path_shapes(lib):
    """vl2_gemm_plan is host logic only: the tile the cost model picks and how many rounds the persistent grid runs for
    the GEMMs of config 2 (148 SMs assumed without a device).  Pins the scheduler against accidental changes."""
    def plan(M, N, K):
        out = (ctypes.c_int32 * 6)()
        assert lib.vl2_gemm_plan(M, N, K, 0, out) == 0
        return dict(zip(("bn", "pair", "tiles", "slots", "rounds", "sms"), out))
    gate_up = plan(1776, 28672, 4096)            # decoder gate/up: the 256 x 256 cta_group::2 tile, 784 tiles on 74 pairs
    assert (gate_up["bn"], gate_up["pair"], gate_up["tiles"], gate_up["slots"], gate_up["rounds"]) == (256, 1, 784, 74, 11)
    for M, N, K in [(1776, 6144, 4096), (1776, 4096, 4096), (1776, 4096, 14336), (9232, 3072, 1024), (9232, 4096, 1024),
                    (9232, 1024, 4096), (11664, 1152, 1152)]:
        p = plan(M, N, K)
        assert p["pair"] == 1 and p["bn"] in (224, 256, 416) and p["sms"] == 148   # 416 = the wide 224 + 192 tile
        tile_m = 256
        assert p["tiles"] == -(-M // tile_m) * -(-N // p["bn"]) and p["rounds"] == -(-p["tiles"] // p["slots"])
    o_proj, down = plan(1776, 4096, 4096), plan(1776, 4096, 14336)   # one round of wide tiles instead of two narrow ones
    assert (o_proj["bn"], o_proj["rounds"], down["bn"], down["rounds"]) == (416, 1, 416, 1)
    assert plan(1776, 6144, 4096)["bn"] == 224 and plan(9232, 1024, 1024)["bn"] == 256
    small = plan(300, 520, 256)                  # a matrix narrower than a wide tile gets a narrow single-CTA tile
    assert small["pair"] == 0 and small["bn"] <= 128 and small["rounds"] == 1
    out = (ctypes.c_int32 * 6)()
    assert lib.vl2_gemm_plan(0, 8, 8, 0, out) < 0
