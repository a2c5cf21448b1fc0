"""Print the per-phase cycle trace of one softmax thread of the attention kernel (debug aid)."""
import ctypes as C
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import _lib
from videollama2_b200._lib import AttnArgs

lib = _lib.load()
for (B, S, Hq, Hkv, D, causal) in [(16, 577, 16, 16, 64, False), (1, 1776, 32, 8, 128, True)]:
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    out = torch.empty(B * S, Hq * D, device="cuda", dtype=torch.bfloat16)
    for rep in range(3):
        a = AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(), ldq=q.stride(0), ldk=k.stride(0),
                     ldv=v.stride(0), ldo=out.stride(0), B=B, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=int(causal),
                     scale=1 / math.sqrt(D), reserved=777)
        assert lib.vl2_attention(C.byref(a), torch.cuda.current_stream().cuda_stream) == 0
    buf = (C.c_longlong * 16)()
    assert lib.vl2_debug_attn_trace(buf) == 0
    n = buf[6]
    names = ["wait S", "TMEM ld", "mask+max+xchg", "wait PV/rescale", "exp+pack+sts", "fence+arrive"]
    print(f"D={D} causal={causal} tiles={n}: " + ", ".join(f"{nm}={buf[i] / max(n, 1):.0f}" for i, nm in enumerate(names)),
          f"| total/iter={sum(buf[:6]) / max(n, 1):.0f} cycles")
