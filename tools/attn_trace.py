"""Print the per-phase cycle trace of one softmax thread of the attention kernel (debug aid)."""
import ctypes as C
import math
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import _lib
from videollama2_b200._lib import AttnArgs

lib = _lib.load()
NAMES = ["wait S", "TMEM ld", "mask+max+xchg", "wait PV/rescale", "exp+pack+sts", "fence+arrive"]
for (B, S, Hq, Hkv, D, causal) in [(16, 577, 16, 16, 64, False), (1, 1776, 32, 8, 128, True)]:
    qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    out = torch.empty(B * S, Hq * D, device="cuda", dtype=torch.bfloat16)
    for mode in (777, 778):     # 777: the first (cold) item of CTA 0; 778: every item of CTA 0 (persistent kernel)
        for rep in range(3):
            a = AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(), ldq=q.stride(0),
                         ldk=k.stride(0), ldv=v.stride(0), ldo=out.stride(0), B=B, S=S, Hq=Hq, Hkv=Hkv, D=D,
                         causal=int(causal), scale=1 / math.sqrt(D), reserved=mode)
            assert lib.vl2_attention(C.byref(a), torch.cuda.current_stream().cuda_stream) == 0
        buf = (C.c_longlong * 16)()
        assert lib.vl2_debug_attn_trace(buf) == 0
        n, items = max(buf[7], 1), max(buf[8], 1)
        print(f"D={D} causal={causal} mode={mode} items={items} tiles={n}: "
              + ", ".join(f"{nm}={buf[i] / n:.0f}" for i, nm in enumerate(NAMES)),
              f"| per tile={sum(buf[:6]) / n:.0f} cycles; item epilogue={buf[6] / items:.0f} per item; "
              f"traced={sum(buf[:7])} of {buf[9]} cycles in the CTA's item loop; repeated (satisfied) S poll={buf[10] / n:.0f}")
    # timeline of CTA 0's second work item (reserved == 779): who waits for whom, in cycles relative to the item's first stamp
    a = AttnArgs(q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(), ldq=q.stride(0), ldk=k.stride(0),
                 ldv=v.stride(0), ldo=out.stride(0), B=B, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=int(causal),
                 scale=1 / math.sqrt(D), reserved=779)
    assert lib.vl2_attention(C.byref(a), torch.cuda.current_stream().cuda_stream) == 0
    tl = (C.c_longlong * 320)()
    assert lib.vl2_debug_attn_timeline(tl) == 0
    t = list(tl)
    n_kv = max(j for j in range(15) if t[8 * j] > 0) + 1 if any(t[8 * j] > 0 for j in range(15)) else 0
    base = min(x for x in (t[250], t[0], t[256], t[288]) if x > 0) if n_kv else 0
    rel = lambda x: (x - base) if x > 0 else -1
    print(f"  timeline D={D} causal={causal}: item of {n_kv} key tiles; cycles since the item's first event")
    print(f"  MMA: Q arrived {rel(t[250])}, QK(0) issued {rel(t[251])} (its K arrived {rel(t[241])})")
    for j in range(min(n_kv, 6)):
        sm = [rel(t[8 * j + i]) for i in range(7)]
        mm = [rel(t[128 + 8 * j + i]) for i in range(6)]
        print(f"  tile {j}: softmax start {sm[0]} S-arrived {sm[1]} S-in-regs {sm[2]} max-xchg {sm[3]} P/O-free {sm[4]} "
              f"P-stored {sm[5]} arrived {sm[6]} | MMA loop-top {mm[0]} K(j+1)-arrived {mm[1]} QK(j+1)-issued {mm[2]} "
              f"P(j)-arrived {mm[3]} V(j)-arrived {mm[4]} PV(j)-issued {mm[5]} | K(j) load issued {rel(t[256 + j])} "
              f"V(j) load issued {rel(t[288 + j])}")
    print(f"  epilogue: start {rel(t[120])} l-exchanged {rel(t[121])} last-PV-complete {rel(t[122])} O-stored+o_free {rel(t[123])}")
