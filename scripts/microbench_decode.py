"""Per-kernel timing of the decode-step kernels (GEMV shapes of Mistral-7B, split-KV attention) with CUDA events.
Weights rotate through a pool larger than L2 so every launch streams from HBM.  Usage: python scripts/microbench_decode.py"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import ops

dev = torch.device("cuda:0")
res = {}


EAGER = "--eager" in sys.argv     # eager launches (for ncu); default: the n launches are captured in one CUDA graph so that
                                  # Python / driver launch overhead (~10 us per call) does not hide the small kernels


def timeit(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if EAGER:
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (3 * n) * 1e3


for name, N, K, act, rms in [("qkv", 6144, 4096, 0, 1e-5), ("wo", 4096, 4096, 0, 0.0), ("gate_up", 28672, 4096, ops.ACT_SWIGLU, 1e-5),
                             ("down", 4096, 14336, 0, 0.0), ("lm_head", 32000, 4096, 0, 1e-5)]:
    nbytes = N * K * 2
    pool = max(2, int(400e6 // nbytes) + 1)
    ws = [torch.randn((N, K), device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(pool)]
    x = torch.randn((1, K), device=dev, dtype=torch.bfloat16)
    out = torch.empty((1, N // 2 if act == ops.ACT_SWIGLU else N), device=dev, dtype=torch.bfloat16)
    us = timeit(lambda i: ops.gemv(x, ws[i % pool], act=act, rms_eps=rms, out=out), 40)
    res[name] = {"N": N, "K": K, "us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1)}
    del ws

Hq, Hkv, D = 32, 8, 128
for n_pos in (1777, 4096):
    width = (Hq + 2 * Hkv) * D
    caches = [torch.randn((n_pos, width), device=dev, dtype=torch.bfloat16) for _ in range(24)]
    q = torch.randn((Hq * D,), device=dev, dtype=torch.bfloat16)
    us = timeit(lambda i: ops.attention_decode(q, caches[i % 24][:, Hq * D:(Hq + Hkv) * D], caches[i % 24][:, (Hq + Hkv) * D:],
                                               n_pos=n_pos, Hq=Hq, Hkv=Hkv, D=D, scale=D ** -0.5), 48)
    res[f"attn_decode_{n_pos}"] = {"us": round(us, 2), "GBps": round(n_pos * 2 * Hkv * D * 2 / us / 1e3, 1)}
    del caches
print(json.dumps(res))
