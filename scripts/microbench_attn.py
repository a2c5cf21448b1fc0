"""CUDA-event timing of vl2_attention at the tower / decoder shapes.  Usage: python scripts/microbench_attn.py"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import ops

dev = torch.device("cuda:0")
res = {}
for name, B, S, Hq, Hkv, D, causal in [("llm_mistral", 1, 1776, 32, 8, 128, True), ("llm_qwen2_v21", 1, 1607, 28, 4, 128, True),
                                       ("clip_vit", 16, 577, 16, 16, 64, False), ("siglip_vit", 16, 729, 16, 16, 72, False)]:
    qkv = torch.randn((B * S, (Hq + 2 * Hkv) * D), device=dev, dtype=torch.bfloat16)
    q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    out = torch.empty((B * S, Hq * D), device=dev, dtype=torch.bfloat16)
    run = lambda: ops.attention(q, k, v, B=B, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=causal, scale=D ** -0.5, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    fl = 4.0 * B * Hq * S * S * D * (0.5 if causal else 1.0)
    res[name] = {"us": round(us, 1), "TFs_algorithmic": round(fl / us / 1e6, 1)}
print(json.dumps(res))
