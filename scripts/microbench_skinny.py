"""CUDA-event timing of vl2_gemm_skinny at the SE-MLP shapes of the STC connector's RegStage blocks (weights rotated
through a pool larger than L2, as in a step where 16 GB of decoder weights pass through between two uses)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import ops

dev = torch.device("cuda:0")
res = {}
for name, M, N, K, a_f32 in [("se_fc1_s1", 16, 1024, 4096, True), ("se_fc2_s1", 16, 4096, 1024, False),
                             ("se_fc1_s2", 9, 1024, 4096, True), ("se_fc1_first", 16, 256, 4096, True)]:
    pool = 24   # 24 x 8.4 MB > 126 MB of L2
    A = torch.randn((M, K), device=dev, dtype=torch.float32 if a_f32 else torch.bfloat16)
    W = [torch.randn((N, K), device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(pool)]
    b = torch.randn((N,), device=dev)
    for hot in (False, True):
        run = lambda i: ops.gemm_skinny(A, W[0 if hot else i % pool], bias=b, act=ops.ACT_SILU)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        n = 48
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / n * 1e3
        res[name + ("_hot" if hot else "_cold")] = {"us": round(us, 1), "GBps": round(2 * N * K / us / 1e3, 1)}
print(json.dumps(res))
