"""CUDA-event timing of vl2_gemm_bf16 at the ViT / connector / decoder shapes (rotating operands > L2).
Usage: python scripts/microbench_gemm.py [--only NAME] [--bn BN]   (BN: 64..256 single-CTA, 1128..1256 pair)"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import ops

dev = torch.device("cuda:0")
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
bn = int(sys.argv[sys.argv.index("--bn") + 1]) if "--bn" in sys.argv else 0
NO_RES, NO_BIAS, HOT = "--no-res" in sys.argv, "--no-bias" in sys.argv, "--hot" in sys.argv   # epilogue ablations
SHAPES = [  # name, M, N, K, bias, residual, act
    ("siglip_qkv", 11664, 3456, 1152, True, False, 0),
    ("siglip_out", 11664, 1152, 1152, True, True, 0),
    ("siglip_fc1", 11664, 4304, 1152, True, False, 5),
    ("siglip_fc2", 11664, 1152, 4304, True, True, 0),
    ("clip_qkv", 9232, 3072, 1024, True, False, 0),
    ("clip_out", 9232, 1024, 1024, True, True, 0),
    ("clip_fc1", 9232, 4096, 1024, True, False, 1),
    ("clip_fc2", 9232, 1024, 4096, True, True, 0),
    ("llm_qkv", 1776, 6144, 4096, False, False, 0),
    ("llm_o", 1776, 4096, 4096, False, True, 0),
    ("llm_gate_up", 1776, 28672, 4096, False, False, 4),
    ("llm_down", 1776, 4096, 14336, False, True, 0),
]
res = {}
for name, M, N, K, has_bias, has_res, act in SHAPES:
    if only and name != only:
        continue
    has_res = has_res and not NO_RES
    has_bias = has_bias and not NO_BIAS
    pool = 1 if HOT else 4
    A = [torch.randn((M, K), device=dev, dtype=torch.bfloat16) for _ in range(pool)]
    W = [torch.randn((N, K), device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(pool)]
    n_out = N // 2 if act == 4 else N
    R = [torch.randn((M, n_out), device=dev, dtype=torch.bfloat16) for _ in range(pool)] if has_res else None
    b = torch.randn((N,), device=dev) if has_bias else None
    out = torch.empty((M, n_out), device=dev, dtype=torch.bfloat16)

    def run(i):
        ops.gemm(A[i % pool], W[i % pool], bias=b, act=act, residual=R[i % pool] if R else None, out=out, bn=bn)

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    res[name] = {"M": M, "N": N, "K": K, "us": round(us, 1), "TFs": round(2 * M * N * K / us / 1e6, 1)}
    if "--trace" in sys.argv:
        ops.gemm(A[0], W[0], bias=b, act=act, residual=R[0] if R else None, out=out, bn=bn, trace=True)
        t = ops.gemm_trace()
        n_t = int(t[0])
        base = t[1]
        rows = []
        for i in range(n_t):
            v = t[8 * i + 1: 8 * i + 7]
            rows.append({"mma_wait_acc": v[1] - v[0], "mma_issue": v[2] - v[1], "epi_wait_acc": v[4] - v[3],
                         "epi_work": v[5] - v[4], "mma_start_at": v[0] - base, "epi_start_at": v[3] - base})
            ph = t[64 + 8 * i: 64 + 8 * i + 6]   # first span of epilogue warp 4: residual staged / chunk loaded / half 0 / half 1 / stored
            rows[-1]["span0"] = [ph[j + 1] - ph[j] for j in range(5)]
        res[name]["trace_cycles"] = rows
print(json.dumps(res))
