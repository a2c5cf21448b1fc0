"""First-contact probe for the tcgen05 kernels on a real B200: each stage runs in its own subprocess with a timeout so
that a trap or a hang in one kernel neither poisons the CUDA context for the others nor burns the gpurun lease."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

STAGES = {
    "gemm_min": """
        import torch
        from videollama2_b200 import ops
        torch.manual_seed(0)
        for (M, N, K) in [(128, 256, 64), (128, 64, 64), (128, 128, 64), (128, 256, 128), (128, 256, 256), (256, 512, 512), (200, 264, 136)]:
            a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
            out = ops.gemm(a, w); torch.cuda.synchronize()
            ref = a.float() @ w.float().t()
            err = (out.float() - ref).abs()
            rel = (out.float() - ref).norm() / ref.norm()
            print(f"gemm {M}x{N}x{K}: rel={rel:.3e} maxabs={err.max():.3e} refmax={ref.abs().max():.3e}", flush=True)
            if rel > 1e-2:
                e = err.cpu()
                rb = e.view(-1, 8, N).amax(dim=(1, 2)) if M % 8 == 0 else None
                print("  per-8-row-block max err:", None if rb is None else [round(x, 2) for x in rb[:16].tolist()])
                cb = e[:, : (N // 8) * 8].reshape(M, -1, 8).amax(dim=(0, 2))
                print("  per-8-col-block max err:", [round(x, 2) for x in cb[:32].tolist()])
                # does a K-chunk subset explain the output?
                for kk in range(0, K, 16):
                    part = a[:, kk:kk + 16].float() @ w[:, kk:kk + 16].float().t()
                    print(f"  corr with k-chunk {kk}: {torch.nn.functional.cosine_similarity(out.float().flatten(), part.flatten(), dim=0):.3f}")
                print("  out[0,:8]", out[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
    """,
    "attn_min": """
        import math, torch
        from videollama2_b200 import ops
        torch.manual_seed(0)
        for (B, S, Hq, Hkv, D, causal) in [(1, 128, 1, 1, 64, False), (1, 128, 1, 1, 128, False), (1, 256, 2, 1, 128, True), (2, 577, 2, 2, 64, False), (1, 300, 4, 2, 128, True)]:
            qkv = torch.randn(B * S, (Hq + 2 * Hkv) * D, device='cuda').bfloat16()
            q = qkv[:, : Hq * D]; k = qkv[:, Hq * D:(Hq + Hkv) * D]; v = qkv[:, (Hq + Hkv) * D:]
            out = ops.attention(q, k, v, B=B, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=causal, scale=1 / math.sqrt(D)); torch.cuda.synchronize()
            qf = q.float().view(B, S, Hq, D).transpose(1, 2)
            kf = k.float().view(B, S, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, 1)
            vf = v.float().view(B, S, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, 1)
            s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
            if causal: s = s.masked_fill(torch.ones(S, S, device='cuda', dtype=torch.bool).triu(1), float('-inf'))
            ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * S, Hq * D)
            rel = (out.float() - ref).norm() / ref.norm()
            print(f"attn B{B} S{S} Hq{Hq} Hkv{Hkv} D{D} causal={causal}: rel={rel:.3e}", flush=True)
            if rel > 2e-2:
                # diagnose: is it the QK part or the PV part?  compare against softmax(S) @ V with V rows permuted hypotheses
                print("  out[0,:8]", out[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
                print("  out[S-1,:8]", out[S - 1, :8].tolist()); print("  ref[S-1,:8]", ref[S - 1, :8].tolist())
    """,
}


def run_stage(name, code, timeout):
    env = dict(os.environ, PYTHONPATH=ROOT)
    try:
        r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        out = r.stdout + r.stderr[-3000:]
        status = f"exit={r.returncode}"
    except subprocess.TimeoutExpired as e:
        out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        status = "TIMEOUT"
    msg = f"===== {name}: {status} =====\n{out}\n"
    print(msg, flush=True)
    with open(os.path.join(OUT, "probe.log"), "a") as fh:
        fh.write(msg)
    return status


if __name__ == "__main__":
    which = sys.argv[1:] or list(STAGES)
    for n in which:
        run_stage(n, STAGES[n], 180)
