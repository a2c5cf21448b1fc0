"""What one rank of the frame-parallel vision stage runs, measured on ONE GPU: the CLIP tower on 2 / 4 / 8 / 16 frames as a
CUDA-graph replay (LayerNorm folding on / off, PDL on / off), and a tile-width sweep of vl2_gemm_bf16 at the 2-frame shapes
(M = 1154) where every GEMM is a partial wave.   python scripts/vit_shard_probe.py [--sweep]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videollama2_b200 import ops, presets
from videollama2_b200.model.encoder import CLIPVisionTower

dev = torch.device("cuda")
cfg = presets.make_config(presets.MISTRAL_7B, 16)
sd = {k: v for k, v in presets.random_state_dict(cfg, dev).items() if "vision_tower" in k}
PFX = "model.vision_tower.vision_tower.vision_model."


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"tower_ms": {}}
for fold in (True, False):
    CLIPVisionTower.fold_layernorm = fold
    tower = CLIPVisionTower("synthetic-clip", cfg, vision_config=cfg.vision_config).load_state_dict(sd, dev, prefix=PFX)
    for pdl in (False, True):
        with ops.pdl(pdl):
            tower.enable_cuda_graphs(True)
            for frames in (2, 4, 8, 16):
                px = torch.randn((frames, 3, 336, 336), device=dev).bfloat16()
                out["tower_ms"][f"fold={int(fold)} pdl={int(pdl)} frames={frames}"] = round(timed(lambda: tower(px)), 4)
            tower.enable_cuda_graphs(False)
    del tower
CLIPVisionTower.fold_layernorm = True

if "--sweep" in sys.argv:
    sweep = {}
    shapes = [("qkv", 1154, 3072, 1024, True, False, 0), ("out", 1154, 1024, 1024, True, True, 0),
              ("fc1", 1154, 4096, 1024, True, False, 1), ("fc2", 1154, 1024, 4096, True, True, 0),
              ("s1_1x1", 1152, 4096, 4096, False, False, 0), ("s1_in", 1152, 4096, 1024, False, False, 0),
              ("stc_1521", 1521, 4096, 4096, False, False, 0), ("conv3d_1521", 1521, 4096, 32768, True, False, 2)]
    for name, M, N, K, has_b, has_r, act in shapes:
        A = [torch.randn((M, K), device=dev, dtype=torch.bfloat16) for _ in range(3)]
        W = [torch.randn((N, K), device=dev, dtype=torch.bfloat16) * K ** -0.5 for _ in range(3)]
        R = [torch.randn((M, N), device=dev, dtype=torch.bfloat16) for _ in range(3)] if has_r else None
        b = torch.randn((N,), device=dev) if has_b else None
        o = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        row = {}
        for bn in (0, 64, 96, 128, 160, 192, 224, 256, 1128, 1160, 1192, 1224, 1256):
            if bn not in (0,) and (bn % 1000) - 32 >= N:
                continue
            cnt = [0]

            def run():
                i = cnt[0] % 3
                cnt[0] += 1
                ops.gemm(A[i], W[i], bias=b, act=act, residual=R[i] if R else None, out=o, bn=bn)
            g = torch.cuda.CUDAGraph()
            run(); run()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(6):
                    run()
            us = timed(g.replay, 10) / 6 * 1e3
            row[str(bn)] = round(us, 2)
        sweep[f"{name} {M}x{N}x{K}"] = row
    out["gemm_us_by_bn"] = sweep

    # attention at the shard shapes
    att = {}
    for frames in (2, 4, 16):
        S, H, D = 577, 16, 64
        qkv = torch.randn((frames * S, 3 * H * D), device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.attention(qkv[:, :1024], qkv[:, 1024:2048], qkv[:, 2048:], B=frames, S=S, Hq=H, Hkv=H, D=D,
                                   causal=False, scale=D ** -0.5)
        g = torch.cuda.CUDAGraph()
        fn(); fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(6):
                fn()
        att[f"frames={frames}"] = round(timed(g.replay, 10) / 6 * 1e3, 2)
    out["attention_us"] = att
if "--conv" in sys.argv:
    # Conv3d sampler of the connector: implicit GEMM (TMA gathers the taps) against explicit tap gather + GEMM
    T, H, W, C = 16, 24, 24, 4096
    x = torch.randn((T, H, W, C), device=dev, dtype=torch.bfloat16)
    wk = torch.randn((C, 8 * C), device=dev, dtype=torch.bfloat16) * (8 * C) ** -0.5
    b = torch.randn((C,), device=dev)
    conv = {}
    for bn in (0, 256, 1256, 1224, 1192):
        conv[f"implicit bn={bn}"] = round(timed(lambda: ops.conv3d_k2s2(x, wk, bias=b, act=ops.ACT_SILU, pad=1, bn=bn), 10) * 1e3, 1)
    conv["explicit im2col+gemm"] = round(timed(lambda: ops.gemm(ops.conv3d_im2col(x, 1), wk, bias=b, act=ops.ACT_SILU), 10) * 1e3, 1)
    out["conv3d_us"] = conv
if "--embed" in sys.argv:
    # ViT embeddings: the fused implicit-GEMM kernel against im2col + GEMM + finish (graph replays, per launch of the chain)
    emb = {}
    tw = CLIPVisionTower("synthetic-clip", cfg, vision_config=cfg.vision_config).load_state_dict(sd, dev, prefix=PFX)
    for frames in (2, 16):
        px = torch.randn((frames, 3, 336, 336), device=dev).bfloat16()
        w = tw.w
        fused = lambda: ops.patch_embed(px, w["patch"], w["pos"], 14, cls=w["cls"], gamma=w["pre_g"], beta=w["pre_b"], eps=1e-5)
        explicit = lambda: ops.clip_embed_finish(ops.gemm(ops.patch_im2col(px, 14, 640), w["patch"]), w["cls"], w["pos"],
                                                 w["pre_g"], w["pre_b"], frames, 1e-5)
        for name, fn in (("fused", fused), ("explicit", explicit)):
            g = torch.cuda.CUDAGraph()
            fn(); fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(4):
                    fn()
            emb[f"{name} frames={frames}"] = round(timed(g.replay, 10) / 4 * 1e3, 1)
    out["patch_embed_us"] = emb
print(json.dumps(out))
