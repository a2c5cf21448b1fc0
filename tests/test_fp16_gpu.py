"""The fp16 build of the kernel library (libvl2_f16.so: the same sources compiled with -DVL2_HALF; the reference's own
inference dtype, videollama2/__init__.py:60, model/__init__.py:71).  Kernel level: GEMM epilogues, attention, LayerNorm,
fused patch embedding, Conv3d front end, RoPE epilogue, GEMV on float16 tensors against fp32 PyTorch.  End to end: the tiny
configurations through the reference-shaped API with dtype=torch.float16 against the fp32 oracle, bar = max(1e-2, 1.25 x the
error of the reference-style fp16 CPU run).  fp16 has 3 more mantissa bits than bf16, so the bars are the bf16 ones or tighter."""
import math

import pytest
import torch

from helpers import engine_config, rel

pytestmark = pytest.mark.gpu
H = torch.float16


def rnd(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(H).to(dev)


def test_library_builds_are_distinct(cuda):
    from videollama2_b200 import _lib
    a, b = _lib.load(torch.bfloat16), _lib.load(torch.float16)
    assert a is not b and int(a.vl2_storage_dtype()) == 0 and int(b.vl2_storage_dtype()) == 1
    from videollama2_b200 import ops
    with pytest.raises(TypeError):
        ops.gemm(rnd((16, 64), cuda), rnd((16, 64), cuda).to(torch.bfloat16))       # mixed storage types in one call


@pytest.mark.parametrize("M,N,K,act,res", [(577, 1024, 1024, 1, True), (1776, 512, 4096, 0, True), (300, 264, 136, 5, False),
                                           (130, 256, 64, 3, False)])
def test_gemm_fp16(cuda, M, N, K, act, res):
    from videollama2_b200 import ops
    a = rnd((M, K), cuda, seed=1)
    w = rnd((N, K), cuda, K ** -0.5, seed=2)
    b = rnd((N,), cuda, 0.1, seed=3).float()
    r = rnd((M, N), cuda, seed=4) if res else None
    out = ops.gemm(a, w, bias=b, act=act, residual=r)
    assert out.dtype == H
    ref = a.float() @ w.float().t() + b
    ref = {0: lambda x: x, 1: lambda x: x * torch.sigmoid(1.702 * x), 3: torch.nn.functional.gelu,
           5: lambda x: torch.nn.functional.gelu(x, approximate="tanh")}[act](ref)
    if res:
        ref = ref + r.float()
    assert rel(out, ref) < 1.5e-3


def test_gemm_swiglu_rms_fp16(cuda):
    from videollama2_b200 import ops
    M, I, K = 260, 384, 512
    x = rnd((M, K), cuda, 2.0, seed=5)
    gate, up = rnd((I, K), cuda, 0.05, seed=6), rnd((I, K), cuda, 0.05, seed=7)
    w = torch.stack([gate, up], 1).reshape(2 * I, K).contiguous()
    out = ops.gemm(x, w, act=ops.ACT_SWIGLU, rms_in=ops.row_sumsq(x), rms_eps=1e-5)
    xn = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-5)
    ref = torch.nn.functional.silu(xn @ gate.float().t()) * (xn @ up.float().t())
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("B,S,Hq,Hkv,D,causal", [(2, 577, 4, 4, 64, False), (1, 1776, 8, 2, 128, True), (2, 130, 4, 4, 72, False)])
def test_attention_fp16(cuda, B, S, Hq, Hkv, D, causal):
    from videollama2_b200 import ops
    qkv = rnd((B * S, (Hq + 2 * Hkv) * D), cuda, 1.0, seed=9)
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    out = ops.attention(q, k, v, B=B, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=causal, scale=D ** -0.5)
    qf = q.float().view(B, S, Hq, D).transpose(1, 2)
    kf = k.float().view(B, S, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    vf = v.float().view(B, S, Hkv, D).transpose(1, 2).repeat_interleave(Hq // Hkv, dim=1)
    s = (qf @ kf.transpose(-1, -2)) * D ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(S, S, device=cuda, dtype=torch.bool).triu(1), float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B * S, Hq * D)
    assert out.dtype == H and rel(out, ref) < 2e-3


def test_row_kernels_and_gemv_fp16(cuda):
    from videollama2_b200 import ops
    x = rnd((1521, 4096), cuda, 2.0, seed=10) + 0.5
    g, b = rnd((4096,), cuda, 0.1, seed=11) + 1.0, rnd((4096,), cuda, 0.1, seed=12)
    y = ops.layernorm(x, g, b, 1e-5, act=ops.ACT_SILU)
    ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(x.float(), (4096,), g.float(), b.float(), 1e-5))
    assert rel(y, ref) < 1e-3
    w = rnd((1000, 4096), cuda, 4096 ** -0.5, seed=13)
    v = rnd((1, 4096), cuda, seed=14)
    out = ops.gemv(v, w, rms_eps=1e-5, out_dtype=torch.float32)
    vn = v.float() * torch.rsqrt(v.float().pow(2).mean() + 1e-5)
    assert rel(out, vn @ w.float().t()) < 1e-3


def test_patch_embed_conv3d_rope_fp16(cuda):
    from videollama2_b200 import ops
    F_, Hh, P, C = 2, 336, 14, 1024
    px = rnd((F_, 3, Hh, Hh), cuda, seed=16)
    wconv = rnd((C, 3, P, P), cuda, 0.05, seed=17)
    K = 3 * P * P
    wpad = torch.zeros((C, 640), device=cuda, dtype=H)
    wpad[:, :K] = wconv.reshape(C, K)
    np_ = (Hh // P) ** 2
    cls, pos = rnd((C,), cuda, seed=18), rnd((np_ + 1, C), cuda, seed=19)
    g, b = rnd((C,), cuda, 0.1, seed=20) + 1, rnd((C,), cuda, 0.1, seed=21)
    tok = ops.patch_embed(px, wpad, pos, P, cls=cls, gamma=g, beta=b, eps=1e-5)
    patch = torch.nn.functional.conv2d(px.float(), wconv.float(), stride=P).flatten(2).transpose(1, 2)
    emb = torch.cat([cls.float().expand(F_, 1, C), patch], 1) + pos.float()
    ref = torch.nn.functional.layer_norm(emb, (C,), g.float(), b.float(), 1e-5).reshape(-1, C)
    assert tok.dtype == H and rel(tok, ref) < 1.5e-3
    # Conv3d front end
    T, Hs, Cc, N = 4, 6, 64, 64
    x = rnd((T, Hs, Hs, Cc), cuda, seed=81)
    wt = rnd((N, Cc, 2, 2, 2), cuda, (8 * Cc) ** -0.5, seed=82)
    wk = wt.permute(0, 2, 3, 4, 1).reshape(N, 8 * Cc).contiguous()
    out = ops.conv3d_k2s2(x, wk, act=ops.ACT_SILU, pad=1)
    refc = torch.nn.functional.silu(torch.nn.functional.conv3d(x.float().permute(3, 0, 1, 2)[None], wt.float(), stride=2, padding=1))
    assert rel(out, refc[0].permute(1, 2, 3, 0).reshape(-1, N)) < 1.5e-3
    # RoPE epilogue (table packed with fp16 cos / sin)
    S, Hq, Hkv, D, Kd = 300, 4, 2, 128, 256
    nqk = (Hq + Hkv) * D
    xs = rnd((S, Kd), cuda, seed=91)
    w = rnd((nqk + Hkv * D, Kd), cuda, Kd ** -0.5, seed=92)
    perm = torch.cat([ops.rope_interleave_rows(Hq + Hkv, D), torch.arange(nqk, nqk + Hkv * D)]).to(cuda)
    tab = ops.rope_table(S + 8, D, 1e6, cuda, H)
    o = ops.gemm(xs, w[perm].contiguous(), rope=(tab, 5, D, nqk))
    y = xs.float() @ w.float().t()
    inv = 1.0 / (1e6 ** (torch.arange(0, D, 2, device=cuda).float() / D))
    ang = torch.outer(torch.arange(5, 5 + S, device=cuda).float(), inv)
    cos, sin = torch.cat([ang, ang], -1).cos().to(H).float(), torch.cat([ang, ang], -1).sin().to(H).float()
    qk = y[:, :nqk].view(S, Hq + Hkv, D)
    rot = torch.cat([-qk[..., D // 2:], qk[..., :D // 2]], -1)
    refr = torch.cat([(qk * cos[:, None] + rot * sin[:, None]).reshape(S, nqk), y[:, nqk:]], 1)[:, perm]
    assert rel(o, refr) < 1.5e-3


@pytest.mark.parametrize("name", ["tiny", "tiny_qwen2", "tiny_siglip", "mid"])
def test_end_to_end_fp16(cuda, name):
    """dtype=torch.float16 through from_state_dict -> forward / generate: logits against the fp32 oracle, greedy tokens equal
    the bf16 engine's unless the top-2 gap is inside the noise; KV-cache graph decode == eager decode."""
    from oracle import synth, torch_ref
    from videollama2_b200.model import VLLMs
    cfg = synth.CONFIGS[name]
    sd = synth.state_dict(cfg)
    px, ids = synth.inputs(cfg)
    gold = torch_ref.full_forward(sd, cfg, px, ids, torch.float32)
    noise = torch_ref.full_forward(sd, cfg, px, ids, torch.float16)           # the reference's dtype, reference-style on CPU
    ec = engine_config(cfg)
    model = VLLMs[ec.model_type].from_state_dict(ec, sd, device=cuda, dtype=torch.float16)
    assert model.dtype == H and model.get_vision_tower().dtype == H
    images = [(px.to(cuda).to(H), "video")]
    out = model(input_ids=ids, attention_mask=torch.ones_like(ids), images=images)
    err = rel(out.logits[0], gold["logits"])
    bar = max(1e-2, 1.25 * rel(noise["logits"], gold["logits"]))
    assert err < bar, (err, bar)
    mm = model.encode_images_or_videos(images)
    assert mm.dtype == H and rel(mm[0], gold["mm"]) < max(1e-2, 1.25 * rel(noise["mm"], gold["mm"]))
    eager = model.generate(ids, images=images, max_new_tokens=4, do_sample=False)
    model.enable_cuda_graphs(True)
    graphed = model.generate(ids, images=images, max_new_tokens=4, do_sample=False)
    model.enable_cuda_graphs(False)
    assert torch.equal(eager, graphed) and int(eager[0, 0]) == int(out.logits[0, -1].argmax())
