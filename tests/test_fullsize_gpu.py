"""Parity at BASELINE.json's real dimensions (CLIP-L/14@336, Mistral-7B, STC C=4096), one layer / block at a time so the
CPU oracle finishes in seconds: each engine layer is fed the oracle's input (bf16-rounded) and compared with the fp32
golden of that layer (relL2 <= 1e-2; the reference-style bf16 run of the same layer sits at 3-6e-3)."""
import pytest
import torch

from helpers import rel

pytestmark = pytest.mark.gpu


def test_clip_layer_real_dims(cuda):
    from oracle import synth, torch_ref
    from videollama2_b200.model.config import VisionConfig
    from videollama2_b200.model.encoder import CLIPVisionTower
    v = synth.VisionCfg(layers=1)                       # real width/heads/patching, one encoder layer
    sd = dict(synth.iter_state(synth.vision_specs(v)))
    px = torch.randn((2, 3, 336, 336), generator=torch.Generator().manual_seed(7)).to(torch.bfloat16)
    hs = torch_ref.vit_hidden_states(sd, v, px, torch.float32)
    args = type("A", (), {"mm_vision_select_layer": 1, "mm_vision_select_feature": "cls_patch"})()
    vc = VisionConfig(num_hidden_layers=1)
    tower = CLIPVisionTower("synthetic-clip", args, vision_config=vc).load_state_dict(
        sd, cuda, prefix="model.vision_tower.vision_tower.vision_model.")
    out = tower(px.to(cuda))
    assert out.shape == (2, 577, 1024)
    assert rel(out, hs[1]) < 1e-2
    args0 = type("A", (), {"mm_vision_select_layer": 0, "mm_vision_select_feature": "patch"})()
    emb = CLIPVisionTower("synthetic-clip", args0, vision_config=vc).load_state_dict(
        sd, cuda, prefix="model.vision_tower.vision_tower.vision_model.")(px.to(cuda))
    assert rel(emb, hs[0][:, 1:]) < 5e-3               # patch conv + cls/pos + pre-LN only


def test_siglip_layer_real_dims(cuda):
    """SigLIP-so400m/14@384 at its real dimensions (1152 wide, 16 heads x 72, MLP 4304, 27x27 patches out of 384 pixels
    with 6 unused ones, patch bias, gelu-tanh): embedding and one encoder layer against the fp32 oracle."""
    from oracle import synth, torch_ref
    from videollama2_b200.model.config import VisionConfig
    from videollama2_b200.model.encoder import SiglipVisionTower
    v = synth.VisionCfg(hidden=1152, inter=4304, layers=1, heads=16, image=384, patch=14, eps=1e-6, kind="siglip")
    pfx = "model.vision_tower.vision_tower.vision_model."
    sd = dict(synth.iter_state(synth.vision_specs(v)))
    px = torch.randn((2, 3, 384, 384), generator=torch.Generator().manual_seed(8)).to(torch.bfloat16)
    hs = torch_ref.vit_hidden_states(sd, v, px, torch.float32)
    vc = VisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=1, num_attention_heads=16,
                      image_size=384, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh",
                      model_type="siglip_vision_model")
    args = type("A", (), {"mm_vision_select_layer": 1, "mm_vision_select_feature": "patch"})()
    tower = SiglipVisionTower("synthetic-siglip", args, vision_config=vc).load_state_dict(sd, cuda, prefix=pfx)
    out = tower(px.to(cuda))
    assert out.shape == (2, 729, 1152) and tower.num_patches == 729
    assert rel(out, hs[1]) < 1e-2
    args0 = type("A", (), {"mm_vision_select_layer": 0, "mm_vision_select_feature": "patch"})()
    emb = SiglipVisionTower("synthetic-siglip", args0, vision_config=vc).load_state_dict(sd, cuda, prefix=pfx)(px.to(cuda))
    assert rel(emb, hs[0]) < 5e-3                      # patch conv + bias + position rows (one GEMM)


def test_mistral_layer_real_dims(cuda):
    from oracle import synth, torch_ref
    from videollama2_b200.model.decoder import DecoderEngine
    from videollama2_b200 import presets as P
    l = synth.LlmCfg(layers=1)
    S = 1776
    sd = dict(synth.iter_state(synth.llm_layer_specs(l, 0)))
    x = (torch.randn((S, l.hidden), generator=torch.Generator().manual_seed(8)) * 0.7).to(torch.bfloat16)
    cos, sin = torch_ref.rope_cos_sin(S, l.head_dim, l.theta, torch.float32)
    ref = torch_ref.decoder_layer(sd, l, 0, x.float(), cos, sin, torch.float32)
    cfg = P.make_config(dict(P.MISTRAL_7B, num_hidden_layers=1), 16)
    eng = DecoderEngine(cfg)
    sd_full = dict(sd)
    sd_full["model.embed_tokens.weight"] = torch.zeros((8, l.hidden), dtype=torch.bfloat16)
    sd_full["model.norm.weight"] = torch.ones((l.hidden,), dtype=torch.bfloat16)
    sd_full["lm_head.weight"] = torch.zeros((8, l.hidden), dtype=torch.bfloat16)
    eng.load_state_dict(sd_full, cuda)
    from videollama2_b200 import ops
    xd = x.to(cuda)
    out, ss_out = eng._layer(eng.layers[0], xd, S, 0, None, ops.row_sumsq(xd))
    assert out.shape == (S, l.hidden)
    assert rel(out, ref) < 1e-2
    # the folded-RMSNorm bookkeeping: the down_proj epilogue left sum(out^2) per row as per-32-column partials
    assert ss_out.shape == (S, l.hidden // 32) and rel(ss_out.sum(-1), out.float().pow(2).sum(-1)) < 1e-4
    out2, _ = eng._layer(eng.layers[0], xd, S, 0, None, ops.row_sumsq(xd))
    assert torch.equal(out, out2)                      # no atomics anywhere: bit-reproducible


def test_stc_block_real_dims(cuda):
    from oracle import synth, torch_ref
    from videollama2_b200.model.projector import STCConnector
    C = 4096
    sd = dict(synth.iter_state(synth.stc_specs(1024, C, depth=2, prefix="")))      # two blocks per stage, real widths
    x = torch.randn((2, 24, 24, 1024), generator=torch.Generator().manual_seed(9)).to(torch.bfloat16)
    r1 = torch_ref.regstage_block(sd, "s1.b1.", x.float(), torch.float32, 1e-5)
    cfg = type("C", (), {"mm_hidden_size": 1024, "hidden_size": C})()
    stc = STCConnector(cfg, depth=2).load_state_dict(sd, cuda)
    out1 = stc._block(stc.blocks["s1"][0], x.to(cuda))                             # 1024 -> 4096 with conv shortcut
    assert rel(out1, r1) < 1e-2
    r1b = r1.to(torch.bfloat16)
    r2 = torch_ref.regstage_block(sd, "s1.b2.", r1b.float(), torch.float32, 1e-5)
    out2 = stc._block(stc.blocks["s1"][1], r1b.to(cuda))                           # identity shortcut
    assert rel(out2, r2) < 1e-2


def test_stc_connector_8192_wide(cuda):
    """The connector of BASELINE config 5 (VideoLLaMA2-72B: hidden 8192): one RegStage block per stage at the real width,
    Conv3d front end with K = 65536, depthwise kernel split over a 2-CTA cluster (LayerNorm statistics exchanged through
    DSMEM) - against the fp32 oracle."""
    from oracle import synth, torch_ref
    from videollama2_b200.model.projector import STCConnector
    C = 8192
    sd = dict(synth.iter_state(synth.stc_specs(1024, C, depth=1, prefix="")))
    x = torch.randn((1, 4, 36, 1024), generator=torch.Generator().manual_seed(19)).to(torch.bfloat16)      # T = 4, 6 x 6 patches
    ref = torch_ref.stc_stages({"model.mm_projector." + k: v for k, v in sd.items()}, x.float(), 1, 1, torch.float32)
    nb = torch_ref.stc_stages({"model.mm_projector." + k: v for k, v in sd.items()}, x, 1, 1, torch.bfloat16)
    cfg = type("C", (), {"mm_hidden_size": 1024, "hidden_size": C})()
    stc = STCConnector(cfg, depth=1).load_state_dict(sd, cuda)
    out = stc(x.to(cuda))
    assert out.shape == ref["out"].shape
    assert rel(out, ref["out"]) < max(1e-2, 1.25 * rel(nb["out"], ref["out"]))
    s1 = stc.run_s1(x[0].to(cuda).view(4, 6, 6, 1024))
    assert rel(s1, ref["s1"]) < max(1e-2, 1.25 * rel(nb["s1"], ref["s1"]))


# ---------------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE config-2 sizes (no CPU oracle needed: bit-exact identities)
# ---------------------------------------------------------------------------------------------------------------
def _rnd(shape, dev, scale, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def test_gemm_tile_variants_agree_bit_exactly_at_full_size(cuda):
    """Every tile configuration (single-CTA widths, cta_group::2 pairs) accumulates K in the same order, so the gate/up
    GEMM of config 2 (M=1776, N=28672, K=4096, SwiGLU) must be bit-identical across them (split-K tail off: it changes
    the summation order of the tiles it cuts); with the split-K tail the result moves only by fp32 summation order; one
    of them is checked against an fp32 matmul on a row sample."""
    from videollama2_b200 import ops
    M, N, K = 1776, 28672, 4096
    a = _rnd((M, K), cuda, 1.0, 100)
    w = _rnd((N, K), cuda, 0.02, 101)
    base = ops.gemm(a, w, act=ops.ACT_SWIGLU, splitk=False)
    for bn in (256, 128, 1256, 1224, 1128):
        assert torch.equal(ops.gemm(a, w, act=ops.ACT_SWIGLU, bn=bn, splitk=False), base), bn
    split = ops.gemm(a, w, act=ops.ACT_SWIGLU, splitk=3)       # forced K-slices for the 44 tiles of the last round
    assert rel(split, base) < 1e-3 and torch.equal(split, ops.gemm(a, w, act=ops.ACT_SWIGLU, splitk=3))
    rows = torch.tensor([0, 1, 127, 128, 1000, 1775], device=cuda)
    acc = a[rows].float() @ w.float().t()
    ref = torch.nn.functional.silu(acc[:, 0::2]) * acc[:, 1::2]
    assert rel(base[rows], ref) < 8e-3


def test_causal_attention_prefix_property_at_full_size(cuda):
    """Causal attention rows [0, 1024) over S=1776 equal the S=1024 problem on the same prefix, bit for bit (same key
    tiles, same order); and no output row depends on later tokens."""
    from videollama2_b200 import ops
    S, Hq, Hkv, D = 1776, 32, 8, 128
    qkv = _rnd((S, (Hq + 2 * Hkv) * D), cuda, 1.0, 102)
    q, k, v = qkv[:, : Hq * D], qkv[:, Hq * D: (Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
    full = ops.attention(q, k, v, B=1, S=S, Hq=Hq, Hkv=Hkv, D=D, causal=True, scale=D ** -0.5)
    S1 = 1024
    pre = ops.attention(qkv[:S1, : Hq * D], qkv[:S1, Hq * D: (Hq + Hkv) * D], qkv[:S1, (Hq + Hkv) * D:], B=1, S=S1, Hq=Hq,
                        Hkv=Hkv, D=D, causal=True, scale=D ** -0.5)
    assert torch.equal(full[:S1], pre)
    qkv2 = qkv.clone()
    qkv2[1500:] = _rnd((S - 1500, qkv.shape[1]), cuda, 3.0, 103)       # perturb the future
    out2 = ops.attention(qkv2[:, : Hq * D], qkv2[:, Hq * D: (Hq + Hkv) * D], qkv2[:, (Hq + Hkv) * D:], B=1, S=S, Hq=Hq,
                         Hkv=Hkv, D=D, causal=True, scale=D ** -0.5)
    assert torch.equal(out2[:1500], full[:1500])


def test_vit_frames_are_independent_at_full_size(cuda):
    """The tower treats frames as batch: encoding 16 frames at once or one by one gives the same bits per frame (this
    is what makes the frame-sharded multi-GPU path exact)."""
    from oracle import synth
    from videollama2_b200.model.config import VisionConfig
    from videollama2_b200.model.encoder import CLIPVisionTower
    v = synth.VisionCfg(layers=2)
    sd = dict(synth.iter_state(synth.vision_specs(v)))
    px = _rnd((16, 3, 336, 336), cuda, 1.0, 104)
    args = type("A", (), {"mm_vision_select_layer": -1, "mm_vision_select_feature": "patch"})()
    tower = CLIPVisionTower("synthetic-clip", args, vision_config=VisionConfig(num_hidden_layers=2)).load_state_dict(
        sd, cuda, prefix="model.vision_tower.vision_tower.vision_model.")
    allf = tower(px)
    assert allf.shape == (16, 576, 1024)
    for i in (0, 7, 15):
        assert torch.equal(tower(px[i:i + 1])[0], allf[i])
    assert torch.equal(tower(px[4:6]), allf[4:6])
