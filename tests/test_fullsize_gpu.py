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
    ss_x = ops.row_sumsq(xd)
    ss_h = torch.zeros_like(ss_x)
    out = eng._layer(eng.layers[0], xd, S, 0, None, ss_x, ss_h)
    assert out.shape == (S, l.hidden)
    assert rel(out, ref) < 1e-2
    # the folded-RMSNorm bookkeeping: ss_x now holds sum(out^2) per row, the scratch buffer is back to zero
    assert rel(ss_x, out.float().pow(2).sum(-1)) < 1e-4 and float(ss_h.abs().max()) == 0.0


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
