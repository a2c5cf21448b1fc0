"""Shared test helpers: synth cfg -> engine config, engine construction."""
import torch


def engine_config(cfg):
    from videollama2_b200.model.config import Videollama2Config, VisionConfig
    l, v = cfg.llm, cfg.vision
    vc = VisionConfig(hidden_size=v.hidden, intermediate_size=v.inter, num_hidden_layers=v.layers,
                      num_attention_heads=v.heads, image_size=v.image, patch_size=v.patch, layer_norm_eps=v.eps,
                      hidden_act="gelu_pytorch_tanh" if v.kind == "siglip" else "quick_gelu",
                      model_type="siglip_vision_model" if v.kind == "siglip" else "clip_vision_model")
    return Videollama2Config(
        model_type="videollama2_qwen2" if l.kind == "qwen2" else "videollama2_mistral",
        hidden_size=l.hidden, intermediate_size=l.inter, num_hidden_layers=l.layers, num_attention_heads=l.heads,
        num_key_value_heads=l.kv_heads, vocab_size=l.vocab, rms_norm_eps=l.eps, rope_theta=l.theta,
        attention_bias=(l.kind == "qwen2"), mm_vision_tower="synthetic-siglip" if v.kind == "siglip" else "synthetic-clip", mm_projector_type=cfg.projector,
        mm_hidden_size=v.hidden, mm_vision_select_layer=cfg.select_layer, num_frames=cfg.frames, vision_config=vc)


def build_engine(cfg, sd, device="cuda"):
    from videollama2_b200.model import VLLMs
    ec = engine_config(cfg)
    return VLLMs[ec.model_type].from_state_dict(ec, sd, device=device)


def rel(a, b):
    a = a.float().cpu()
    b = b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
