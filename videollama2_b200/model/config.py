"""Configuration objects carrying the same attribute names the reference reads from its HF configs
(videollama2_arch.py:49-68: mm_vision_tower, mm_projector_type, mm_hidden_size, mm_vision_select_layer,
mm_vision_select_feature, num_frames; HF Mistral/Qwen2 config fields; HF CLIPVisionConfig fields)."""
from __future__ import annotations

import dataclasses
import json
import os
from typing import Optional


@dataclasses.dataclass
class VisionConfig:
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    layer_norm_eps: float = 1e-5
    hidden_act: str = "quick_gelu"
    model_type: str = "clip_vision_model"

    @classmethod
    def from_dir(cls, path: str) -> "VisionConfig":
        with open(os.path.join(path, "config.json")) as fh:
            d = json.load(fh)
        d = d.get("vision_config", d)
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in d.items() if k in names})


@dataclasses.dataclass
class Videollama2Config:
    model_type: str = "videollama2_mistral"      # or videollama2_qwen2
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    vocab_size: int = 32000
    rms_norm_eps: float = 1e-5
    rope_theta: float = 1e6
    max_position_embeddings: int = 32768
    attention_bias: bool = False                 # True for qwen2 (q/k/v bias)
    # multimodal attributes (same names as the reference)
    mm_vision_tower: Optional[str] = None
    mm_projector_type: str = "stc_connector"
    mm_hidden_size: int = 1024
    mm_vision_select_layer: int = -2
    mm_vision_select_feature: str = "patch"
    num_frames: int = 8
    vision_config: Optional[VisionConfig] = None
    eos_token_id: Optional[int] = 2
    pad_token_id: Optional[int] = None

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @classmethod
    def from_dict(cls, d: dict) -> "Videollama2Config":
        names = {f.name for f in dataclasses.fields(cls)}
        kw = {k: v for k, v in d.items() if k in names and k != "vision_config"}
        cfg = cls(**kw)
        if "qwen2" in cfg.model_type:
            cfg.attention_bias = True
        return cfg
