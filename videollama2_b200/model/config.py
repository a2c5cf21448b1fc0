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

    # defaults of the two tower families when a config.json omits a key (HF CLIPVisionConfig / SiglipVisionConfig defaults)
    _FAMILY_DEFAULTS = {
        "clip": dict(layer_norm_eps=1e-5, hidden_act="quick_gelu", model_type="clip_vision_model"),
        "siglip": dict(layer_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh", model_type="siglip_vision_model"),
    }
    # the two towers the reference's released checkpoints name by hub id (README.md:117-126); used when the id is not a
    # local directory (there is no hub access in this engine)
    _KNOWN_TOWERS = {
        "clip-vit-large-patch14-336": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24,
                                           num_attention_heads=16, image_size=336, patch_size=14),
        "siglip-so400m-patch14-384": dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                                          num_attention_heads=16, image_size=384, patch_size=14),
    }

    @classmethod
    def _family(cls, d: dict, hint: str = "") -> str:
        mt = str(d.get("model_type", "")).lower()
        return "siglip" if ("siglip" in mt or (not mt and "siglip" in hint.lower())) else "clip"

    @classmethod
    def from_dict(cls, d: dict, hint: str = "") -> "VisionConfig":
        """A (possibly partial) HF vision config dict -> VisionConfig; keys the file omits take the defaults of ITS
        family (SigLIP: eps 1e-6 / gelu_pytorch_tanh, CLIP: eps 1e-5 / quick_gelu), not CLIP's for both."""
        d = d.get("vision_config", d)
        names = {f.name for f in dataclasses.fields(cls)}
        kw = dict(cls._FAMILY_DEFAULTS[cls._family(d, hint)])
        kw.update({k: v for k, v in d.items() if k in names})
        return cls(**kw)

    @classmethod
    def from_dir(cls, path: str) -> "VisionConfig":
        """`path`: a local directory with config.json, or one of the hub ids the reference checkpoints carry in
        `mm_vision_tower` (resolved from the built-in table: no network)."""
        cfg_file = os.path.join(path, "config.json")
        if os.path.isfile(cfg_file):
            with open(cfg_file) as fh:
                return cls.from_dict(json.load(fh), hint=path)
        base = os.path.basename(os.path.normpath(path)).lower()
        for key, dims in cls._KNOWN_TOWERS.items():
            if key in base:
                return cls(**dims, **cls._FAMILY_DEFAULTS["siglip" if "siglip" in key else "clip"])
        raise FileNotFoundError(f"vision tower '{path}' is neither a local directory with config.json nor a known tower "
                                f"({', '.join(cls._KNOWN_TOWERS)})")


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
    # 16-bit storage type of weights / activations: "bfloat16" (default) or "float16" - the reference's own inference
    # dtype (videollama2/__init__.py:60, model/__init__.py:71); selects which build of the kernel library runs
    torch_dtype: str = "bfloat16"

    @property
    def storage_dtype(self):
        import torch
        name = str(self.torch_dtype).replace("torch.", "")
        if name in ("float16", "half", "fp16"):
            return torch.float16
        if name in ("bfloat16", "bf16"):
            return torch.bfloat16
        raise ValueError(f"torch_dtype {self.torch_dtype!r} is not a storage type of the engine (bfloat16 / float16)")

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
