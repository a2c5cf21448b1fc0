"""Model registry + loader with the reference's names (videollama2/model/__init__.py:31-193)."""
from __future__ import annotations

import json
import os

import torch

from .config import Videollama2Config, VisionConfig
from .encoder import CLIPVisionTower, SiglipVisionTower, build_vision_tower
from .projector import STCConnector, STCConnectorV35, build_vision_projector, load_mm_projector
from .videollama2_mistral import Videollama2MistralConfig, Videollama2MistralForCausalLM
from .videollama2_qwen2 import Videollama2Qwen2Config, Videollama2Qwen2ForCausalLM

VLLMs = {
    "videollama2": Videollama2MistralForCausalLM,
    "videollama2_mistral": Videollama2MistralForCausalLM,
    "videollama2_qwen2": Videollama2Qwen2ForCausalLM,
}

VLLMConfigs = {
    "videollama2": Videollama2MistralConfig,
    "videollama2_mistral": Videollama2MistralConfig,
    "videollama2_qwen2": Videollama2Qwen2Config,
}


def _read_checkpoint(model_path: str):
    """HF checkpoint directory -> state dict (sharded safetensors or pytorch_model.bin)."""
    idx = os.path.join(model_path, "model.safetensors.index.json")
    files = []
    if os.path.exists(idx):
        with open(idx) as fh:
            files = sorted(set(json.load(fh)["weight_map"].values()))
    elif os.path.exists(os.path.join(model_path, "model.safetensors")):
        files = ["model.safetensors"]
    sd = {}
    if files:
        from safetensors.torch import load_file
        for f in files:
            sd.update(load_file(os.path.join(model_path, f)))
        return sd
    binf = os.path.join(model_path, "pytorch_model.bin")
    if os.path.exists(binf):
        return torch.load(binf, map_location="cpu")
    raise FileNotFoundError(f"no model.safetensors[.index.json] / pytorch_model.bin under {model_path}")


def load_pretrained_model(model_path, model_base=None, model_name=None, load_8bit=False, load_4bit=False,
                          device_map="auto", device="cuda", use_flash_attn=False, **kwargs):
    """SFT-checkpoint branch of the reference loader (model/__init__.py:165-193) for LOCAL directories:
    returns (tokenizer, model, image_processor, context_len).  LoRA / 4-bit / 8-bit branches are training artefacts
    outside the accelerated path."""
    if load_8bit or load_4bit:
        raise NotImplementedError("bitsandbytes loading is not supported by the B200 engine")
    if model_base is not None or (model_name and "lora" in model_name.lower()):
        raise NotImplementedError("LoRA / base+projector loading is not supported by the B200 engine")
    with open(os.path.join(model_path, "config.json")) as fh:
        raw = json.load(fh)
    model_type = raw.get("model_type", "videollama2_mistral")
    if model_type not in VLLMs:
        raise ValueError(f"unsupported model_type {model_type} (supported: {sorted(VLLMs)})")
    config = VLLMConfigs[model_type].from_dict(raw)
    config.model_type = model_type
    tower = config.mm_vision_tower
    if tower is not None and os.path.isdir(tower):
        config.vision_config = VisionConfig.from_dir(tower)
    elif "vision_config" in raw:
        config.vision_config = VisionConfig(**raw["vision_config"])
    model = VLLMs[model_type].from_state_dict(config, _read_checkpoint(model_path), device=device)
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(model_path, use_fast=True)
    processor = None
    if config.mm_vision_tower is not None:
        processor = model.get_vision_tower().image_processor
    context_len = getattr(config, "max_sequence_length", 2048)
    return tokenizer, model, processor, context_len
