"""Import the UNMODIFIED reference classes from /root/reference on CPU (TEST INFRASTRUCTURE; build container only).

Three shims, none of which touches the reference's arithmetic:
  1. `timm` is absent            -> oracle/shims/timm (RegStage / LayerNorm2d restated from timm 1.0.3) on sys.path
  2. `transformers.TRANSFORMERS_CACHE` was removed in transformers 5.x (projector.py:24)  -> define the attribute
  3. encoder.py:24 forces flash_attention_2 (needs a GPU) and encoder.py:21-23 fetch the CLIP config from the hub
     -> a local "clip" directory with config.json / preprocessor_config.json, and a CLIPVisionModel subclass that
        switches the attention backend to sdpa/eager before construction.
`videollama2/__init__.py` imports decord/imageio (video decoding, out of scope): the package is registered as a bare
namespace pointing at the reference directory so that `videollama2.model` / `videollama2.mm_utils` import directly,
with stub `decord` / `imageio` modules.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import torch

REF_ROOT = os.environ.get("VL2_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "videollama2", "model"))


_loaded = None


def load():
    """Returns the reference `videollama2.model` module (with `mm_utils` importable as videollama2.mm_utils)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    shim_dir = os.path.join(_HERE, "shims")
    if shim_dir not in sys.path:
        sys.path.insert(0, shim_dir)
    import transformers
    if not hasattr(transformers, "TRANSFORMERS_CACHE"):
        transformers.TRANSFORMERS_CACHE = os.path.join(tempfile.gettempdir(), "hf_cache_unused")
    for name in ("decord", "imageio"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.VideoReader = object
            m.cpu = lambda *a, **k: None
            sys.modules[name] = m
    if "videollama2" not in sys.modules:
        pkg = types.ModuleType("videollama2")
        pkg.__path__ = [os.path.join(REF_ROOT, "videollama2")]
        sys.modules["videollama2"] = pkg
    import importlib
    enc = importlib.import_module("videollama2.model.encoder")

    class _CpuCLIPVisionModel(enc.CLIPVisionModel):
        def __init__(self, config=None, *a, **k):
            config._attn_implementation = os.environ.get("VL2_ORACLE_ATTN", "sdpa")
            super().__init__(config, *a, **k)

    enc.CLIPVisionModel = _CpuCLIPVisionModel

    class _CpuSiglipVisionModel(enc.SiglipVisionModel):     # encoder.py:96 forces flash_attention_2 the same way
        def __init__(self, config=None, *a, **k):
            config._attn_implementation = os.environ.get("VL2_ORACLE_ATTN", "sdpa")
            super().__init__(config, *a, **k)

    enc.SiglipVisionModel = _CpuSiglipVisionModel
    model = importlib.import_module("videollama2.model")
    _loaded = model
    return model


def _write_json_atomic(path: str, obj) -> None:
    """Concurrent test processes share these temp directories: never expose a half-written file."""
    tmp = f"{path}.{os.getpid()}.tmp"
    with open(tmp, "w") as fh:
        json.dump(obj, fh)
    os.replace(tmp, path)


def clip_dir(vcfg) -> str:
    """A local directory whose path contains 'clip' (encoder.py:157) holding the tower + processor configs."""
    d = os.path.join(tempfile.gettempdir(), f"vl2_oracle_clip_{vcfg.hidden}_{vcfg.layers}_{vcfg.image}")
    os.makedirs(d, exist_ok=True)
    _write_json_atomic(os.path.join(d, "config.json"), {"model_type": "clip_vision_model", "hidden_size": vcfg.hidden, "intermediate_size": vcfg.inter,
                   "num_hidden_layers": vcfg.layers, "num_attention_heads": vcfg.heads, "image_size": vcfg.image,
                   "patch_size": vcfg.patch, "hidden_act": "quick_gelu", "layer_norm_eps": vcfg.eps,
                   "projection_dim": 768, "num_channels": 3})
    _write_json_atomic(os.path.join(d, "preprocessor_config.json"), {"crop_size": vcfg.image, "do_center_crop": True, "do_normalize": True, "do_resize": True,
                   "feature_extractor_type": "CLIPFeatureExtractor", "image_processor_type": "CLIPImageProcessor",
                   "image_mean": [0.48145466, 0.4578275, 0.40821073], "image_std": [0.26862954, 0.26130258, 0.27577711],
                   "resample": 3, "size": vcfg.image})
    return d


def siglip_dir(vcfg) -> str:
    """A local directory whose path contains 'siglip' (encoder.py:159) holding the tower + processor configs."""
    d = os.path.join(tempfile.gettempdir(), f"vl2_oracle_siglip_{vcfg.hidden}_{vcfg.layers}_{vcfg.image}")
    os.makedirs(d, exist_ok=True)
    _write_json_atomic(os.path.join(d, "config.json"), {"model_type": "siglip_vision_model", "hidden_size": vcfg.hidden, "intermediate_size": vcfg.inter,
                   "num_hidden_layers": vcfg.layers, "num_attention_heads": vcfg.heads, "image_size": vcfg.image,
                   "patch_size": vcfg.patch, "hidden_act": "gelu_pytorch_tanh", "layer_norm_eps": vcfg.eps,
                   "num_channels": 3})
    _write_json_atomic(os.path.join(d, "preprocessor_config.json"), {"do_resize": True, "do_rescale": True, "do_normalize": True, "image_processor_type": "SiglipImageProcessor",
                   "image_mean": [0.5, 0.5, 0.5], "image_std": [0.5, 0.5, 0.5], "resample": 3, "rescale_factor": 1 / 255,
                   "size": {"height": vcfg.image, "width": vcfg.image}})
    return d


def build_reference_model(cfg, dtype=torch.float32, state=None):
    """Instantiate the reference's Videollama2{Mistral,Qwen2}ForCausalLM for `cfg` (oracle.synth.ModelCfg) on CPU and
    load the synthetic HF-named weights.  dtype float32 -> "G32" golden (fp32 math on bf16-rounded weights);
    dtype bfloat16 -> "Hbf16" (what the reference literally computes)."""
    from . import synth
    model_mod = load()
    l = cfg.llm
    common = dict(hidden_size=l.hidden, intermediate_size=l.inter, num_hidden_layers=l.layers,
                  num_attention_heads=l.heads, num_key_value_heads=l.kv_heads, vocab_size=l.vocab,
                  rms_norm_eps=l.eps, rope_theta=l.theta, max_position_embeddings=32768, tie_word_embeddings=False,
                  attn_implementation=os.environ.get("VL2_ORACLE_ATTN", "sdpa"))
    if l.kind == "qwen2":
        hf_cfg = model_mod.Videollama2Qwen2Config(**common, use_sliding_window=False)
        cls = model_mod.Videollama2Qwen2ForCausalLM
    else:
        hf_cfg = model_mod.Videollama2MistralConfig(**common, sliding_window=None)
        cls = model_mod.Videollama2MistralForCausalLM
    hf_cfg.mm_vision_tower = siglip_dir(cfg.vision) if cfg.vision.kind == "siglip" else clip_dir(cfg.vision)
    hf_cfg.mm_projector_type = cfg.projector
    hf_cfg.mm_hidden_size = cfg.vision.hidden
    hf_cfg.mm_vision_select_layer = cfg.select_layer
    hf_cfg.mm_vision_select_feature = "patch"
    hf_cfg.num_frames = cfg.frames
    model = cls(hf_cfg)
    sd = state if state is not None else synth.state_dict(cfg)
    missing, unexpected = model.load_state_dict({k: v.to(torch.float32) for k, v in sd.items()}, strict=False)
    real_missing = [k for k in missing if "rotary_emb" not in k and "position_ids" not in k]
    if real_missing or unexpected:
        raise RuntimeError(f"state-dict mismatch: missing={real_missing[:8]} unexpected={list(unexpected)[:8]}")
    model = model.to(dtype).eval()
    return model
