"""Configs, deterministic synthetic weights (HF state-dict names of the reference model) and synthetic inputs.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Dimensions follow the upstream HF configs named in the reference README
(README.md:117-126): CLIP-ViT-L/14@336, Mistral-7B-Instruct-v0.2, Qwen2-7B-Instruct.

Every tensor is generated from its own torch CPU generator seeded with crc32(name) ^ SEED, so any subset (one layer,
one stage) can be regenerated independently and identically on any box with the same torch build."""
from __future__ import annotations

import dataclasses
import zlib
from typing import Dict, Iterator, List, Tuple

import torch

SEED = 20240603


@dataclasses.dataclass(frozen=True)
class VisionCfg:
    hidden: int = 1024
    inter: int = 4096
    layers: int = 24
    heads: int = 16
    image: int = 336
    patch: int = 14
    eps: float = 1e-5
    kind: str = "clip"             # "clip" (CLS + pre-LN, quick_gelu) | "siglip" (no CLS, patch bias, gelu-tanh)

    @property
    def seq(self) -> int:
        return self.num_patches + (1 if self.kind == "clip" else 0)

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def num_patches(self) -> int:
        return self.grid ** 2


@dataclasses.dataclass(frozen=True)
class LlmCfg:
    kind: str = "mistral"          # "mistral" | "qwen2"  (qwen2: q/k/v bias)
    hidden: int = 4096
    inter: int = 14336
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    vocab: int = 32000
    eps: float = 1e-5
    theta: float = 1e6

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclasses.dataclass(frozen=True)
class ModelCfg:
    name: str
    vision: VisionCfg
    llm: LlmCfg
    frames: int
    prompt: int
    select_layer: int = -2
    projector: str = "stc_connector"
    stc_depth: int = 4

    @property
    def stc_pad(self) -> int:
        return 0 if self.projector == "stc_connector_v35" else 1

    @property
    def vis_tokens(self) -> int:
        p = self.stc_pad
        t = (self.frames + 2 * p - 2) // 2 + 1
        g = (self.vision.grid + 2 * p - 2) // 2 + 1
        return t * g * g

    @property
    def seq(self) -> int:
        return self.prompt - 1 + self.vis_tokens


CLIP_L_336 = VisionCfg()
# google/siglip-so400m-patch14-384 (README.md:125-126; dims from the upstream HF config)
SIGLIP_SO400M_384 = VisionCfg(hidden=1152, inter=4304, layers=27, heads=16, image=384, patch=14, eps=1e-6, kind="siglip")
MISTRAL_7B = LlmCfg()
QWEN2_7B = LlmCfg(kind="qwen2", hidden=3584, inter=18944, layers=28, heads=28, kv_heads=4, vocab=152064, eps=1e-6)

TINY_VIT = VisionCfg(hidden=128, inter=256, layers=4, heads=2, image=56, patch=14)
TINY_LLM = LlmCfg(hidden=256, inter=512, layers=2, heads=4, kv_heads=2, vocab=512)
TINY_QWEN = LlmCfg(kind="qwen2", hidden=256, inter=512, layers=2, heads=2, kv_heads=1, vocab=512, eps=1e-6)
# head_dim 72, odd 5x5 grid, intermediate size that is a multiple of 8 but not of 64: the so400m tower's awkward shapes
TINY_SIGLIP = VisionCfg(hidden=288, inter=304, layers=3, heads=4, image=70, patch=14, eps=1e-6, kind="siglip")
# "mid": real head dims / tile-tail shapes at a size the CPU oracle finishes in seconds
MID_VIT = VisionCfg(hidden=256, inter=512, layers=3, heads=4, image=112, patch=14)
MID_LLM = LlmCfg(hidden=512, inter=1024, layers=2, heads=4, kv_heads=2, vocab=1024)

CONFIGS: Dict[str, ModelCfg] = {
    "tiny": ModelCfg("tiny", TINY_VIT, TINY_LLM, frames=4, prompt=12),
    "tiny_qwen2": ModelCfg("tiny_qwen2", TINY_VIT, TINY_QWEN, frames=4, prompt=12),
    "tiny_v35": ModelCfg("tiny_v35", TINY_VIT, TINY_LLM, frames=4, prompt=12, projector="stc_connector_v35"),
    "tiny_siglip": ModelCfg("tiny_siglip", TINY_SIGLIP, TINY_QWEN, frames=4, prompt=12, projector="stc_connector_v35"),
    "mid": ModelCfg("mid", MID_VIT, MID_LLM, frames=6, prompt=40),
    "cfg1": ModelCfg("cfg1", CLIP_L_336, MISTRAL_7B, frames=8, prompt=32),
    "cfg2": ModelCfg("cfg2", CLIP_L_336, MISTRAL_7B, frames=16, prompt=256),
    "cfg3": ModelCfg("cfg3", CLIP_L_336, QWEN2_7B, frames=16, prompt=256),
    "cfg3_v21": ModelCfg("cfg3_v21", SIGLIP_SO400M_384, QWEN2_7B, frames=16, prompt=256, projector="stc_connector_v35"),
}


# ----------------------------------------------------------------------------------------------------------------
# parameter specs: (name, shape, kind) in HF state-dict naming (SURVEY.md §8b weight contract)
# ----------------------------------------------------------------------------------------------------------------
Spec = Tuple[str, Tuple[int, ...], str]


def vision_specs(v: VisionCfg, prefix: str = "model.vision_tower.vision_tower.vision_model.") -> List[Spec]:
    if v.kind == "siglip":
        return siglip_specs(v, prefix)
    s: List[Spec] = [
        (prefix + "embeddings.class_embedding", (v.hidden,), "emb"),
        (prefix + "embeddings.patch_embedding.weight", (v.hidden, 3, v.patch, v.patch), "w"),
        (prefix + "embeddings.position_embedding.weight", (v.num_patches + 1, v.hidden), "emb"),
        (prefix + "pre_layrnorm.weight", (v.hidden,), "gain"),
        (prefix + "pre_layrnorm.bias", (v.hidden,), "bias"),
    ]
    for i in range(v.layers):
        p = f"{prefix}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s += [(p + f"self_attn.{nm}.weight", (v.hidden, v.hidden), "w"), (p + f"self_attn.{nm}.bias", (v.hidden,), "bias")]
        s += [(p + "layer_norm1.weight", (v.hidden,), "gain"), (p + "layer_norm1.bias", (v.hidden,), "bias"),
              (p + "mlp.fc1.weight", (v.inter, v.hidden), "w"), (p + "mlp.fc1.bias", (v.inter,), "bias"),
              (p + "mlp.fc2.weight", (v.hidden, v.inter), "w"), (p + "mlp.fc2.bias", (v.hidden,), "bias"),
              (p + "layer_norm2.weight", (v.hidden,), "gain"), (p + "layer_norm2.bias", (v.hidden,), "bias")]
    s += [(prefix + "post_layernorm.weight", (v.hidden,), "gain"), (prefix + "post_layernorm.bias", (v.hidden,), "bias")]
    return s


def _encoder_layer_specs(v: VisionCfg, p: str) -> List[Spec]:
    s: List[Spec] = []
    for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
        s += [(p + f"self_attn.{nm}.weight", (v.hidden, v.hidden), "w"), (p + f"self_attn.{nm}.bias", (v.hidden,), "bias")]
    s += [(p + "layer_norm1.weight", (v.hidden,), "gain"), (p + "layer_norm1.bias", (v.hidden,), "bias"),
          (p + "mlp.fc1.weight", (v.inter, v.hidden), "w"), (p + "mlp.fc1.bias", (v.inter,), "bias"),
          (p + "mlp.fc2.weight", (v.hidden, v.inter), "w"), (p + "mlp.fc2.bias", (v.hidden,), "bias"),
          (p + "layer_norm2.weight", (v.hidden,), "gain"), (p + "layer_norm2.bias", (v.hidden,), "bias")]
    return s


def siglip_specs(v: VisionCfg, prefix: str) -> List[Spec]:
    """HF SiglipVisionModel state-dict names (encoder.py:84-101 wraps it as `vision_tower`).  The attention-pooling
    head and post_layernorm exist in real checkpoints but never feed hidden_states[-2]."""
    s: List[Spec] = [
        (prefix + "embeddings.patch_embedding.weight", (v.hidden, 3, v.patch, v.patch), "w"),
        (prefix + "embeddings.patch_embedding.bias", (v.hidden,), "bias"),
        (prefix + "embeddings.position_embedding.weight", (v.num_patches, v.hidden), "emb"),
    ]
    for i in range(v.layers):
        s += _encoder_layer_specs(v, f"{prefix}encoder.layers.{i}.")
    s += [(prefix + "post_layernorm.weight", (v.hidden,), "gain"), (prefix + "post_layernorm.bias", (v.hidden,), "bias"),
          (prefix + "head.probe", (1, 1, v.hidden), "emb"),
          (prefix + "head.attention.in_proj_weight", (3 * v.hidden, v.hidden), "w"),
          (prefix + "head.attention.in_proj_bias", (3 * v.hidden,), "bias"),
          (prefix + "head.attention.out_proj.weight", (v.hidden, v.hidden), "w"),
          (prefix + "head.attention.out_proj.bias", (v.hidden,), "bias"),
          (prefix + "head.layernorm.weight", (v.hidden,), "gain"), (prefix + "head.layernorm.bias", (v.hidden,), "bias"),
          (prefix + "head.mlp.fc1.weight", (v.inter, v.hidden), "w"), (prefix + "head.mlp.fc1.bias", (v.inter,), "bias"),
          (prefix + "head.mlp.fc2.weight", (v.hidden, v.inter), "w"), (prefix + "head.mlp.fc2.bias", (v.hidden,), "bias")]
    return s


def stc_specs(cin: int, c: int, depth: int = 4, prefix: str = "model.mm_projector.") -> List[Spec]:
    s: List[Spec] = []
    for stage, first_in in (("s1", cin), ("s2", c)):
        for b in range(1, depth + 1):
            bin_ = first_in if b == 1 else c
            p = f"{prefix}{stage}.b{b}."
            rd = int(round(bin_ * 0.25))
            s += [(p + "conv1.conv.weight", (c, bin_, 1, 1), "w"), (p + "conv1.bn.weight", (c,), "gain"), (p + "conv1.bn.bias", (c,), "bias"),
                  (p + "conv2.conv.weight", (c, 1, 3, 3), "w"), (p + "conv2.bn.weight", (c,), "gain"), (p + "conv2.bn.bias", (c,), "bias"),
                  (p + "se.fc1.weight", (rd, c, 1, 1), "w"), (p + "se.fc1.bias", (rd,), "bias"),
                  (p + "se.fc2.weight", (c, rd, 1, 1), "w"), (p + "se.fc2.bias", (c,), "bias"),
                  (p + "conv3.conv.weight", (c, c, 1, 1), "w"), (p + "conv3.bn.weight", (c,), "gain"), (p + "conv3.bn.bias", (c,), "bias")]
            if bin_ != c:
                s += [(p + "downsample.conv.weight", (c, bin_, 1, 1), "w"), (p + "downsample.bn.weight", (c,), "gain"),
                      (p + "downsample.bn.bias", (c,), "bias")]
    s += [(prefix + "sampler.0.weight", (c, c, 2, 2, 2), "w"), (prefix + "sampler.0.bias", (c,), "bias"),
          (prefix + "readout.0.weight", (c, c), "w"), (prefix + "readout.0.bias", (c,), "bias"),
          (prefix + "readout.2.weight", (c, c), "w"), (prefix + "readout.2.bias", (c,), "bias")]
    return s


def llm_layer_specs(l: LlmCfg, i: int) -> List[Spec]:
    p = f"model.layers.{i}."
    d = l.head_dim
    s: List[Spec] = []
    for nm, n in (("q_proj", l.heads * d), ("k_proj", l.kv_heads * d), ("v_proj", l.kv_heads * d)):
        s.append((p + f"self_attn.{nm}.weight", (n, l.hidden), "w"))
        if l.kind == "qwen2":
            s.append((p + f"self_attn.{nm}.bias", (n,), "bias"))
    s += [(p + "self_attn.o_proj.weight", (l.hidden, l.heads * d), "w"),
          (p + "mlp.gate_proj.weight", (l.inter, l.hidden), "w"), (p + "mlp.up_proj.weight", (l.inter, l.hidden), "w"),
          (p + "mlp.down_proj.weight", (l.hidden, l.inter), "w"),
          (p + "input_layernorm.weight", (l.hidden,), "gain"), (p + "post_attention_layernorm.weight", (l.hidden,), "gain")]
    return s


def llm_specs(l: LlmCfg) -> List[Spec]:
    s: List[Spec] = [("model.embed_tokens.weight", (l.vocab, l.hidden), "emb")]
    for i in range(l.layers):
        s += llm_layer_specs(l, i)
    s += [("model.norm.weight", (l.hidden,), "gain"), ("lm_head.weight", (l.vocab, l.hidden), "w")]
    return s


def model_specs(cfg: ModelCfg) -> List[Spec]:
    return llm_specs(cfg.llm) + vision_specs(cfg.vision) + stc_specs(cfg.vision.hidden, cfg.llm.hidden, cfg.stc_depth)


# ----------------------------------------------------------------------------------------------------------------
# deterministic fill
# ----------------------------------------------------------------------------------------------------------------
def make_tensor(name: str, shape: Tuple[int, ...], kind: str, dtype=torch.bfloat16) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(name.encode()) ^ SEED) & 0x7FFFFFFF)
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "w":
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        x *= fan_in ** -0.5
    elif kind == "gain":
        x = 1.0 + 0.1 * x
    elif kind == "bias":
        x *= 0.02
    elif kind == "emb":
        x *= 0.05
    else:
        raise ValueError(kind)
    return x.to(dtype)


def iter_state(specs: List[Spec], dtype=torch.bfloat16) -> Iterator[Tuple[str, torch.Tensor]]:
    for name, shape, kind in specs:
        yield name, make_tensor(name, shape, kind, dtype)


def state_dict(cfg: ModelCfg, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    return dict(iter_state(model_specs(cfg), dtype))


def inputs(cfg: ModelCfg) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pixels bf16 [T,3,H,W], input_ids int64 [1,P] with <video> = -201 at index 4)  — SURVEY.md §8d."""
    g = torch.Generator(device="cpu").manual_seed(1234)
    px = torch.randn((cfg.frames, 3, cfg.vision.image, cfg.vision.image), generator=g).to(torch.bfloat16)
    g2 = torch.Generator(device="cpu").manual_seed(1235)
    ids = torch.randint(3, cfg.llm.vocab, (1, cfg.prompt), generator=g2, dtype=torch.int64)
    ids[0, 4] = -201
    return px, ids
