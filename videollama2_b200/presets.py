"""Dimension presets of the models named in BASELINE.json (upstream HF configs named in the reference README:
openai/clip-vit-large-patch14-336, mistralai/Mistral-7B-Instruct-v0.2, Qwen/Qwen2-7B-Instruct) and the FLOP model of
the video->text prefill path used for every roofline fraction."""
from __future__ import annotations

from .model.config import Videollama2Config, VisionConfig

CLIP_L_336 = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                  image_size=336, patch_size=14, layer_norm_eps=1e-5)

# google/siglip-so400m-patch14-384: the tower of the released VideoLLaMA2.1 checkpoints (README.md:125-126)
SIGLIP_SO400M_384 = dict(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27, num_attention_heads=16,
                         image_size=384, patch_size=14, layer_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh",
                         model_type="siglip_vision_model")

MISTRAL_7B = dict(model_type="videollama2_mistral", hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                  num_attention_heads=32, num_key_value_heads=8, vocab_size=32000, rms_norm_eps=1e-5, rope_theta=1e6)
QWEN2_7B = dict(model_type="videollama2_qwen2", hidden_size=3584, intermediate_size=18944, num_hidden_layers=28,
                num_attention_heads=28, num_key_value_heads=4, vocab_size=152064, rms_norm_eps=1e-6, rope_theta=1e6,
                attention_bias=True)


# Qwen/Qwen2-72B-Instruct (BASELINE.json configs[4]: VideoLLaMA2-72B; SURVEY.md §8a row "HF Qwen2*")
QWEN2_72B = dict(model_type="videollama2_qwen2", hidden_size=8192, intermediate_size=29568, num_hidden_layers=80,
                 num_attention_heads=64, num_key_value_heads=8, vocab_size=152064, rms_norm_eps=1e-6, rope_theta=1e6,
                 attention_bias=True)


def random_tp_shard(cfg: Videollama2Config, rank: int, world: int, device, seed: int = 20240603):
    """This rank's slices of a random decoder checkpoint of `cfg` (HF names, shapes of tp_decoder.shard_state_dict), drawn on
    the device: the 72B model's 145 GB never exist in one place.  Replicated tensors (embedding, norms) use a rank-independent
    generator so that every rank holds the same values."""
    import torch
    from .model.tp_decoder import shard_plan
    plan = shard_plan(cfg, rank, world)
    H, D = cfg.hidden_size, plan["D"]
    g_loc = torch.Generator(device=device).manual_seed(seed + 1000 * (rank + 1))
    g_rep = torch.Generator(device=device).manual_seed(seed)

    def w(shape, fan_in, g):
        return (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * fan_in ** -0.5).to(torch.bfloat16)

    def gain():
        return (1.0 + 0.1 * torch.randn((H,), generator=g_rep, device=device, dtype=torch.float32)).to(torch.bfloat16)

    sd = {"model.embed_tokens.weight": (0.05 * torch.randn((cfg.vocab_size, H), generator=g_rep, device=device,
                                                            dtype=torch.float32)).to(torch.bfloat16)}
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"] = gain()
        sd[p + "post_attention_layernorm.weight"] = gain()
        for nm, n in (("q_proj", plan["Hq"] * D), ("k_proj", plan["Hkv"] * D), ("v_proj", plan["Hkv"] * D)):
            sd[p + f"self_attn.{nm}.weight"] = w((n, H), H, g_loc)
            if cfg.attention_bias:
                sd[p + f"self_attn.{nm}.bias"] = (0.02 * torch.randn((n,), generator=g_loc, device=device)).to(torch.bfloat16)
        sd[p + "self_attn.o_proj.weight"] = w((H, plan["Hq"] * D), cfg.num_attention_heads * D, g_loc)
        sd[p + "mlp.gate_proj.weight"] = w((plan["I"], H), H, g_loc)
        sd[p + "mlp.up_proj.weight"] = w((plan["I"], H), H, g_loc)
        sd[p + "mlp.down_proj.weight"] = w((H, plan["I"]), cfg.intermediate_size, g_loc)
    sd["model.norm.weight"] = gain()
    sd["lm_head.weight"] = w((plan["V"], H), H, g_loc)
    return sd


def random_vision_state(cfg: Videollama2Config, device, seed: int = 20240603):
    """Random vision tower + connector weights of `cfg` (HF names), the same on every rank (rank-independent generator)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed + 7)
    sd = {}
    for name, shape, kind in state_dict_specs(cfg):
        if not (name.startswith("model.vision_tower.") or name.startswith("model.mm_projector.")):
            continue
        x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            x *= fan_in ** -0.5
        elif kind == "gain":
            x = 1.0 + 0.1 * x
        elif kind == "bias":
            x *= 0.02
        else:
            x *= 0.05
        sd[name] = x.to(torch.bfloat16)
    return sd


def make_config(llm: dict, frames: int, projector: str = "stc_connector", vision: dict = CLIP_L_336) -> Videollama2Config:
    vc = VisionConfig(**vision)
    siglip = "siglip" in vc.model_type
    return Videollama2Config(**llm, mm_vision_tower="synthetic-siglip-so400m-patch14-384" if siglip
                             else "synthetic-clip-vit-large-patch14-336", mm_projector_type=projector,
                             mm_hidden_size=vc.hidden_size, mm_vision_select_layer=-2, num_frames=frames,
                             vision_config=vc)


def state_dict_specs(cfg: Videollama2Config):
    """(name, shape, kind) of every tensor in the reference checkpoint format."""
    v = cfg.vision_config
    H, I, D = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    s = [("model.embed_tokens.weight", (cfg.vocab_size, H), "emb")]
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}."
        for nm, n in (("q_proj", cfg.num_attention_heads * D), ("k_proj", cfg.num_key_value_heads * D),
                      ("v_proj", cfg.num_key_value_heads * D)):
            s.append((p + f"self_attn.{nm}.weight", (n, H), "w"))
            if cfg.attention_bias:
                s.append((p + f"self_attn.{nm}.bias", (n,), "bias"))
        s += [(p + "self_attn.o_proj.weight", (H, cfg.num_attention_heads * D), "w"),
              (p + "mlp.gate_proj.weight", (I, H), "w"), (p + "mlp.up_proj.weight", (I, H), "w"),
              (p + "mlp.down_proj.weight", (H, I), "w"),
              (p + "input_layernorm.weight", (H,), "gain"), (p + "post_attention_layernorm.weight", (H,), "gain")]
    s += [("model.norm.weight", (H,), "gain"), ("lm_head.weight", (cfg.vocab_size, H), "w")]
    vp = "model.vision_tower.vision_tower.vision_model."
    C, Iv = v.hidden_size, v.intermediate_size
    npatch = (v.image_size // v.patch_size) ** 2
    siglip = "siglip" in v.model_type
    if siglip:   # HF SiglipVisionModel naming; the pooling head (never on the path) is omitted from random init
        s += [(vp + "embeddings.patch_embedding.weight", (C, 3, v.patch_size, v.patch_size), "w"),
              (vp + "embeddings.patch_embedding.bias", (C,), "bias"),
              (vp + "embeddings.position_embedding.weight", (npatch, C), "emb")]
    else:
        s += [(vp + "embeddings.class_embedding", (C,), "emb"),
              (vp + "embeddings.patch_embedding.weight", (C, 3, v.patch_size, v.patch_size), "w"),
              (vp + "embeddings.position_embedding.weight", (npatch + 1, C), "emb"),
              (vp + "pre_layrnorm.weight", (C,), "gain"), (vp + "pre_layrnorm.bias", (C,), "bias")]
    for i in range(v.num_hidden_layers):
        p = f"{vp}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s += [(p + f"self_attn.{nm}.weight", (C, C), "w"), (p + f"self_attn.{nm}.bias", (C,), "bias")]
        s += [(p + "layer_norm1.weight", (C,), "gain"), (p + "layer_norm1.bias", (C,), "bias"),
              (p + "mlp.fc1.weight", (Iv, C), "w"), (p + "mlp.fc1.bias", (Iv,), "bias"),
              (p + "mlp.fc2.weight", (C, Iv), "w"), (p + "mlp.fc2.bias", (C,), "bias"),
              (p + "layer_norm2.weight", (C,), "gain"), (p + "layer_norm2.bias", (C,), "bias")]
    s += [(vp + "post_layernorm.weight", (C,), "gain"), (vp + "post_layernorm.bias", (C,), "bias")]
    pp = "model.mm_projector."
    for stage, first_in in (("s1", C), ("s2", H)):
        for b in range(1, 5):
            bi = first_in if b == 1 else H
            p = f"{pp}{stage}.b{b}."
            rd = int(round(bi * 0.25))
            s += [(p + "conv1.conv.weight", (H, bi, 1, 1), "w"), (p + "conv1.bn.weight", (H,), "gain"), (p + "conv1.bn.bias", (H,), "bias"),
                  (p + "conv2.conv.weight", (H, 1, 3, 3), "w"), (p + "conv2.bn.weight", (H,), "gain"), (p + "conv2.bn.bias", (H,), "bias"),
                  (p + "se.fc1.weight", (rd, H, 1, 1), "w"), (p + "se.fc1.bias", (rd,), "bias"),
                  (p + "se.fc2.weight", (H, rd, 1, 1), "w"), (p + "se.fc2.bias", (H,), "bias"),
                  (p + "conv3.conv.weight", (H, H, 1, 1), "w"), (p + "conv3.bn.weight", (H,), "gain"), (p + "conv3.bn.bias", (H,), "bias")]
            if bi != H:
                s += [(p + "downsample.conv.weight", (H, bi, 1, 1), "w"), (p + "downsample.bn.weight", (H,), "gain"),
                      (p + "downsample.bn.bias", (H,), "bias")]
    s += [(pp + "sampler.0.weight", (H, H, 2, 2, 2), "w"), (pp + "sampler.0.bias", (H,), "bias"),
          (pp + "readout.0.weight", (H, H), "w"), (pp + "readout.0.bias", (H,), "bias"),
          (pp + "readout.2.weight", (H, H), "w"), (pp + "readout.2.bias", (H,), "bias")]
    return s


def random_state_dict(cfg: Videollama2Config, device, seed: int = 20240603):
    """Random-init weights of the named architecture, generated on `device` (no checkpoints offline)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape, kind in state_dict_specs(cfg):
        x = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            x *= fan_in ** -0.5
        elif kind == "gain":
            x = 1.0 + 0.1 * x
        elif kind == "bias":
            x *= 0.02
        else:
            x *= 0.05
        sd[name] = x.to(torch.bfloat16)
    return sd


SYNTH_SEED = 20240603


def synth_tensor(name: str, shape, kind: str):
    """One tensor of the deterministic synthetic checkpoint: its own torch CPU generator seeded with
    crc32(name) ^ SYNTH_SEED, so any subset regenerates identically on any box with the same torch build.  Byte-identical
    to the weight factory the parity goldens were produced with (tests/test_host_logic.py pins the two together)."""
    import zlib

    import torch
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(name.encode()) ^ SYNTH_SEED) & 0x7FFFFFFF)
    x = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
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
    return x.to(torch.bfloat16)


def synthetic_state_dict(cfg: Videollama2Config, device, threads: int = 8):
    """The deterministic synthetic checkpoint of `cfg`, generated on the HOST RNG (so that it equals the weights the
    committed full-depth goldens were computed with) and moved to `device` tensor by tensor."""
    import concurrent.futures as cf

    import torch
    dev = torch.device(device)
    specs = state_dict_specs(cfg)

    def one(spec):
        name, shape, kind = spec
        return name, synth_tensor(name, shape, kind).to(dev, non_blocking=False)

    with cf.ThreadPoolExecutor(max(1, threads)) as ex:
        return dict(ex.map(one, specs))


def synthetic_inputs(cfg: Videollama2Config, frames: int, prompt: int):
    """(pixels bf16 [T,3,H,W], input_ids int64 [1,P] with <video> = -201 at index 4) — SURVEY.md §8d."""
    import torch
    v = cfg.vision_config
    g = torch.Generator(device="cpu").manual_seed(1234)
    px = torch.randn((frames, 3, v.image_size, v.image_size), generator=g).to(torch.bfloat16)
    g2 = torch.Generator(device="cpu").manual_seed(1235)
    ids = torch.randint(3, cfg.vocab_size, (1, prompt), generator=g2, dtype=torch.int64)
    ids[0, 4] = -201
    return px, ids


def flops(cfg: Videollama2Config, frames: int, prompt: int, all_logits: bool = False) -> dict:
    """Algorithmic FLOPs of one video->text prefill (2*M*N*K per GEMM; attention 4*S^2*d per head, causal halved;
    ViT counted for the layers actually consumed; last-position logits).  Matches BASELINE.md §2."""
    v = cfg.vision_config
    C, Iv = v.hidden_size, v.intermediate_size
    g = v.image_size // v.patch_size
    npatch = g * g
    nl = v.num_hidden_layers + 1 + cfg.mm_vision_select_layer if cfg.mm_vision_select_layer < 0 else cfg.mm_vision_select_layer
    seq = npatch + (0 if "siglip" in v.model_type else 1)
    Mv = frames * seq
    vit_patch = 2 * frames * npatch * C * 3 * v.patch_size ** 2
    vit_gemm = nl * 2 * Mv * (4 * C * C + 2 * C * Iv)
    vit_attn = nl * frames * v.num_attention_heads * 4 * seq ** 2 * (C // v.num_attention_heads)
    H = cfg.hidden_size
    pad = 0 if cfg.mm_projector_type.endswith("v35") else 1
    to, go = (frames + 2 * pad - 2) // 2 + 1, (g + 2 * pad - 2) // 2 + 1
    M1, M2 = frames * npatch, to * go * go
    s1 = 2 * M1 * (2 * C * H + 7 * H * H)
    conv3d = 2 * M2 * 8 * H * H
    s2 = 2 * M2 * 8 * H * H
    readout = 2 * M2 * 2 * H * H
    dw = 2 * 9 * H * (4 * M1 + 4 * M2)
    S = prompt - 1 + M2
    D = cfg.head_dim
    qkv_n = (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * D
    llm_gemm = cfg.num_hidden_layers * 2 * S * (H * qkv_n + cfg.num_attention_heads * D * H + 3 * H * cfg.intermediate_size)
    llm_attn = cfg.num_hidden_layers * cfg.num_attention_heads * 4 * S * S * D // 2
    head = 2 * (S if all_logits else 1) * H * cfg.vocab_size
    out = {"S": S, "vis_tokens": M2, "vit": vit_patch + vit_gemm + vit_attn, "stc": s1 + conv3d + s2 + readout + dw,
           "llm": llm_gemm + llm_attn + head, "vit_gemm": vit_patch + vit_gemm, "vit_attn": vit_attn,
           "llm_gemm": llm_gemm, "llm_attn": llm_attn}
    out["total"] = out["vit"] + out["stc"] + out["llm"]
    return out
