"""Plain-PyTorch CPU restatement of the reference's video->text forward path.  TEST INFRASTRUCTURE (oracle/__init__.py).

Functional, state-dict driven (HF names), no nn.Modules; `dtype` selects the arithmetic (float32 = "G32" golden on
bf16-rounded weights).  Each function cites the reference (or the third-party library the reference calls; HF: =
transformers/models, pinned transformers==4.40.0 / installed 5.5.0, timm==1.0.3 restated in SURVEY.md §8c).
Pinned against the real reference classes by tests/test_oracle_vs_reference.py and tests/golden/*.pt."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
VPFX = "model.vision_tower.vision_tower.vision_model."
PPFX = "model.mm_projector."
MODAL_IDS = (-200, -201, -202)  # videollama2/constants.py:28-32


def _w(sd: SD, name: str, dtype) -> torch.Tensor:
    return sd[name].to(dtype)


# ----------------------------------------------------------------------------------------------------------------
# CLIP ViT tower — encoder.py:41-53 -> HF:clip/modeling_clip.py:202-218 (embeddings), :282-336 (attention),
# :339-351 (MLP, quick_gelu), :354-385 (layer), :647-696 (encoder + hidden_states), feature_select encoder.py:31-39
# ----------------------------------------------------------------------------------------------------------------
def vit_hidden_states(sd: SD, v, pixels: torch.Tensor, dtype=torch.float32, n_layers: Optional[int] = None,
                      pfx: str = VPFX, sdpa: bool = False) -> List[torch.Tensor]:
    x = pixels.to(dtype)
    Fn = x.shape[0]
    siglip = getattr(v, "kind", "clip") == "siglip"
    if siglip:
        # HF SiglipVisionEmbeddings (HF:siglip/modeling_siglip.py): conv WITH bias, no class token, no pre-LN
        pe = F.conv2d(x, _w(sd, pfx + "embeddings.patch_embedding.weight", dtype),
                      _w(sd, pfx + "embeddings.patch_embedding.bias", dtype), stride=v.patch)
        h = (pe.flatten(2).transpose(1, 2) + _w(sd, pfx + "embeddings.position_embedding.weight", dtype)).contiguous()
    else:
        pe = F.conv2d(x, _w(sd, pfx + "embeddings.patch_embedding.weight", dtype), stride=v.patch)  # no bias
        pe = pe.flatten(2).transpose(1, 2)                                                         # [F, np, C]
        cls = _w(sd, pfx + "embeddings.class_embedding", dtype).expand(Fn, 1, -1)
        h = torch.cat([cls, pe], dim=1) + _w(sd, pfx + "embeddings.position_embedding.weight", dtype)
        h = F.layer_norm(h, (v.hidden,), _w(sd, pfx + "pre_layrnorm.weight", dtype), _w(sd, pfx + "pre_layrnorm.bias", dtype), v.eps)
    hs = [h]
    d = v.hidden // v.heads
    for i in range(v.layers if n_layers is None else n_layers):
        p = f"{pfx}encoder.layers.{i}."
        r = h
        y = F.layer_norm(h, (v.hidden,), _w(sd, p + "layer_norm1.weight", dtype), _w(sd, p + "layer_norm1.bias", dtype), v.eps)
        q = F.linear(y, _w(sd, p + "self_attn.q_proj.weight", dtype), _w(sd, p + "self_attn.q_proj.bias", dtype))
        k = F.linear(y, _w(sd, p + "self_attn.k_proj.weight", dtype), _w(sd, p + "self_attn.k_proj.bias", dtype))
        vv = F.linear(y, _w(sd, p + "self_attn.v_proj.weight", dtype), _w(sd, p + "self_attn.v_proj.bias", dtype))
        S = y.shape[1]
        q = q.view(Fn, S, v.heads, d).transpose(1, 2)
        k = k.view(Fn, S, v.heads, d).transpose(1, 2)
        vv = vv.view(Fn, S, v.heads, d).transpose(1, 2)
        if sdpa:    # what HF runs with attn_implementation="sdpa" (the timing baseline); the eager form below is the golden
            o = F.scaled_dot_product_attention(q, k, vv, scale=d ** -0.5).transpose(1, 2).reshape(Fn, S, v.hidden)
        else:
            att = torch.softmax((q @ k.transpose(-1, -2)).float() * d ** -0.5, dim=-1).to(dtype)
            o = (att @ vv).transpose(1, 2).reshape(Fn, S, v.hidden)
        h = r + F.linear(o, _w(sd, p + "self_attn.out_proj.weight", dtype), _w(sd, p + "self_attn.out_proj.bias", dtype))
        r = h
        y = F.layer_norm(h, (v.hidden,), _w(sd, p + "layer_norm2.weight", dtype), _w(sd, p + "layer_norm2.bias", dtype), v.eps)
        y = F.linear(y, _w(sd, p + "mlp.fc1.weight", dtype), _w(sd, p + "mlp.fc1.bias", dtype))
        y = F.gelu(y, approximate="tanh") if siglip else y * torch.sigmoid(1.702 * y)
        h = r + F.linear(y, _w(sd, p + "mlp.fc2.weight", dtype), _w(sd, p + "mlp.fc2.bias", dtype))
        hs.append(h)
    return hs


def vit_features(sd: SD, v, pixels: torch.Tensor, select_layer: int = -2, dtype=torch.float32, sdpa: bool = False) -> torch.Tensor:
    """CLIPVisionTower.forward + feature_select('patch'): hidden_states[select_layer][:, 1:] (encoder.py:31-53).
    Layers after the selected one are skipped (they do not influence the result)."""
    n = v.layers + 1 + select_layer if select_layer < 0 else select_layer
    hs = vit_hidden_states(sd, v, pixels, dtype, n_layers=n, sdpa=sdpa)
    return hs[n] if getattr(v, "kind", "clip") == "siglip" else hs[n][:, 1:]


# ----------------------------------------------------------------------------------------------------------------
# STC connector — projector.py:133-215, channels-last restatement; timm regnet.Bottleneck
# ----------------------------------------------------------------------------------------------------------------
def _ln_c(x, sd, name, dtype, eps):
    c = x.shape[-1]
    return F.layer_norm(x, (c,), _w(sd, name + ".weight", dtype), _w(sd, name + ".bias", dtype), eps)


def regstage_block(sd: SD, p: str, x: torch.Tensor, dtype, eps: float) -> torch.Tensor:
    """x: [F,H,W,Cin] channels-last.  timm Bottleneck(bottle_ratio=1, group_size=1, se_ratio=.25) with LayerNormAct2d+SiLU."""
    w1 = _w(sd, p + "conv1.conv.weight", dtype)
    c, cin = w1.shape[0], w1.shape[1]
    y = F.silu(_ln_c(x @ w1.view(c, cin).t(), sd, p + "conv1.bn", dtype, eps))
    wd = _w(sd, p + "conv2.conv.weight", dtype)                                       # [C,1,3,3] depthwise
    y = F.conv2d(y.permute(0, 3, 1, 2), wd, padding=1, groups=c).permute(0, 2, 3, 1)
    y = F.silu(_ln_c(y, sd, p + "conv2.bn", dtype, eps))
    s = y.mean(dim=(1, 2))                                                            # [F,C]  SE squeeze (per frame)
    f1 = _w(sd, p + "se.fc1.weight", dtype)
    f2 = _w(sd, p + "se.fc2.weight", dtype)
    s = F.silu(s @ f1.view(f1.shape[0], c).t() + _w(sd, p + "se.fc1.bias", dtype))
    s = torch.sigmoid(s @ f2.view(c, f2.shape[1]).t() + _w(sd, p + "se.fc2.bias", dtype))
    y = y * s[:, None, None, :]
    y = _ln_c(y @ _w(sd, p + "conv3.conv.weight", dtype).view(c, c).t(), sd, p + "conv3.bn", dtype, eps)
    if p + "downsample.conv.weight" in sd:
        r = _ln_c(x @ _w(sd, p + "downsample.conv.weight", dtype).view(c, cin).t(), sd, p + "downsample.bn", dtype, eps)
    else:
        r = x
    return F.silu(y + r)


def stc_stages(sd: SD, x: torch.Tensor, pad: int = 1, depth: int = 4, dtype=torch.float32, eps: float = 1e-5,
               pfx: str = PPFX) -> Dict[str, torch.Tensor]:
    """x: [b,T,np,Cin] -> dict(s1, sampler, s2, out[b, T'*H'*W', C]).  projector.py:189-215."""
    b, T, n, cin = x.shape
    hw = int(n ** 0.5)                                                                 # projector.py:198
    a = x.to(dtype).reshape(b * T, hw, hw, cin)
    for i in range(1, depth + 1):
        a = regstage_block(sd, f"{pfx}s1.b{i}.", a, dtype, eps)
    s1 = a
    c = a.shape[-1]
    # Conv3d(k=s=2, padding=pad) + SiLU   projector.py:164-174
    vol = a.view(b, T, hw, hw, c).permute(0, 4, 1, 2, 3)
    smp = F.silu(F.conv3d(vol, _w(sd, pfx + "sampler.0.weight", dtype), _w(sd, pfx + "sampler.0.bias", dtype), stride=2, padding=pad))
    To, Ho, Wo = smp.shape[2:]
    a = smp.permute(0, 2, 3, 4, 1).reshape(b * To, Ho, Wo, c)
    sampler = a
    for i in range(1, depth + 1):
        a = regstage_block(sd, f"{pfx}s2.b{i}.", a, dtype, eps)
    s2 = a
    y = a.reshape(b, To * Ho * Wo, c)                                                  # 'b (t h w) d'
    y = F.linear(y, _w(sd, pfx + "readout.0.weight", dtype), _w(sd, pfx + "readout.0.bias", dtype))
    y = F.gelu(y)                                                                      # nn.GELU() = erf
    y = F.linear(y, _w(sd, pfx + "readout.2.weight", dtype), _w(sd, pfx + "readout.2.bias", dtype))
    return {"s1": s1, "sampler": sampler, "s2": s2, "out": y}


def stc_forward(sd: SD, x: torch.Tensor, pad: int = 1, depth: int = 4, dtype=torch.float32, eps: float = 1e-5) -> torch.Tensor:
    return stc_stages(sd, x, pad, depth, dtype, eps)["out"]


# ----------------------------------------------------------------------------------------------------------------
# token / embedding splice — mm_utils.py:277-302 and videollama2_arch.py:161-263 (batch 1..n, index math exact)
# ----------------------------------------------------------------------------------------------------------------
def tokenizer_multimodal_token(prompt: str, tokenizer, multimodal_token: str = "<image>") -> List[int]:
    """mm_utils.py:277-302: split on the tag, tokenize chunks without special tokens, interleave the modal index."""
    idx = {"<image>": -200, "<video>": -201, "<audio>": -202}.get(multimodal_token)
    if idx is None:
        return tokenizer(prompt, add_special_tokens=False).input_ids
    chunks = [tokenizer(c, add_special_tokens=False).input_ids for c in prompt.split(multimodal_token)]
    out: List[int] = []
    for i, ch in enumerate(chunks):
        if i > 0:
            out.append(idx)
        out.extend(ch)
    return out


def splice_plan(ids_row: List[int], n_mm_tokens: List[int]) -> Tuple[List[Tuple[str, int, int]], int]:
    """Segment list for one sample: ('text', src_start, length) | ('mm', mm_index, length); total length.
    videollama2_arch.py:177-224: every modal id is replaced by ALL tokens of the next mm feature."""
    segs: List[Tuple[str, int, int]] = []
    total = 0
    mm = 0
    start = 0
    for i, t in enumerate(ids_row):
        if t in MODAL_IDS:
            if i > start:
                segs.append(("text", start, i - start))
                total += i - start
            segs.append(("mm", mm, n_mm_tokens[mm]))
            total += n_mm_tokens[mm]
            mm += 1
            start = i + 1
    if len(ids_row) > start:
        segs.append(("text", start, len(ids_row) - start))
        total += len(ids_row) - start
    return segs, total


def splice_embeddings(ids: torch.Tensor, embed: torch.Tensor, mm_features: torch.Tensor) -> torch.Tensor:
    """Batch-1 splice (videollama2_arch.py:198-220): ids [P], embed [V,H], mm_features [L,H] -> [P-1+L, H]."""
    row = ids.tolist()
    segs, total = splice_plan(row, [mm_features.shape[0]] * sum(t in MODAL_IDS for t in row))
    parts = []
    for kind, a, n in segs:
        parts.append(embed[ids[a:a + n]] if kind == "text" else mm_features)
    out = torch.cat(parts, dim=0)
    assert out.shape[0] == total
    return out


# ----------------------------------------------------------------------------------------------------------------
# Mistral / Qwen2 decoder — HF:mistral/modeling_mistral.py:35-48 (MLP), :51-82 (RoPE), :122-177 (attention),
# :182-199 (RMSNorm), :202-239 (layer), :262-323 (rotary), :328-398 (model), :402-470 (lm head);
# HF:qwen2/modeling_qwen2.py:187-246 (q/k/v bias).
# ----------------------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def rope_cos_sin(S: int, D: int, theta: float, dtype, pos0: int = 0):
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    fr = torch.outer(torch.arange(pos0, pos0 + S, dtype=torch.float32), inv_freq)
    emb = torch.cat([fr, fr], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def decoder_layer(sd: SD, l, i: int, h: torch.Tensor, cos, sin, dtype, sdpa: bool = False) -> torch.Tensor:
    p = f"model.layers.{i}."
    S = h.shape[0]
    d = l.head_dim
    r = h
    y = rmsnorm(h, _w(sd, p + "input_layernorm.weight", dtype), l.eps)

    def proj(nm):
        b = sd.get(p + f"self_attn.{nm}.bias")
        return F.linear(y, _w(sd, p + f"self_attn.{nm}.weight", dtype), None if b is None else b.to(dtype))

    q = proj("q_proj").view(S, l.heads, d).transpose(0, 1)
    k = proj("k_proj").view(S, l.kv_heads, d).transpose(0, 1)
    v = proj("v_proj").view(S, l.kv_heads, d).transpose(0, 1)
    q = q * cos + _rot_half(q) * sin
    k = k * cos + _rot_half(k) * sin
    g = l.heads // l.kv_heads
    k = k.repeat_interleave(g, dim=0)
    v = v.repeat_interleave(g, dim=0)
    if sdpa:
        o = F.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=True, scale=d ** -0.5)[0]
        o = o.transpose(0, 1).reshape(S, l.heads * d)
    else:
        s = (q @ k.transpose(-1, -2)).float() * d ** -0.5
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
        o = (torch.softmax(s, dim=-1).to(dtype) @ v).transpose(0, 1).reshape(S, l.heads * d)
    h = r + F.linear(o, _w(sd, p + "self_attn.o_proj.weight", dtype))
    r = h
    y = rmsnorm(h, _w(sd, p + "post_attention_layernorm.weight", dtype), l.eps)
    y = F.silu(F.linear(y, _w(sd, p + "mlp.gate_proj.weight", dtype))) * F.linear(y, _w(sd, p + "mlp.up_proj.weight", dtype))
    return r + F.linear(y, _w(sd, p + "mlp.down_proj.weight", dtype))


def decoder_forward(sd: SD, l, embeds: torch.Tensor, dtype=torch.float32, all_logits: bool = True,
                    return_hidden: bool = False):
    """embeds [S,H] -> logits [S,V] (or [1,V] for the last position)."""
    h = embeds.to(dtype)
    cos, sin = rope_cos_sin(h.shape[0], l.head_dim, l.theta, dtype)
    hidden = [h]
    for i in range(l.layers):
        h = decoder_layer(sd, l, i, h, cos, sin, dtype)
        hidden.append(h)
    hn = rmsnorm(h, _w(sd, "model.norm.weight", dtype), l.eps)
    if not all_logits:
        hn = hn[-1:]
    logits = F.linear(hn, _w(sd, "lm_head.weight", dtype))
    return (logits, hidden) if return_hidden else logits


# ----------------------------------------------------------------------------------------------------------------
# whole path — videollama2_mistral.py:63-108 (forward) with videollama2_arch.py:114-134,161-263
# ----------------------------------------------------------------------------------------------------------------
def encode_video(sd: SD, cfg, pixels: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    feats = vit_features(sd, cfg.vision, pixels, cfg.select_layer, dtype)
    return stc_forward(sd, feats[None], cfg.stc_pad, cfg.stc_depth, dtype)[0]


def full_forward(sd: SD, cfg, pixels: torch.Tensor, ids: torch.Tensor, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    feats = vit_features(sd, cfg.vision, pixels, cfg.select_layer, dtype)
    stc = stc_stages(sd, feats[None], cfg.stc_pad, cfg.stc_depth, dtype)
    mm = stc["out"][0]
    embeds = splice_embeddings(ids[0], _w(sd, "model.embed_tokens.weight", dtype), mm)
    logits, hidden = decoder_forward(sd, cfg.llm, embeds, dtype, return_hidden=True)
    return {"vit": feats, "stc_s1": stc["s1"], "stc_sampler": stc["sampler"], "stc_s2": stc["s2"], "mm": mm,
            "inputs_embeds": embeds, "hidden_last": hidden[-1], "logits": logits}
