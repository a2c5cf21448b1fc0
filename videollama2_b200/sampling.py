"""Token selection for `generate(do_sample=True, ...)`: the logits warpers HF's GenerationMixin applies for the arguments
the reference's mm_infer passes (videollama2/__init__.py:93-110: temperature, top_p; top_k stays at GenerationConfig's
default 50) — temperature -> top-k -> top-p -> softmax -> multinomial.  Plain torch on the logits row (any device)."""
from __future__ import annotations

from typing import Optional

import torch


def warp_logits(logits: torch.Tensor, temperature: float = 1.0, top_k: int = 50, top_p: float = 1.0,
                min_tokens_to_keep: int = 1) -> torch.Tensor:
    """[..., V] fp32 scores after TemperatureLogitsWarper, TopKLogitsWarper and TopPLogitsWarper (HF semantics: top-p
    removes the low-probability tail whose cumulative probability is <= 1 - top_p, always keeping the best token)."""
    scores = logits.float()
    if temperature is not None and temperature != 1.0:
        if temperature <= 0:
            raise ValueError("temperature must be positive when sampling")
        scores = scores / temperature
    if top_k is not None and top_k > 0:
        k = min(max(int(top_k), min_tokens_to_keep), scores.shape[-1])
        kth = torch.topk(scores, k, dim=-1).values[..., -1, None]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        sorted_scores, sorted_idx = torch.sort(scores, descending=False, dim=-1)
        cum = sorted_scores.softmax(dim=-1).cumsum(dim=-1)
        remove = cum <= (1.0 - top_p)
        remove[..., -min_tokens_to_keep:] = False
        scores = scores.masked_fill(remove.scatter(-1, sorted_idx, remove), float("-inf"))
    return scores


def sample_token(logits: torch.Tensor, temperature: float, top_p: float, top_k: int = 50,
                 generator: Optional[torch.Generator] = None) -> int:
    probs = warp_logits(logits.reshape(1, -1), temperature, top_k, top_p).softmax(dim=-1)
    return int(torch.multinomial(probs, 1, generator=generator).item())
