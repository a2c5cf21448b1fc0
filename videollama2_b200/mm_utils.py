"""Host-side integer work of the hot path (exact parity with the reference required).

tokenizer_multimodal_token  <- videollama2/mm_utils.py:277-302
splice_plan / build_splice  <- the index logic of prepare_inputs_labels_for_multimodal, videollama2/model/videollama2_arch.py:177-261
KeywordsStoppingCriteria    <- videollama2/mm_utils.py:314-345 (token-id tail match only)
Video decoding / resizing (process_video, mm_utils.py:132-202) is CPU I/O outside the accelerated path.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from .constants import DEFAULT_IMAGE_TOKEN, IGNORE_INDEX, MODAL_INDEX_MAP

_MODAL_IDS = tuple(MODAL_INDEX_MAP.values())


def tokenizer_multimodal_token(prompt, tokenizer, multimodal_token=DEFAULT_IMAGE_TOKEN, return_tensors=None):
    """Tokenize text and multimodal tag to input_ids: split on the tag, tokenize each chunk WITHOUT special tokens and
    join the chunks with the (negative) modal index."""
    multimodal_token_index = MODAL_INDEX_MAP.get(multimodal_token, None)
    if multimodal_token_index is None:
        input_ids = tokenizer(prompt, add_special_tokens=False).input_ids
    else:
        chunks = [tokenizer(chunk, add_special_tokens=False).input_ids for chunk in prompt.split(multimodal_token)]
        input_ids = []
        for n, chunk in enumerate(chunks):
            if n:
                input_ids.append(multimodal_token_index)
            input_ids.extend(chunk)
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(input_ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return input_ids


def splice_plan(ids_row: Sequence[int], mm_lengths: Sequence[int], mm_start: int = 0):
    """For one sample: (segments, new_len, n_mm_used).  segments = [('text', src_start, n) | ('mm', mm_idx, n)].
    Every modal placeholder id is replaced by ALL tokens of the next multimodal feature (arch.py:198-211);
    a sample without placeholders still consumes one (empty) feature slot (arch.py:181-191)."""
    segs: List[Tuple[str, int, int]] = []
    total, mm, start = 0, mm_start, 0
    n_ph = 0
    for i, t in enumerate(ids_row):
        if t in _MODAL_IDS:
            n_ph += 1
            if i > start:
                segs.append(("text", start, i - start))
                total += i - start
            segs.append(("mm", mm, mm_lengths[mm]))
            total += mm_lengths[mm]
            mm += 1
            start = i + 1
    if len(ids_row) > start:
        segs.append(("text", start, len(ids_row) - start))
        total += len(ids_row) - start
    if n_ph == 0:
        mm += 1
    return segs, total, mm - mm_start


def build_splice(input_ids: torch.Tensor, mm_lengths: Sequence[int]):
    """Batch index plan.  Returns dict with
       new_len [B], max_len, text_src (b, src_pos) and text_dst (flat row in the [B*max_len] output) index lists,
       mm_dst: list of (mm_idx, b, dst_start, n)."""
    ids = input_ids.tolist()
    plans, lens, mm = [], [], 0
    for row in ids:
        segs, total, used = splice_plan(row, mm_lengths, mm)
        mm += used
        plans.append(segs)
        lens.append(total)
    max_len = max(lens)
    text_b, text_src, text_dst, mm_dst = [], [], [], []
    for b, segs in enumerate(plans):
        pos = 0
        for kind, a, n in segs:
            if kind == "text":
                text_b.extend([b] * n)
                text_src.extend(range(a, a + n))
                text_dst.extend(range(b * max_len + pos, b * max_len + pos + n))
            else:
                mm_dst.append((a, b, pos, n))
            pos += n
    return {"new_len": lens, "max_len": max_len, "text_b": text_b, "text_src": text_src, "text_dst": text_dst,
            "mm_dst": mm_dst}


def spliced_attention_mask(attention_mask: torch.Tensor, old_len: int, new_lens: Sequence[int], max_len: int):
    """arch.py:239-261: True for the inserted positions (left-extended), original mask, False right padding."""
    B = attention_mask.shape[0]
    out = torch.zeros((B, max_len), dtype=attention_mask.dtype, device=attention_mask.device)
    for b in range(B):
        grow = new_lens[b] - old_len
        out[b, :grow] = True
        out[b, grow:grow + old_len] = attention_mask[b]
    return out


def spliced_labels(labels: torch.Tensor, input_ids: torch.Tensor, mm_lengths: Sequence[int], max_len: int):
    """arch.py:205-208,232-237: labels of inserted visual tokens and padding are IGNORE_INDEX."""
    B = labels.shape[0]
    out = torch.full((B, max_len), IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
    mm = 0
    for b in range(B):
        segs, _, used = splice_plan(input_ids[b].tolist(), mm_lengths, mm)
        mm += used
        pos = 0
        for kind, a, n in segs:
            if kind == "text":
                out[b, pos:pos + n] = labels[b, a:a + n]
            pos += n
    return out


class KeywordsStoppingCriteria:
    """Stop when the tail of the generated ids equals one of the keyword id sequences (mm_utils.py:314-345)."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keyword_ids = []
        for kw in keywords:
            cur = tokenizer(kw).input_ids
            if len(cur) > 1 and cur[0] == tokenizer.bos_token_id:
                cur = cur[1:]
            self.keyword_ids.append(torch.tensor(cur))
        self.start_len = input_ids.shape[1]

    def __call__(self, output_ids: torch.Tensor, scores=None, **kw) -> bool:
        for kid in self.keyword_ids:
            n = kid.numel()
            if output_ids.shape[1] >= n and torch.equal(output_ids[0, -n:].cpu(), kid):
                return True
        return False
