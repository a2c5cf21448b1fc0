"""Host-side integer work of the hot path (exact parity with the reference required).

tokenizer_multimodal_token  <- videollama2/mm_utils.py:277-302
splice_plan / build_splice  <- the index logic of prepare_inputs_labels_for_multimodal, videollama2/model/videollama2_arch.py:177-261
KeywordsStoppingCriteria    <- videollama2/mm_utils.py:314-345 (token-id tail match only)
frame_sample                <- videollama2/mm_utils.py:106-129 (frame index arithmetic, exact)
process_video / process_image <- videollama2/mm_utils.py:91-103,132-202 for frames that are already decoded (arrays, PIL
                               images, uint8 tensors): pad-to-square, Pillow-exact resize, crop, normalise on the GPU
                               (videollama2_b200/preprocess.py).  Container decoding (decord / imageio) stays CPU I/O.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from .constants import (DEFAULT_IMAGE_TOKEN, IGNORE_INDEX, MAX_FRAMES, MODAL_INDEX_MAP, NUM_FRAMES,
                        NUM_FRAMES_PER_SECOND)

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
    """mm_utils.py:314-345.  A row stops when (a) the tail of its ids equals one of the keyword id sequences, or (b) the
    decoded text of its last `offset` ids contains a keyword, offset = min(len(ids) - start_len, longest keyword) exactly
    as the reference computes it (with `generate(inputs_embeds=...)` the ids hold only NEW tokens while start_len is the
    prompt length, so the window `ids[-offset:]` follows Python's slicing of a possibly negative / zero offset); the batch
    stops when every row does."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = list(keywords)
        self.keyword_ids = []
        self.max_keyword_len = 0
        for kw in self.keywords:
            cur = tokenizer(kw).input_ids
            if len(cur) > 1 and cur[0] == tokenizer.bos_token_id:
                cur = cur[1:]
            self.max_keyword_len = max(self.max_keyword_len, len(cur))
            self.keyword_ids.append(torch.tensor(cur))
        self.tokenizer = tokenizer
        self.start_len = input_ids.shape[1]

    def _row_stops(self, row: torch.Tensor) -> bool:
        row = row.cpu()
        for kid in self.keyword_ids:
            n = kid.numel()
            tail = row[-n:]
            if tail.numel() == n and torch.equal(tail, kid):
                return True
        decode = getattr(self.tokenizer, "batch_decode", None)
        if decode is None:
            return False
        offset = min(row.numel() - self.start_len, self.max_keyword_len)
        text = decode(row[-offset:].unsqueeze(0), skip_special_tokens=True)[0]
        return any(kw in text for kw in self.keywords)

    def __call__(self, output_ids: torch.Tensor, scores=None, **kw) -> bool:
        return all(self._row_stops(output_ids[i]) for i in range(output_ids.shape[0]))


# ----------------------------------------------------------------------------------------------------------------
# frame sampling + preprocessing of decoded frames
# ----------------------------------------------------------------------------------------------------------------
def frame_sample(duration, mode="uniform", num_frames=None, fps=None):
    """Indices of the frames to keep out of `duration` decoded frames (mm_utils.py:106-129).
    uniform: the midpoint of each of `num_frames` equal segments of [0, duration-1], rounded half-up via +1e-6;
    fps: one frame per `fps // NUM_FRAMES_PER_SECOND` source frames, starting half a segment in."""
    if mode == "uniform":
        assert num_frames is not None, "Number of frames must be provided for uniform sampling."
        seg = float(duration - 1) / num_frames
        mids = [(seg * i + seg * (i + 1)) / 2 for i in range(num_frames)]
        return np.round(np.array(mids) + 1e-6).astype(int)
    if mode == "fps":
        assert fps is not None, "FPS must be provided for FPS sampling."
        step = min(fps // NUM_FRAMES_PER_SECOND, duration)
        return np.arange(step // 2, duration, step, dtype=int)
    raise ImportError(f"Unsupported frame sampling mode: {mode}")


def _processor_kind(processor) -> Tuple[str, int]:
    size = processor.size
    if "shortest_edge" in size:
        return "clip", int(processor.crop_size["height"])
    return "siglip", int(size["height"])


def _as_uint8_frames(video) -> List[np.ndarray]:
    if isinstance(video, torch.Tensor):
        return [video]                                   # handled as one [T,H,W,3] block by the caller
    if isinstance(video, np.ndarray):
        return [f for f in video] if video.ndim == 4 else [video]
    if isinstance(video, (list, tuple)) and len(video) and isinstance(video[0], str):
        raise NotImplementedError("decode the files first (PIL / decord are CPU I/O); pass frames as arrays or images")
    if isinstance(video, (list, tuple)):
        return [np.asarray(f.convert("RGB")) if hasattr(f, "convert") else np.asarray(f) for f in video]
    raise ValueError(f"Unsupported video path type: {type(video)}")


def process_video(video_path, processor, s=None, e=None, aspect_ratio="pad", num_frames=NUM_FRAMES, device="cuda"):
    """Decoded frames -> bf16 pixel_values [T,3,S,S] on `device` (mm_utils.py:132-202 from step 4 on):
    missing frames are appended as black frames (with the reference's swapped width/height, mm_utils.py:190-191),
    at most MAX_FRAMES are kept, then every frame goes through expand2square + processor.preprocess on the GPU."""
    from . import preprocess
    if isinstance(video_path, str):
        raise NotImplementedError("container decoding (decord / imageio, mm_utils.py:143-176) is CPU I/O outside the "
                                  "accelerated path: decode the sampled frames (frame_sample) and pass them in")
    kind, size = _processor_kind(processor)
    kw = dict(kind=kind, aspect_ratio=aspect_ratio)
    if isinstance(video_path, torch.Tensor) and video_path.dim() == 4 and \
            (num_frames is None or video_path.shape[0] >= num_frames):
        frames = video_path[:MAX_FRAMES].to(device)
        return preprocess.preprocess_frames(frames, size, processor.image_mean, processor.image_std, **kw)
    frames = _as_uint8_frames(video_path.cpu().numpy() if isinstance(video_path, torch.Tensor) else video_path)
    while num_frames is not None and len(frames) < num_frames:
        h, w = frames[-1].shape[:2]
        frames.append(np.zeros((w, h, 3), dtype=np.uint8))          # (*PIL.size, 3) = (width, height, 3) in the reference
    frames = frames[:MAX_FRAMES]
    out: List[torch.Tensor] = []
    i = 0
    while i < len(frames):                                          # consecutive frames of one shape share a launch
        j = i
        while j < len(frames) and frames[j].shape == frames[i].shape:
            j += 1
        block = torch.from_numpy(np.ascontiguousarray(np.stack(frames[i:j]))).to(device)
        out.append(preprocess.preprocess_frames(block, size, processor.image_mean, processor.image_std, **kw))
        i = j
    return torch.cat(out, 0) if len(out) > 1 else out[0]


def process_image(image, processor, aspect_ratio="pad", device="cuda"):
    """One decoded image (array / PIL image / uint8 tensor [H,W,3]) -> bf16 [1,3,S,S] (mm_utils.py:91-103)."""
    if isinstance(image, str):
        raise NotImplementedError("decode the file first (PIL is CPU I/O); pass the image as an array or a PIL image")
    if isinstance(image, torch.Tensor):
        image = image.cpu().numpy()
    arr = np.asarray(image.convert("RGB")) if hasattr(image, "convert") else np.asarray(image)
    return process_video([arr], processor, aspect_ratio=aspect_ratio, num_frames=None, device=device)
