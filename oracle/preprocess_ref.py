"""CPU restatement of the frame preprocessing in front of the vision tower.  TEST INFRASTRUCTURE.

Reference call chain: videollama2/mm_utils.py:132-202 (process_video) / :91-103 (process_image): PIL frames ->
`expand2square` (mm_utils.py:27-38, background = int(255 * image_mean)) -> `processor.preprocess` =
transformers 4.40.0 (pinned in the reference's requirements.txt; NOT in this image) CLIPImageProcessor /
SiglipImageProcessor: resize (PIL `Image.resize`, resample = BICUBIC) -> center crop (CLIP) -> `image * (1/255)` in
float64 cast to float32 -> `(image - mean) / std` in float32 -> channels-first.

The resampling itself lives in Pillow (third-party, libImaging/Resample.c; the image here has Pillow 12.2, the algorithm
is unchanged since Pillow 4): separable convolution with an antialiasing bicubic kernel (a = -0.5, support 2 * max(scale, 1)),
coefficients normalised in double then rounded to 22-bit fixed point, horizontal pass first, every pass rounded and
clipped to uint8.  `resample_u8` restates it in integer numpy arithmetic; tests/test_preprocess.py pins it bit-exactly
against Pillow itself and against the committed goldens (tests/golden/preprocess.pt, oracle/make_golden_preprocess.py)."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
SIGLIP_MEAN = (0.5, 0.5, 0.5)
SIGLIP_STD = (0.5, 0.5, 0.5)


def bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box.
    Returns (bounds int32 [out,2] = (first tap, tap count), kk int32 [out, ksize], ksize)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size      # box edges are C floats
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass along `axis` of an [H, W, C] uint8 image (int64 accumulation == C int32: no overflow)."""
    in_size = img.shape[axis]
    bounds, kk, _ = pil_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resample_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), BICUBIC)` of an [H, W, 3] uint8 image: horizontal pass, then vertical pass; a pass
    whose size does not change is skipped (ImagingResample: need_horizontal / need_vertical)."""
    h, w = img.shape[:2]
    if w != out_w:
        img = _pass(img, out_w, 1)
    if h != out_h:
        img = _pass(img, out_h, 0)
    return np.ascontiguousarray(img)


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """mm_utils.py:27-38 on an [H, W, 3] uint8 array."""
    h, w = img.shape[:2]
    if w == h:
        return img
    s = max(w, h)
    out = np.empty((s, s, 3), dtype=np.uint8)
    out[:] = np.asarray(background, dtype=np.uint8)
    if w > h:
        top = (w - h) // 2
        out[top:top + h] = img
    else:
        left = (h - w) // 2
        out[:, left:left + w] = img
    return out


def background_color(mean: Sequence[float]) -> Tuple[int, ...]:
    return tuple(int(x * 255) for x in mean)        # mm_utils.py:99,196


def resize_target(h: int, w: int, size: int, kind: str) -> Tuple[int, int]:
    """CLIP: shortest edge -> size keeping the aspect (transformers 4.40 get_resize_output_image_size,
    default_to_square=False); SigLIP: (size, size)."""
    if kind == "siglip":
        return size, size
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def normalise_lut(mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """float32 [3, 256]: `rescale` (uint8 * float64 scale -> float32) then `normalize` ((x - mean) / std in float32)."""
    v = (np.arange(256, dtype=np.uint8) * (1 / 255)).astype(np.float32)
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def preprocess_frames(frames: Sequence[np.ndarray], size: int, kind: str = "clip", aspect_ratio: str = "pad",
                      mean=None, std=None) -> Tuple[np.ndarray, np.ndarray]:
    """frames: [H, W, 3] uint8 arrays -> (resized+cropped uint8 [T, size, size, 3], pixel_values float32 [T, 3, size, size])."""
    mean = mean or (SIGLIP_MEAN if kind == "siglip" else CLIP_MEAN)
    std = std or (SIGLIP_STD if kind == "siglip" else CLIP_STD)
    lut = normalise_lut(mean, std)
    u8: List[np.ndarray] = []
    for f in frames:
        if aspect_ratio == "pad":
            f = expand2square(f, background_color(mean))
        oh, ow = resize_target(f.shape[0], f.shape[1], size, kind)
        r = resample_u8(f, oh, ow)
        if kind != "siglip":                                   # CLIPImageProcessor center crop
            top, left = (oh - size) // 2, (ow - size) // 2
            r = r[top:top + size, left:left + size]
        u8.append(r)
    u8a = np.stack(u8)
    px = np.stack([lut[c][u8a[..., c]] for c in range(3)], axis=1)
    return u8a, px
