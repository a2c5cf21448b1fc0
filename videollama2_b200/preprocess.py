"""Device-side frame preprocessing (the step in front of the tower).

Host logic mirrored from the reference's call chain — `expand2square` (videollama2/mm_utils.py:27-38), the resize target
and centre crop of transformers 4.40 CLIPImageProcessor / SiglipImageProcessor (`processor.preprocess`,
mm_utils.py:101,197-201) and Pillow's coefficient precomputation (libImaging/Resample.c precompute_coeffs +
normalize_coeffs_8bpc) — feeding `vl2_preprocess_frames`, which does the per-pixel work on the GPU.  The resized uint8
image is bit-identical to Pillow's; the bf16 pixel_values are the bf16 rounding of the reference's float32 values."""
from __future__ import annotations

import ctypes as C
import functools
import math
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@functools.lru_cache(maxsize=64)
def resample_tables(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """(bounds int32 [out,2], kk int32 [out,ksize], ksize): Pillow's antialiased bicubic taps for one axis in 22-bit
    fixed point.  in_size == out_size yields the identity (Pillow skips that pass; the result is the same)."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / filterscale
    one = 1 << PRECISION_BITS
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * inv) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * one) if k < 0 else int(0.5 + k * one)
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def normalise_lut(mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """float32 [3,256]: transformers 4.40 `rescale` (uint8 * float64(1/255) -> float32) then `normalize`."""
    v = (np.arange(256, dtype=np.uint8) * (1 / 255)).astype(np.float32)
    m = np.array(mean, dtype=np.float32)
    s = np.array(std, dtype=np.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def geometry(h: int, w: int, size: int, kind: str, aspect_ratio: str):
    """-> (canvas_h, canvas_w, off_y, off_x, out_h, out_w, crop_top, crop_left)."""
    ch, cw, oy, ox = h, w, 0, 0
    if aspect_ratio == "pad" and h != w:                       # expand2square (mm_utils.py:27-38)
        s = max(h, w)
        ch = cw = s
        if w > h:
            oy = (w - h) // 2
        else:
            ox = (h - w) // 2
    if kind == "siglip":                                        # SiglipImageProcessor: resize to (size, size), no crop
        return ch, cw, oy, ox, size, size, 0, 0
    short, long = (cw, ch) if cw <= ch else (ch, cw)            # CLIPImageProcessor: shortest edge, then centre crop
    new_long = int(size * long / short)
    out_h, out_w = (new_long, size) if cw <= ch else (size, new_long)
    return ch, cw, oy, ox, out_h, out_w, (out_h - size) // 2, (out_w - size) // 2


class PreprocessArgs(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("canvas_h", C.c_int32), ("canvas_w", C.c_int32), ("off_y", C.c_int32), ("off_x", C.c_int32),
                ("pad_rgb", C.c_uint8 * 4), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("crop_top", C.c_int32), ("crop_left", C.c_int32), ("crop", C.c_int32),
                ("bounds_h", C.c_void_p), ("kk_h", C.c_void_p), ("bounds_v", C.c_void_p), ("kk_v", C.c_void_p),
                ("ksize_h", C.c_int32), ("ksize_v", C.c_int32), ("lut", C.c_void_p), ("tmp", C.c_void_p),
                ("out_bf16", C.c_void_p), ("out_u8", C.c_void_p)]


_dev_tables = {}


def _tables_on(device, in_size: int, out_size: int):
    key = (torch.device(device).index, in_size, out_size)
    t = _dev_tables.get(key)
    if t is None:
        b, k, ks = resample_tables(in_size, out_size)
        t = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), ks)
        _dev_tables[key] = t
    return t


def preprocess_frames(frames: torch.Tensor, size: int, mean: Sequence[float], std: Sequence[float], *,
                      kind: str = "clip", aspect_ratio: str = "pad", return_u8: bool = False,
                      dtype: torch.dtype = torch.bfloat16):
    """frames: uint8 CUDA tensor [T,H,W,3] (RGB, HWC — what decord / PIL hand out) -> bf16 [T,3,size,size]
    (+ the resized uint8 window [T,size,size,3] when return_u8)."""
    if not frames.is_cuda:
        raise _lib.Vl2Error("preprocess_frames needs a CUDA tensor (no CPU fallback)")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected uint8 [T,H,W,3] frames, got {frames.dtype} {tuple(frames.shape)}")
    frames = frames.contiguous()
    T, H, W, _ = frames.shape
    dev = frames.device
    ch, cw, oy, ox, out_h, out_w, ctop, cleft = geometry(H, W, size, kind, aspect_ratio)
    bh, kh, ksh = _tables_on(dev, cw, out_w)
    bv, kv, ksv = _tables_on(dev, ch, out_h)
    lut = torch.from_numpy(normalise_lut(mean, std)).to(dev)
    out = torch.empty((T, 3, size, size), device=dev, dtype=dtype)
    u8 = torch.empty((T, size, size, 3), device=dev, dtype=torch.uint8) if return_u8 else None
    a = PreprocessArgs()
    a.frames, a.T, a.H, a.W = frames.data_ptr(), T, H, W
    a.canvas_h, a.canvas_w, a.off_y, a.off_x = ch, cw, oy, ox
    for i, v in enumerate(int(x * 255) for x in mean):          # expand2square background (mm_utils.py:99,196)
        a.pad_rgb[i] = v
    a.out_h, a.out_w, a.crop_top, a.crop_left, a.crop = out_h, out_w, ctop, cleft, size
    a.bounds_h, a.kk_h, a.ksize_h = bh.data_ptr(), kh.data_ptr(), ksh
    a.bounds_v, a.kk_v, a.ksize_v = bv.data_ptr(), kv.data_ptr(), ksv
    a.lut = lut.data_ptr()
    lib = _lib.load(dtype)
    tmp = torch.empty((int(lib.vl2_preprocess_workspace(C.byref(a))),), device=dev, dtype=torch.uint8)
    a.tmp, a.out_bf16, a.out_u8 = tmp.data_ptr(), out.data_ptr(), (u8.data_ptr() if u8 is not None else None)
    check(lib.vl2_preprocess_frames(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
          "vl2_preprocess_frames")
    return (out, u8) if return_u8 else out
