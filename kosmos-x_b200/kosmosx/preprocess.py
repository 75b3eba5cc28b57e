"""Host side of the device preprocessing: geometry and coefficient tables of CLIPImageProcessor's
resize + centre crop (reference: kosmosx/model.py:36-38,81-97 -> transformers 4.35 `CLIPImageProcessor`:
``get_resize_output_image_size(shortest_edge)``, ``image_transforms.resize`` = ``PIL.Image.resize(BICUBIC)``,
``center_crop``).  Only index arithmetic and a few KB of filter taps are computed here, once per input size (cached);
the pixels are resampled on the device by kx_resize_crop_u8 in PIL's own fixed-point arithmetic, bit for bit.
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch

from . import _abi, ops
from ._abi import check, lib

PRECISION_BITS = 32 - 8 - 2          # PIL ImagingResample, 8 bits per channel


def resize_output_size(height: int, width: int, shortest_edge: int) -> tuple[int, int]:
    """transformers ``get_resize_output_image_size(image, size=shortest_edge, default_to_square=False)``."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = shortest_edge, int(shortest_edge * long / short)
    return (new_long, new_short) if width <= height else (new_short, new_long)


def _bicubic(x: float) -> float:
    a = -0.5                         # PIL's bicubic_filter
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@functools.lru_cache(maxsize=256)
def bicubic_taps(in_size: int, out_size: int) -> tuple[np.ndarray, np.ndarray]:
    """PIL ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for one axis: int32 taps [out_size, ksize] and
    (first input index, tap count) pairs [out_size, 2].  Same operations in the same order as the C code, in float64."""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    taps = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):    # C: (int)(±0.5 + v * (1 << PRECISION_BITS)), truncation toward zero
            taps[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return taps, bounds


@functools.lru_cache(maxsize=64)
def _plan(in_h: int, in_w: int, size: int, crop: int, device_index: int):
    new_h, new_w = resize_output_size(in_h, in_w, size)
    if crop > new_h or crop > new_w:
        raise ValueError(f"crop {crop} exceeds the resized image {new_h}x{new_w} (transformers pads there; not built)")
    top, left = (new_h - crop) // 2, (new_w - crop) // 2             # transformers center_crop
    kx, bx = bicubic_taps(in_w, new_w)
    ky, by = bicubic_taps(in_h, new_h)
    kx, bx = kx[left:left + crop], bx[left:left + crop]
    ky, by = ky[top:top + crop], by[top:top + crop]
    y0 = int(by[:, 0].min())
    y1 = int((by[:, 0] + by[:, 1]).max())
    dev = torch.device("cuda", device_index)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return dict(kx=t(kx), bx=t(bx), ky=t(ky), by=t(by), ksize_x=kx.shape[1], ksize_y=ky.shape[1], y0=y0, rows=y1 - y0,
                resized=(new_h, new_w), top=top, left=left)


def resize_center_crop_u8(pixels: torch.Tensor, size: int = 224, crop: int | None = None) -> torch.Tensor:
    """uint8 CUDA images (N,H,W,3) or (N,3,H,W) of any size -> uint8 (N, crop, crop, 3): CLIPImageProcessor's
    shortest-edge bicubic resize + centre crop, bit-identical to the PIL path of the reference's processor."""
    crop = size if crop is None else crop
    if not pixels.is_cuda or pixels.dtype != torch.uint8 or pixels.ndim != 4:
        raise TypeError("resize_center_crop_u8 takes a 4-D uint8 CUDA tensor")
    if pixels.shape[-1] == 3 and pixels.shape[1] != 3:
        cl, (in_h, in_w) = 1, pixels.shape[1:3]
    elif pixels.shape[1] == 3:
        cl, (in_h, in_w) = 0, pixels.shape[2:4]
    else:
        raise ValueError(f"uint8 images must be (N,H,W,3) or (N,3,H,W), got {tuple(pixels.shape)}")
    pixels = pixels.contiguous()
    n = pixels.shape[0]
    p = _plan(int(in_h), int(in_w), int(size), int(crop), pixels.device.index or 0)
    tmp = torch.empty(n, p["rows"], crop, 3, dtype=torch.uint8, device=pixels.device)
    out = torch.empty(n, crop, crop, 3, dtype=torch.uint8, device=pixels.device)
    check(lib.kx_resize_crop_u8(pixels.data_ptr(), cl, n, int(in_h), int(in_w), p["kx"].data_ptr(), p["bx"].data_ptr(), p["ksize_x"],
                                p["ky"].data_ptr(), p["by"].data_ptr(), p["ksize_y"], p["y0"], p["rows"], crop, crop,
                                tmp.data_ptr(), out.data_ptr(), ops._stream()), "kx_resize_crop_u8")
    return out
