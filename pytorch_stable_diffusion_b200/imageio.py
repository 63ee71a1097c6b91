"""Image pre-processing on the device: the reference's `input_image.resize((WIDTH, HEIGHT))`
+ `np.array` + `rescale((0, 255), (-1, 1))` (sd/pipeline.py:156-173) as kernels.

PIL.Image.resize defaults to bicubic resampling with antialiasing (the filter support widens by the down-scaling
factor), computed in 8-bit fixed point: per output sample a window of source samples, taps rounded to 22 fractional
bits, horizontal pass then vertical pass, each rounded and clipped to uint8. The tap tables depend only on
(input size, output size): they are computed here on the host exactly as Pillow's precompute_coeffs /
normalize_coeffs_8bpc do (double arithmetic, same operation order) and the two passes run on the device
(sdb_resample_u8), the second one fused with the rescale to [-1, 1]. Byte-exact against Pillow."""
import functools
import math

import numpy as np
import torch

from . import _ext, ops

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@functools.lru_cache(maxsize=64)
def resample_tables(in_size, out_size):
    """(bounds int32 [out, 2], coef int32 [out, ksize]) of Pillow's bicubic resampler for in_size -> out_size."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        w = [_bicubic((x + lo - center + 0.5) * inv) for x in range(hi - lo)]
        total = 0.0
        for v in w:
            total += v
        for x, v in enumerate(w):
            if total != 0.0:
                v = v / total
            coef[o, x] = int(0.5 + v * (1 << PRECISION_BITS)) if v >= 0 else int(-0.5 + v * (1 << PRECISION_BITS))
        bounds[o] = (lo, hi - lo)
    return bounds, coef


def _pass(src, out_size, axis, want_image):
    lib = _ext.lib()
    n, h, w, c = src.shape
    bounds, coef = resample_tables(src.shape[2 - axis], out_size)
    bd = torch.from_numpy(bounds).to(src.device)
    cf = torch.from_numpy(coef).to(src.device)
    shape = (n, out_size, w, c) if axis else (n, h, out_size, c)
    dst = torch.empty(shape, device=src.device, dtype=torch.uint8)
    img = torch.empty(shape, device=src.device, dtype=torch.float32) if want_image else None
    _ext.check(lib.sdb_resample_u8(ops._p(src), ops._p(dst), ops._p(img), n, h, w, c, out_size, axis, ops._p(bd),
                                   ops._p(cf), coef.shape[1], ops._stream()), "sdb_resample_u8")
    return dst, img


def resize_u8(images, width, height, to_image=False):
    """uint8 NHWC [N, H, W, C] on the device -> [N, height, width, C], as PIL.Image.resize((width, height)) would
    produce for each image. to_image=True also returns the fp32 NHWC tensor rescaled to [-1, 1]."""
    if images.dtype != torch.uint8 or images.dim() != 4 or not images.is_cuda:
        raise ValueError("images must be a uint8 NHWC CUDA tensor")
    x = images.contiguous()
    n, h, w, c = x.shape
    img = None
    passes = ([(width, 0)] if w != width else []) + ([(height, 1)] if h != height else [])
    for i, (size, axis) in enumerate(passes):
        x, img = _pass(x, size, axis, to_image and i == len(passes) - 1)
    if to_image and img is None:                    # already the target size: Pillow returns a copy
        img = ops.uint8_to_image(x, out_fp32=True)
    return (x, img) if to_image else x


def load_images(images, width, height, device):
    """PIL images / HWC uint8 arrays (one, or a list - one per sample) -> (uint8, fp32 in [-1, 1]) NHWC device tensors
    of the target size. Images of equal size are resized as one batch."""
    if not isinstance(images, (list, tuple)):
        images = [images]
    arrs = []
    for im in images:
        a = np.asarray(im)
        if a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8:
            raise ValueError(f"input images must be RGB uint8 (H, W, 3); got shape {a.shape} dtype {a.dtype}")
        arrs.append(a)
    out_u8, out_f = [], []
    i = 0
    while i < len(arrs):
        j = i
        while j < len(arrs) and arrs[j].shape == arrs[i].shape:
            j += 1
        batch = torch.from_numpy(np.stack(arrs[i:j])).to(device)
        u8, f = resize_u8(batch, width, height, to_image=True)
        out_u8.append(u8)
        out_f.append(f)
        i = j
    return torch.cat(out_u8), torch.cat(out_f)
