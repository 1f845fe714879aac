"""ORACLE (test infrastructure, never imported by loans_b200): the loader's frame path, restated in numpy.

What the reference's loader does with a decoded frame (common/datasets/image_dataset.py:16-28, 75-98, 147-182):

    pil_image = Image.fromarray(image.transpose(1, 2, 0).astype('uint8')).convert('RGB')
    pil_image = pil_image.resize((image_size[1], image_size[0]), Image.LANCZOS)         # resize_image, :16-28
    image = numpy.asarray(pil_image).transpose(2, 0, 1).astype(numpy.float32)
    return image / 255                                                                   # :98, :181

The arithmetic of ``Image.resize(..., LANCZOS)`` lives in Pillow (a dependency of the reference, requirements.txt), not in
the reference tree: ``ingest()`` restates Pillow's 8-bit resampling -- ``precompute_coeffs`` (float64 Lanczos-3 windowed
sinc, support 3 * max(scale, 1), coefficients normalised to sum 1), ``normalize_coeffs_8bpc`` (22-bit fixed point, rounded
half away from zero), horizontal pass then vertical pass, each accumulated in int32 from 1 << 21 and clipped to uint8 --
and is PINNED: tests/test_oracle_ingest.py checks it bit for bit against the real PIL (Pillow is importable in this image;
the outputs of the real ``Image.resize`` are also committed as tests/golden/ingest_lanczos.npz).  ``/ 255`` is numpy's
float32 true division.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _lanczos(x):
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a, b = x * math.pi, x / 3.0 * math.pi
        return (math.sin(a) / a) * (math.sin(b) / b if b != 0.0 else 1.0)
    return 0.0


def lanczos_coeffs(in_size, out_size):
    """(bounds[out_size, 2] = (first tap, tap count), kk[out_size, ksize] int32 fixed-point coefficients), as Pillow's
    precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = (0, in_size))."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            kk[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One resampling pass of a uint8 (H, W, C) image along ``axis`` (0 = vertical, 1 = horizontal)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)            # (n_in, other, C)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :n].astype(np.int64), src[lo:lo + n], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_lanczos_u8(img, out_h, out_w):
    """uint8 (H, W, C) -> uint8 (out_h, out_w, C): Pillow's ImagingResample for 8-bit images (horizontal pass first, each
    pass skipped when that size does not change)."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    if w != out_w:
        b, k = lanczos_coeffs(w, out_w)
        img = _pass(img, b, k, 1)
    if h != out_h:
        b, k = lanczos_coeffs(h, out_h)
        img = _pass(img, b, k, 0)
    return img


def ingest(frames_u8_hwc, image_size=None):
    """uint8 (B, H, W, 3) decoded frames -> float32 (B, 3, oH, oW) in [0, 1]: the loader's resize_image + ``/ 255``."""
    frames = np.ascontiguousarray(frames_u8_hwc, np.uint8)
    b, h, w, c = frames.shape
    oh, ow = (h, w) if image_size is None else (int(image_size[0]), int(image_size[1]))
    out = np.empty((b, c, oh, ow), np.float32)
    for i in range(b):
        r = resize_lanczos_u8(frames[i], oh, ow)
        out[i] = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    return out


def pil_reference(frames_u8_hwc, image_size):
    """The reference's own statement sequence run through the real PIL (used to pin ``ingest``; needs Pillow)."""
    from PIL import Image
    outs = []
    for f in frames_u8_hwc:
        image = f.transpose(2, 0, 1).astype(np.float32)                      # what the dataset hands to resize_image
        pil_image = Image.fromarray(image.transpose(1, 2, 0).astype('uint8'))
        pil_image = pil_image.convert('RGB')
        pil_image = pil_image.resize((image_size[1], image_size[0]), Image.LANCZOS)
        image = np.asarray(pil_image).transpose(2, 0, 1).astype(np.float32)
        outs.append(image / 255)
    return np.stack(outs).astype(np.float32)
