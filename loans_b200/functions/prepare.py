"""``prepare_images`` -- the localizer's frame preparation on the device.

Mirrors ``SheepLocalizer.prepare_images`` / ``Resnet50SheepLocalizer.prepare_images`` (reference
sheep/sheep_localizer.py:72-82): the reference copies the whole batch to the host, loops over the images in Python
through ``chainer.links.model.vision.resnet.prepare(image, size=None)`` (uint8 quantisation, RGB -> BGR, mean
subtraction) and copies the result back.  Here it is one streaming CUDA kernel behind ``loans_stn_prepare_images``;
no host round trip, no CPU fallback.
"""
import torch

from loans_b200 import _lib
from loans_b200.functions.spatial_transformer import InvalidType, _expect, _need_cuda, _on_device, _ptr, _stream, resolve


def prepare_images(images, scale=1.0):
    """``images`` (B,3,H,W) float32 on a CUDA device -> (B,3,H,W) float32, BGR, mean-subtracted trunk input.

    ``scale=1.0``: ``images`` is what the reference passes to the method, ``images.copy() * 255`` (:45).
    ``scale=255``: ``images`` is the raw [0,1] frame batch; the multiply is folded into the kernel.
    The result carries no gradient, as in the reference (it rebuilds the arrays on the host).
    """
    images = resolve(images)
    _need_cuda(images)
    _expect(images.dtype == torch.float32, "images.dtype.char == 'f' (got %s)" % images.dtype)
    _expect(images.dim() == 4 and images.shape[1] == 3, "images.shape == (B, 3, H, W) (got %s)" % (tuple(images.shape),))
    x = images.detach().contiguous()
    b, c, h, w = x.shape
    out = torch.empty_like(x)
    with _on_device(x):
        _lib.check(_lib.lib().loans_stn_prepare_images(_ptr(x), float(scale), _ptr(out), b, c, h, w, _stream()),
                   "loans_stn_prepare_images")
    return out


__all__ = ["prepare_images", "InvalidType"]
