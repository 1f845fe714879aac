"""``rotation_dropout`` -- same module name (sic), names and semantics as reference functions/rotation_droput.py.

train mode (``config.train`` true): ONE Bernoulli draw per call for the whole batch, ``rand(1) < ratio``,
written into mask[:,0,1] and mask[:,1,0]  (reference :38-45) -- ``ratio`` is the probability of KEEPING the
rotation terms.  test mode: those two entries are scaled by ``ratio`` (:30-36).  Backward is ``gy * mask``
(:47-48).  LoANs calls it with ``ratio=0.0`` (sheep/sheep_localizer.py:61): the rotation terms are always
zeroed.  The draw comes from numpy's global stream on the host, as in the reference's CPU path; the
multiply runs on the device (and is folded into the fused kernel by ``stn_crop``).
"""
import numpy
import torch

from loans_b200 import _lib
from loans_b200.configuration import config
from loans_b200.functions.spatial_transformer import (Deferred, InvalidType, _modified_error, _need_cuda, _on_device, _ptr,
                                                       _stream, _unchanged, resolve)


def draw_mask_value(ratio):
    """The scalar that lands in mask[:,0,1] / mask[:,1,0] for this call (reference :33-35, :41-43)."""
    if not config.train:
        return float(ratio)
    return float(bool(numpy.random.rand(1)[0] < ratio))


class _RotationDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask_value, can_backprop):
        x = x.contiguous()
        y = torch.empty_like(x)
        with _on_device(x):
            _lib.check(_lib.lib().loans_stn_rotation_dropout(_ptr(x), mask_value, _ptr(y), x.shape[0], _stream()),
                       "loans_stn_rotation_dropout")
        ctx.meta = (mask_value, can_backprop)
        return y

    @staticmethod
    def backward(ctx, gy):
        mask_value, can_backprop = ctx.meta
        if not can_backprop:
            # reference :47-48 multiplies by self.mask, which a test-mode forward (:30-36) never creates
            raise AttributeError("'RotationDropout' object has no attribute 'mask' "
                                 "(backward after a test-mode forward, as in the reference)")
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        with _on_device(gy):
            _lib.check(_lib.lib().loans_stn_rotation_dropout(_ptr(gy), mask_value, _ptr(gx), gy.shape[0], _stream()),
                       "loans_stn_rotation_dropout")
        return gx, None, None


class RotationDropout(object):
    """Dropout regularisation for training the rotation of a spatial transformer (reference :9-48)."""

    def __init__(self, dropout_ratio):
        self.dropout_ratio = dropout_ratio

    def check_type_forward(self, x):
        if not (torch.is_tensor(x) and x.dtype.is_floating_point):
            raise InvalidType("x.dtype.kind == 'f'")
        if x.dim() != 3:
            raise InvalidType("x.ndim == 3 (got %d)" % x.dim())
        if x.shape[1] != 2 or x.shape[2] != 3:
            raise InvalidType("x.shape[1:] == (2, 3) (got %s)" % (tuple(x.shape[1:]),))
        if x.dtype != torch.float32:
            raise InvalidType("loans_b200 computes this path in float32 (got %s)" % x.dtype)

    def __call__(self, x):
        x = resolve(x)
        self.check_type_forward(x)
        _need_cuda(x)
        train = bool(config.train)
        if train:
            if not hasattr(self, "mask_value"):
                self.mask_value = draw_mask_value(self.dropout_ratio)       # one draw per call for the whole batch (:41)
            value = self.mask_value
        else:
            value = float(self.dropout_ratio)                               # :33-35
        if not config.defer:
            return _RotationDropoutFn.apply(x, value, train)
        # deferred: the grid node that normally follows hands (x, value) to the fused kernel and this multiply is never
        # launched on its own; anything else that touches the result materialises it with loans_stn_rotation_dropout
        note = {"theta": x, "theta_version": x._version, "theta_ptr": x.data_ptr(), "mask01": value, "can_backprop": train}

        def materialise():
            if not _unchanged(x, note["theta_version"], note["theta_ptr"]):
                raise _modified_error("the input of rotation_dropout")
            return _RotationDropoutFn.apply(x, value, train)

        return Deferred(x.shape, x.dtype, x.device, x.requires_grad, materialise, ("rotation_dropout", note))


def rotation_dropout(x, ratio=.5, **kwargs):
    return RotationDropout(ratio)(x)
