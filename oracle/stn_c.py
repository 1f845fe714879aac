"""ctypes front end of the plain-C oracle (oracle/stn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Same call signatures as the numpy restatement's composite functions, numpy arrays in and out.
Built by ``make -C oracle`` (``__graft_entry__.build()`` does that); ``load()`` builds it on demand
when gcc is around, which it is on both the CPU container and the GPU box.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_stn.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.c_int


def build(force=False):
    src = os.path.join(_HERE, "stn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_SO)
        lib.oracle_rotation_dropout.argtypes = [_f, ctypes.c_float, _f, _i]
        lib.oracle_grid_forward.argtypes = [_f, _f, _i, _i, _i]
        lib.oracle_grid_backward.argtypes = [_f, _f, _i, _i, _i]
        lib.oracle_sampler_forward.argtypes = [_f, _f, _f] + [_i] * 7
        lib.oracle_sampler_backward.argtypes = [_f, _f, _f, _f, _f] + [_i] * 7
        lib.oracle_crop_forward.argtypes = [_f, _f, ctypes.c_float, _f, _f] + [_i] * 7
        lib.oracle_crop_backward.argtypes = [_f, _f, ctypes.c_float, _f, _f, _f, _f, _f] + [_i] * 7
        for fn in ("oracle_rotation_dropout", "oracle_grid_forward", "oracle_grid_backward",
                   "oracle_sampler_forward", "oracle_sampler_backward", "oracle_crop_forward",
                   "oracle_crop_backward"):
            getattr(lib, fn).restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_f)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def rotation_dropout_forward(theta, mask_value):
    theta = _c(theta)
    out = np.empty_like(theta)
    load().oracle_rotation_dropout(_p(theta), float(mask_value), _p(out), theta.shape[0])
    return out


rotation_dropout_backward = rotation_dropout_forward       # same arithmetic (gy * mask)


def grid_forward(theta, output_shape):
    theta = _c(theta)
    oh, ow = output_shape
    grid = np.empty((theta.shape[0], 2, oh, ow), np.float32)
    load().oracle_grid_forward(_p(theta), _p(grid), theta.shape[0], oh, ow)
    return grid


def grid_backward(ggrid):
    ggrid = _c(ggrid)
    n, _, oh, ow = ggrid.shape
    gtheta = np.empty((n, 2, 3), np.float32)
    load().oracle_grid_backward(_p(ggrid), _p(gtheta), n, oh, ow)
    return gtheta


def sampler_forward(x, grid, crops_per_frame=1):
    x, grid = _c(x), _c(grid)
    b, c, h, w = x.shape
    n, _, oh, ow = grid.shape
    assert n == b * crops_per_frame
    y = np.empty((n, c, oh, ow), np.float32)
    load().oracle_sampler_forward(_p(x), _p(grid), _p(y), n, crops_per_frame, c, h, w, oh, ow)
    return y


def sampler_backward(x, grid, gy, crops_per_frame=1, need_gx=True, need_ggrid=True):
    x, grid, gy = _c(x), _c(grid), _c(gy)
    b, c, h, w = x.shape
    n, _, oh, ow = grid.shape
    assert n == b * crops_per_frame
    gx = np.empty_like(x) if need_gx else None
    ggrid = np.empty_like(grid) if need_ggrid else None
    load().oracle_sampler_backward(_p(x), _p(grid), _p(gy), _p(gx), _p(ggrid),
                                   n, crops_per_frame, c, h, w, oh, ow)
    return gx, ggrid


def crop_forward(x, theta, output_shape, mask_value=1.0, crops_per_frame=1):
    x, theta = _c(x), _c(theta)
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = output_shape
    assert n == b * crops_per_frame
    y = np.empty((n, c, oh, ow), np.float32)
    grid = np.empty((n, 2, oh, ow), np.float32)
    load().oracle_crop_forward(_p(x), _p(theta), float(mask_value), _p(y), _p(grid),
                               n, crops_per_frame, c, h, w, oh, ow)
    return y, grid


def crop_backward(x, theta, output_shape, gy, ggrid_upstream=None, mask_value=1.0, crops_per_frame=1,
                  need_gx=True):
    x, theta, gy, ggrid_upstream = _c(x), _c(theta), _c(gy), _c(ggrid_upstream)
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = output_shape
    assert n == b * crops_per_frame
    gtheta = np.empty((n, 2, 3), np.float32)
    gx = np.empty_like(x) if need_gx else None
    ggrid = np.empty((n, 2, oh, ow), np.float32)
    load().oracle_crop_backward(_p(x), _p(theta), float(mask_value), _p(gy), _p(ggrid_upstream),
                                _p(gtheta), _p(gx), _p(ggrid), n, crops_per_frame, c, h, w, oh, ow)
    return gtheta, gx, ggrid
