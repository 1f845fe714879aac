"""Helpers for the -m gpu parity tests: numpy in, C-ABI call on cuda:0 (torch holds the memory), numpy out."""
import numpy as np
import torch

from loans_b200 import _lib

DEV = "cuda:0"


def dev(a, dtype=None):
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def crop_fwd(x, theta, osz, mask=1.0, k=1, bf16=False, want_grid=True):
    xd, td = dev(x), dev(theta)
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = osz
    y = torch.full((n, c, oh, ow), float("nan"), dtype=torch.bfloat16 if bf16 else torch.float32, device=DEV)
    grid = torch.full((n, 2, oh, ow), float("nan"), dtype=torch.float32, device=DEV) if want_grid else None
    _lib.check(_lib.lib().loans_stn_crop_fwd(ptr(xd), ptr(td), float(mask), ptr(y), ptr(grid), n, k, c, h, w, oh, ow,
                                             _lib.BF16 if bf16 else _lib.F32, stream()), "crop_fwd")
    torch.cuda.synchronize()
    return y.float().cpu().numpy(), (None if grid is None else grid.cpu().numpy())


def crop_bwd(x, theta, osz, gy, ggrid_up=None, mask=1.0, k=1, bf16=False, need_gx=True, want_ggrid=True):
    xd, td = dev(x), dev(theta)
    gyd = dev(gy, torch.bfloat16 if bf16 else torch.float32)
    ggd = dev(ggrid_up)
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = osz
    gt = torch.full((n, 2, 3), float("nan"), dtype=torch.float32, device=DEV)
    gx = torch.full((b, c, h, w), float("nan"), dtype=torch.float32, device=DEV) if need_gx else None
    ggo = torch.full((n, 2, oh, ow), float("nan"), dtype=torch.float32, device=DEV) if want_ggrid else None
    _lib.check(_lib.lib().loans_stn_crop_bwd(ptr(xd), ptr(td), float(mask), ptr(gyd), ptr(ggd), ptr(gt), ptr(gx), ptr(ggo),
                                             n, k, c, h, w, oh, ow, _lib.BF16 if bf16 else _lib.F32, stream()), "crop_bwd")
    torch.cuda.synchronize()
    return (gt.cpu().numpy(), None if gx is None else gx.cpu().numpy(), None if ggo is None else ggo.cpu().numpy())


def bf16_round(a):
    """round-to-nearest-even float32 -> bfloat16 -> float32, in numpy"""
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def rel_max(a, ref):
    return float(np.abs(a.astype(np.float64) - ref.astype(np.float64)).max() / max(float(np.abs(ref).max()), 1e-30))
