"""Synthetic workloads of the STN crop path: the five BASELINE.json configs and their input generator.

numpy only (no torch, no CUDA): shared by tests/, bench.py and __graft_entry__.smoke().  Shapes and
distributions are the ones SURVEY.md section 8(d) fixes:

* frames ``x``  : uniform [0,1) float32 NCHW, C=3 -- what ``common/datasets/image_dataset.py:75-98``
  (reference) feeds the localizer (pixel / 255), and the worst case for interpolation parity;
  ``smooth=True`` gives a low-frequency variant for tolerance analysis.
* ``theta``     : ``[[sx, r01, tx], [r10, sy, ty]]`` around the localizer's initial bias
  ``[0.8,0,0,0,0.8,0]`` (reference ``sheep/sheep_localizer.py:30-33``): s ~ U(0.5,0.9), t ~ U(-0.1,0.1),
  r ~ U(-0.2,0.2); 5 % of crops are pushed partly out of the frame (|t| up to 0.6 or s up to 1.3).
* ``gy``        : standard normal, shape of the crops.
"""
from collections import namedtuple

import numpy as np

Workload = namedtuple("Workload", "name batch crops_per_frame channels height width out_h out_w "
                                  "out_dtype rotation_ratio description")

# rotation_ratio: the ``ratio`` handed to rotation_dropout in front of the grid; None = no dropout node
# (theta used as is).  LoANs itself always passes 0.0 (sheep/sheep_localizer.py:61,169).
WORKLOADS = {
    "cfg1": Workload("cfg1", 4, 1, 3, 224, 224, 75, 75, "f32", 0.0,
                     "STN stage of the SheepLocalizer ResNet-18 train step, batch 4, 3x224x224 -> 75x75"),
    "cfg2": Workload("cfg2", 64, 1, 3, 224, 224, 64, 64, "f32", None,
                     "STN grid+sampler fwd+bwd only, batch 64, 3x224x224 -> 64x64, fp32"),
    "cfg3": Workload("cfg3", 256, 1, 3, 512, 512, 75, 75, "bf16", 0.0,
                     "rotation_dropout + STN sampler, batch 256, 3x512x512 -> 75x75, bf16 out"),
    "cfg4": Workload("cfg4", 128, 16, 3, 512, 512, 75, 75, "f32", 0.0,
                     "assessor feed: 16 jittered boxes per image, batch 128, 512x512 -> 75x75, fwd+bwd"),
    "cfg5": Workload("cfg5", 1024, 1, 3, 224, 224, 75, 75, "f32", 0.0,
                     "STN stage of the localizer+assessor step, global batch 1024 sharded over the GPUs"),
}


def algorithmic_bytes(wl, need_gx=True, batch=None):
    """(fwd_bytes, bwd_bytes) per step -- the SURVEY.md 8(d) / BASELINE.md formula, per crop times crops.

    fwd = 24 + 16*C*oH*oW + s_y*C*oH*oW + 8*oH*oW ; bwd = s_y*C*oH*oW + 16*C*oH*oW + 24 [+ 4*C*H*W/K].
    """
    b = wl.batch if batch is None else batch
    n = b * wl.crops_per_frame
    sy = 2 if wl.out_dtype == "bf16" else 4
    px = wl.out_h * wl.out_w
    fwd = 24 + 16 * wl.channels * px + sy * wl.channels * px + 8 * px
    bwd = sy * wl.channels * px + 16 * wl.channels * px + 24
    gx = 4 * wl.channels * wl.height * wl.width * b if need_gx else 0
    return n * fwd, n * bwd + gx


def make_theta(rng, n, rotate=True):
    s = rng.uniform(0.5, 0.9, (n, 2))
    t = rng.uniform(-0.1, 0.1, (n, 2))
    r = rng.uniform(-0.2, 0.2, (n, 2)) if rotate else np.zeros((n, 2))
    wild = rng.random(n) < 0.05
    pick = rng.random(n) < 0.5
    t_w = rng.uniform(-0.6, 0.6, (n, 2))
    s_w = rng.uniform(0.9, 1.3, (n, 2))
    t = np.where((wild & pick)[:, None], t_w, t)
    s = np.where((wild & ~pick)[:, None], s_w, s)
    theta = np.empty((n, 2, 3), np.float32)
    theta[:, 0, 0] = s[:, 0]
    theta[:, 0, 1] = r[:, 0]
    theta[:, 0, 2] = t[:, 0]
    theta[:, 1, 0] = r[:, 1]
    theta[:, 1, 1] = s[:, 1]
    theta[:, 1, 2] = t[:, 1]
    return theta


def make_frames(rng, b, c, h, w, smooth=False):
    if not smooth:
        return rng.random((b, c, h, w), dtype=np.float32)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    out = np.zeros((b, c, h, w), np.float32)
    for _ in range(4):
        fx = rng.uniform(0.5, 3.0, (b, c, 1, 1)).astype(np.float32) * (2 * np.pi / w)
        fy = rng.uniform(0.5, 3.0, (b, c, 1, 1)).astype(np.float32) * (2 * np.pi / h)
        ph = rng.uniform(0, 2 * np.pi, (b, c, 1, 1)).astype(np.float32)
        out += 0.125 * np.sin(fx * xx + fy * yy + ph)
    out += 0.5 + 0.05 * rng.random((b, c, h, w), dtype=np.float32)
    return out.astype(np.float32)


def make_inputs(wl, seed=1234, batch=None, smooth=False, rotate=None, with_ggrid=False):
    """dict(x, theta, gy[, ggrid]) of float32 numpy arrays for ``wl`` (optionally a smaller batch)."""
    rng = np.random.default_rng(seed)
    b = wl.batch if batch is None else batch
    k = wl.crops_per_frame
    n = b * k
    rotate = (wl.rotation_ratio is None) if rotate is None else rotate
    x = make_frames(rng, b, wl.channels, wl.height, wl.width, smooth=smooth)
    if k == 1:
        theta = make_theta(rng, n, rotate=rotate)
    else:
        base = make_theta(rng, b, rotate=rotate)
        theta = np.repeat(base, k, axis=0)
        js = rng.uniform(0.8, 1.25, (n, 2)).astype(np.float32)
        jt = rng.uniform(-0.15, 0.15, (n, 2)).astype(np.float32)
        theta[:, 0, 0] *= js[:, 0]
        theta[:, 1, 1] *= js[:, 1]
        theta[:, 0, 2] += jt[:, 0]
        theta[:, 1, 2] += jt[:, 1]
    gy = rng.standard_normal((n, wl.channels, wl.out_h, wl.out_w), dtype=np.float32)
    out = {"x": x, "theta": theta.astype(np.float32), "gy": gy}
    if with_ggrid:
        gg = np.zeros((n, 2, wl.out_h, wl.out_w), np.float32)
        # what the Direction/OutOfImage regularisers send back: gradients on three corner elements
        # (reference common/utils.py:152-157)
        for (i, j) in ((0, 0), (0, wl.out_w - 1), (wl.out_h - 1, 0)):
            gg[:, :, i, j] = rng.standard_normal((n, 2), dtype=np.float32)
        out["ggrid"] = gg
    return out
