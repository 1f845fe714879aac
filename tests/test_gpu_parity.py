"""GPU parity: the CUDA path behind the C ABI vs the CPU oracle, on the same seeded inputs.

Bars (BASELINE.json north_star): forward crops within 1e-5 relative, theta / input gradients within 1e-4
relative.  What is actually asserted is tighter wherever the arithmetic allows it:
  * grid, crops (fp32) and the per-pixel grid gradient are BIT-EXACT against oracle/stn_oracle.c
    (same float32 rounding sequence; the kernels use _rn intrinsics where numpy rounds);
  * gx comes from a gather that adds the same products in a different order: <= 2e-6 of max|gx|;
  * gtheta is a 4096..5625-term float32 reduction vs the oracle's float64 sum: <= 1e-4 of max|gtheta|.
"""
import os

import numpy as np
import pytest

from loans_b200 import workloads as W
from oracle import stn_c as oc
from oracle import stn_numpy as on

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

FWD_TOL = 1e-5      # north_star, forward
GRAD_TOL = 1e-4     # north_star, gradients


@pytest.fixture(scope="module")
def G():
    import torch
    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    from tests import gpu_util
    return gpu_util


def _full_check(G, x, theta, osz, mask=1.0, k=1, seed=0, with_up=True):
    rng = np.random.default_rng(seed)
    n, c = theta.shape[0], x.shape[1]
    gy = rng.standard_normal((n, c) + tuple(osz), dtype=np.float32)
    gg = rng.standard_normal((n, 2) + tuple(osz), dtype=np.float32) if with_up else None
    y, grid = G.crop_fwd(x, theta, osz, mask, k)
    gt, gx, ggo = G.crop_bwd(x, theta, osz, gy, gg, mask, k)
    y0, grid0 = oc.crop_forward(x, theta, osz, mask, k)
    gt0, gx0, gg0 = oc.crop_backward(x, theta, osz, gy, gg, mask, k)
    assert np.array_equal(grid, grid0), "grid not bit-exact"
    assert np.array_equal(y, y0), "crops not bit-exact (max diff %g)" % np.abs(y - y0).max()
    assert np.array_equal(ggo, gg0), "ggrid not bit-exact"
    assert not np.isnan(gx).any() and not np.isnan(gt).any()
    assert G.rel_max(gx, gx0) <= 2e-6 <= GRAD_TOL, ("gx", G.rel_max(gx, gx0))
    assert np.abs(gt - gt0).max() <= GRAD_TOL * max(1.0, np.abs(gt0).max()), ("gtheta", np.abs(gt - gt0).max())
    return y, grid, gt, gx


@pytest.mark.parametrize("name,batch,mask", [
    ("cfg1", None, 0.0), ("cfg1", None, 1.0),
    ("cfg2", None, 1.0), ("cfg2", 16, 0.0),
    ("cfg3", 8, 0.0), ("cfg3", 4, 0.5),
    ("cfg4", 4, 0.0), ("cfg4", 2, 1.0),
    ("cfg5", 32, 0.0),
])
def test_workloads_against_oracle(G, name, batch, mask):
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True)
    _full_check(G, d["x"], d["theta"], (wl.out_h, wl.out_w), mask, wl.crops_per_frame)


def test_numpy_restatement_agrees_too(G):
    # the literal numpy restatement (BLAS order in the grid contraction): within the north-star bars
    wl = W.WORKLOADS["cfg2"]
    d = W.make_inputs(wl, batch=8)
    osz = (wl.out_h, wl.out_w)
    y, grid = G.crop_fwd(d["x"], d["theta"], osz)
    gt, gx, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"])
    y0, grid0 = on.crop_forward(d["x"], d["theta"], osz)
    gt0, gx0, _ = on.crop_backward(d["x"], d["theta"], osz, d["gy"])
    assert G.rel_max(grid, grid0) <= 2.5e-7
    assert G.rel_max(y, y0) <= FWD_TOL
    assert G.rel_max(gx, gx0) <= GRAD_TOL and G.rel_max(gt, gt0) <= GRAD_TOL


def test_bf16_crops_are_the_rounded_fp32_result(G):
    wl = W.WORKLOADS["cfg3"]
    d = W.make_inputs(wl, batch=6)
    osz = (wl.out_h, wl.out_w)
    y, grid = G.crop_fwd(d["x"], d["theta"], osz, 0.0, 1, bf16=True)
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    assert np.array_equal(grid, grid0)
    assert np.array_equal(y, G.bf16_round(y0))
    # backward consumes bf16 gy: identical to the oracle fed the same rounded gy
    gyr = G.bf16_round(d["gy"])
    gt, gx, ggo = G.crop_bwd(d["x"], d["theta"], osz, gyr, None, 0.0, 1, bf16=True)
    gt0, gx0, gg0 = oc.crop_backward(d["x"], d["theta"], osz, gyr, None, 0.0)
    assert np.array_equal(ggo, gg0)
    assert G.rel_max(gx, gx0) <= 2e-6 and G.rel_max(gt, gt0) <= GRAD_TOL


HARD_THETAS = {
    "identity": [[1, 0, 0], [0, 1, 0]],
    "flip_x": [[-0.8, 0, 0.1], [0, 0.7, 0]],
    "flip_both": [[-0.6, 0, 0], [0, -0.9, 0.05]],
    "rot90": [[0, 0.8, 0], [-0.8, 0, 0]],
    "rot45": [[0.5, -0.5, 0.1], [0.5, 0.5, -0.1]],
    "shear": [[0.7, 0.6, 0], [0, 0.5, 0]],
    "singular_rank1": [[0.5, 0.25, 0], [1.0, 0.5, 0.1]],
    "zero": [[0, 0, 0.2], [0, 0, -0.3]],
    "zero_x_only": [[0, 0, 0.2], [0, 0.8, 0]],
    "upsample_8x": [[0.05, 0.01, 0.3], [-0.01, 0.06, -0.2]],
    "huge_scale": [[40.0, 3.0, 0.5], [-2.0, 55.0, 0.1]],
    "far_outside": [[0.5, 0, 7.0], [0, 0.5, -9.0]],
    "half_outside": [[0.9, 0.1, 0.8], [0.05, 0.9, -0.7]],
}


@pytest.mark.parametrize("shape", [(3, 24, 24, 9, 9), (2, 17, 31, 12, 7), (1, 8, 8, 16, 16), (3, 20, 12, 1, 5),
                                   (4, 13, 9, 6, 1), (5, 33, 45, 20, 30)])
def test_hard_transforms_and_ragged_shapes(G, shape):
    c, h, w, oh, ow = shape
    names = sorted(HARD_THETAS)
    theta = np.array([HARD_THETAS[nm] for nm in names], np.float32)
    rng = np.random.default_rng(sum(shape))
    x = rng.random((len(names), c, h, w), dtype=np.float32)
    _full_check(G, x, theta, (oh, ow), 1.0, 1, seed=5)


def test_identity_at_full_frame_size(G):
    # BASELINE config 3's frame size, crop as large as the frame: bit-exact vs the oracle, and the frame itself up to
    # the float32 rounding of linspace (the identity grid does not land exactly on pixel centres at 512 points)
    x = np.random.default_rng(0).random((2, 3, 512, 512), dtype=np.float32)
    theta = np.tile(np.array([[1, 0, 0], [0, 1, 0]], np.float32), (2, 1, 1))
    y, grid = G.crop_fwd(x, theta, (512, 512))
    y0, grid0 = oc.crop_forward(x, theta, (512, 512))
    assert np.array_equal(grid, grid0) and np.array_equal(y, y0)
    assert np.abs(y - x).max() < 1e-4
    x8 = x[:, :, :9, :17].copy()                       # 9 x 17: every linspace point is exact -> exact copy
    y8, _ = G.crop_fwd(x8, theta, (9, 17))
    assert np.array_equal(y8, x8)


@pytest.mark.parametrize("name", ["cfg3", "cfg4"])
def test_full_size_adjoint_and_determinism(G, name):
    """<crop(x), gy> == <x, gx> (the gather really is the transpose of the sampler), at full size; and the
    fused path is bit-reproducible run to run (no atomics anywhere in it)."""
    import torch
    wl = W.WORKLOADS[name]
    batch = 64 if name == "cfg3" else 32
    d = W.make_inputs(wl, batch=batch, rotate=False)
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    y, _ = G.crop_fwd(d["x"], d["theta"], osz, 0.0, k, want_grid=False)
    gt, gx, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], None, 0.0, k, want_ggrid=False)
    lhs = float((y.astype(np.float64) * d["gy"]).sum())
    rhs = float((d["x"].astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-3, (lhs, rhs)
    gt2, gx2, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], None, 0.0, k, want_ggrid=False)
    assert np.array_equal(gx, gx2) and np.array_equal(gt, gt2)
    # oracle spot check on the first frames
    nb = 2
    gt0, gx0, _ = oc.crop_backward(d["x"][:nb], d["theta"][:nb * k], osz, d["gy"][:nb * k], None, 0.0, k)
    assert G.rel_max(gx[:nb], gx0) <= 2e-6 and G.rel_max(gt[:nb * k], gt0) <= GRAD_TOL


def test_shard_invariance_on_device(G):
    wl = W.WORKLOADS["cfg5"]
    d = W.make_inputs(wl, batch=16)
    osz = (wl.out_h, wl.out_w)
    y, grid = G.crop_fwd(d["x"], d["theta"], osz, 0.0)
    gt, gx, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], None, 0.0)
    for lo, hi in ((0, 8), (8, 16)):
        ys, gs = G.crop_fwd(d["x"][lo:hi], d["theta"][lo:hi], osz, 0.0)
        gts, gxs, _ = G.crop_bwd(d["x"][lo:hi], d["theta"][lo:hi], osz, d["gy"][lo:hi], None, 0.0)
        assert np.array_equal(ys, y[lo:hi]) and np.array_equal(gs, grid[lo:hi])
        assert np.array_equal(gts, gt[lo:hi]) and np.array_equal(gxs, gx[lo:hi])


def test_optional_outputs_and_empty_batch(G):
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=3)
    osz = (wl.out_h, wl.out_w)
    y, grid = G.crop_fwd(d["x"], d["theta"], osz, 1.0, 1, want_grid=False)
    assert grid is None
    y0, _ = oc.crop_forward(d["x"], d["theta"], osz)
    assert np.array_equal(y, y0)
    gt, gx, ggo = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], None, 1.0, 1, need_gx=False, want_ggrid=False)
    assert gx is None and ggo is None
    gt0, _, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"])
    assert np.abs(gt - gt0).max() <= GRAD_TOL * np.abs(gt0).max()
    # empty batch: nothing launched, nothing touched
    e = G.crop_fwd(np.zeros((0, 3, 8, 8), np.float32), np.zeros((0, 2, 3), np.float32), (4, 4))
    assert e[0].shape == (0, 3, 4, 4)


def test_golden_fixture_on_device(G):
    g = np.load(os.path.join(GOLDEN, "stn_small.npz"))
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        osz = tuple(int(v) for v in g[p + "out_size"])
        mask = float(g[p + "mask"])
        y, grid = G.crop_fwd(g[p + "x"], g[p + "theta"], osz, mask)
        gt, gx, ggo = G.crop_bwd(g[p + "x"], g[p + "theta"], osz, g[p + "gy"], g[p + "ggrid_up"], mask)
        assert np.array_equal(y, g[p + "y"]) and np.array_equal(grid, g[p + "grid"])
        assert np.array_equal(ggo, g[p + "ggrid"])
        assert G.rel_max(gx, g[p + "gx"]) <= 2e-6 and G.rel_max(gt, g[p + "gtheta"]) <= GRAD_TOL


@pytest.mark.parametrize("name,batch", [("cfg1", None), ("cfg2", 16), ("cfg3", 6), ("cfg4", 3)])
def test_axis_aligned_kernels_are_bitwise_the_general_kernels(G, name, batch):
    """The kernels written for axis-aligned crops (mask01 == 0: band / table-driven backward, and in a -DSTN_DEVEL build the
    TMA-staged forward) against the general kernels (LOANS_STN_CFG_FORCE_GENERAL) on the same inputs."""
    from loans_b200 import _lib
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True, with_ggrid=True)
    d["theta"][::5, :, 2] += 0.9                      # some crops hanging out of the frame
    d["theta"][1::7, 0, 0] *= -1.0                    # mirrored crops
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    try:
        _lib.force_general(True)
        y0, g0 = G.crop_fwd(d["x"], d["theta"], osz, 0.0, k)
        b0 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, k)
        assert _lib.last_kernel() == "stn_bwd_kernel"
    finally:
        _lib.force_general(False)
    n0 = _lib.launch_count()
    devel = _lib.lib().loans_stn_configure(_lib.CFG_TMA_FORWARD, 1) == 0
    try:
        y1, g1 = G.crop_fwd(d["x"], d["theta"], osz, 0.0, k)
        _lib.band_backward(True)
        b1 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, k)
        assert _lib.last_kernel() in (("stn_bwd_band_kernel/row", "stn_bwd_band_kernel/cta") if k == 1
                                      else ("stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel",))
    finally:
        _lib.band_backward(None)
        if devel:
            _lib.tma_forward(False)
    assert _lib.launch_count() - n0 == (2 if k == 1 else 3)          # several crops per frame: gtheta and gx are two kernels
    assert np.array_equal(y0, y1) and np.array_equal(g0, g1)
    assert np.array_equal(b0[2], b1[2])                                   # ggrid bit-exact
    assert G.rel_max(b1[1], b0[1]) <= 2e-6 and G.rel_max(b1[0], b0[0]) <= 1e-5
    yo, go = oc.crop_forward(d["x"], d["theta"], osz, 0.0, k)
    assert np.array_equal(y1, yo) and np.array_equal(g1, go)


@pytest.mark.parametrize("shape", [(3, 20, 24, 7, 9), (1, 8, 8, 16, 16), (2, 12, 20, 1, 5), (4, 16, 12, 6, 1), (3, 64, 2048, 5, 33)])
def test_axis_aligned_kernels_on_ragged_shapes(G, shape):
    c, h, w, oh, ow = shape
    rng = np.random.default_rng(sum(shape))
    theta = np.array([[[1, 0, 0], [0, 1, 0]], [[-0.8, 0.3, 0.1], [0.2, 0.7, 0]], [[0.05, 0, 0.3], [0, 0.06, -0.2]],
                      [[0.5, 0, 7.0], [0, 0.5, -9.0]], [[0.9, 0.1, 0.8], [0.05, -0.9, -0.7]], [[0, 0, 0.2], [0, 0, -0.3]],
                      [[40.0, 3.0, 0.5], [-2.0, 55.0, 0.1]]], np.float32)
    x = rng.random((len(theta), c, h, w), dtype=np.float32)
    from loans_b200 import _lib
    devel = _lib.lib().loans_stn_configure(_lib.CFG_TMA_FORWARD, 1) == 0      # the TMA-staged forward too, where it is built in
    try:
        _lib.band_backward(True)
        _full_check(G, x, theta, (oh, ow), 0.0, 1, seed=7)
    finally:
        _lib.band_backward(None)
        if devel:
            _lib.tma_forward(False)


# ------------------------------------------------------------------------------------------------ band backward
BAND_THETAS = np.array([
    [[1, 0, 0], [0, 1, 0]], [[0.8, 0, 0], [0, 0.8, 0]], [[0.6, 0, 0.3], [0, 0.7, -0.25]], [[0.9, 0, 0.8], [0, 0.9, -0.7]],
    [[0.5, 0, 7.0], [0, 0.5, -9.0]], [[0.5, 0, 0.0], [0, 0.5, -1.6]], [[1.7, 0, 0.1], [0, 2.5, -0.2]],
    [[40.0, 0, 0.5], [0, 55.0, 0.1]], [[0.9, 0, 0.0], [0, 0.15, 0.1]], [[0.2, 0, 0.3], [0, 0.25, -0.2]],
    [[0.05, 0, 0.3], [0, 0.06, -0.2]], [[-0.8, 0, 0.1], [0, 0.7, 0]], [[0, 0, 0.2], [0, 0.8, 0]],
    [[0.7, 0.4, 0], [-0.3, 0.6, 0.1]], [[0.3, 0, -0.5], [0, 0.33, 0.4]], [[0.55, 0, 0.05], [0, 0.52, 0.0]],
], np.float32)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])      # automatic, CTA bands (1, 2 pixels in flight), row bands
@pytest.mark.parametrize("shape", [(3, 24, 24, 9, 9), (3, 17, 32, 12, 7), (1, 8, 8, 16, 16), (3, 20, 12, 1, 5),
                                   (4, 13, 8, 6, 1), (3, 64, 48, 5, 33), (3, 40, 40, 37, 3), (3, 96, 128, 75, 75)])
def test_band_backward_hard_boxes(G, shape, variant):
    """The band kernel (mask01 == 0, one crop per frame) on boxes that exercise every branch of its plan: compact and
    dense tiles, halos, clipped rows, crops it declines (mirrored, 20x up-sampling, zero scale) -- against the oracle."""
    from loans_b200 import _lib
    c, h, w, oh, ow = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.random((len(BAND_THETAS), c, h, w), dtype=np.float32)
    try:
        _lib.band_backward(True)
        n0 = _lib.launch_count()
        for cs, rows in ((0, 0), (1, 1), (2, 3), (4, 0), (5, 0)):
            _lib.band_tuning(cs=cs, rows=rows, variant=variant)
            _full_check(G, x, BAND_THETAS, (oh, ow), 0.0, 1, seed=11)
        assert _lib.launch_count() - n0 == 10           # one forward + one backward launch per check
    finally:
        _lib.band_tuning()
        _lib.band_backward(None)


@pytest.mark.parametrize("name,batch", [("cfg1", None), ("cfg2", 16), ("cfg3", 6), ("cfg5", 24)])
def test_band_backward_matches_the_general_kernel(G, name, batch):
    from loans_b200 import _lib
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True, with_ggrid=True)
    d["theta"][::5, :, 2] += 0.9                      # some crops hanging out of the frame
    d["theta"][1::7, 0, 0] *= -1.0                    # mirrored crops: declined, general roles inside the band launch
    d["theta"][2::9, :, :2] *= 0.2                    # up-sampling crops: phased, dense tiles
    osz = (wl.out_h, wl.out_w)
    try:
        _lib.band_backward(True)
        n0 = _lib.launch_count()
        b1 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
        assert _lib.launch_count() - n0 == 1
        b2 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
        _lib.band_backward(False)
        b0 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
    finally:
        _lib.band_backward(None)
    assert np.array_equal(b0[2], b1[2])                                   # ggrid bit-exact
    assert G.rel_max(b1[1], b0[1]) <= 2e-6 and G.rel_max(b1[0], b0[0]) <= 1e-5
    assert np.array_equal(b1[0], b2[0]) and np.array_equal(b1[1], b2[1])   # bit-reproducible
    gt0, gx0, gg0 = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
    assert np.array_equal(b1[2], gg0)
    assert G.rel_max(b1[1], gx0) <= 2e-6 and G.rel_max(b1[0], gt0) <= GRAD_TOL


@pytest.mark.parametrize("cs", [1, 2, 3, 5])
def test_row_bands_with_few_ctas_per_crop(G, cs):
    """Row bands with 1 / 2 / 3 / 5 CTAs per crop: every warp then works through several crop rows, one after the other,
    on its private tile (at full size the launcher takes 2 CTAs per crop from 512 crops of 75 rows).  Same results,
    bit for bit, as with the automatic choice, and the oracle's within the usual bars."""
    from loans_b200 import _lib
    wl = W.WORKLOADS["cfg5"]
    d = W.make_inputs(wl, batch=12, rotate=True, with_ggrid=True, seed=31)
    d["theta"][::5, :, 2] += 0.9
    d["theta"][2::4, :, :2] *= 0.3                    # up-sampling crops: phased scatter, halo rows
    osz = (wl.out_h, wl.out_w)
    try:
        _lib.band_backward(True)
        _lib.band_tuning(variant=3)
        b0 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
        _lib.band_tuning(variant=3, cs=cs)
        b1 = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
    finally:
        _lib.band_tuning()
        _lib.band_backward(None)
    assert np.array_equal(b1[1], b0[1]) and np.array_equal(b1[2], b0[2])          # gx, ggrid: bit-identical
    assert G.rel_max(b1[0], b0[0]) <= 1e-5                                        # gtheta: the partial sums group differently
    gt0, gx0, gg0 = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, 1)
    assert np.array_equal(b1[2], gg0)
    assert G.rel_max(b1[1], gx0) <= 2e-6 and G.rel_max(b1[0], gt0) <= GRAD_TOL


def test_band_backward_full_size_adjoint(G):
    from loans_b200 import _lib
    wl = W.WORKLOADS["cfg2"]
    d = W.make_inputs(wl, rotate=False)
    osz = (wl.out_h, wl.out_w)
    y, _ = G.crop_fwd(d["x"], d["theta"], osz, 0.0, 1, want_grid=False)
    try:
        _lib.band_backward(True)
        gt, gx, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], None, 0.0, 1, want_ggrid=False)
    finally:
        _lib.band_backward(None)
    lhs = float((y.astype(np.float64) * d["gy"]).sum())
    rhs = float((d["x"].astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), 1.0) + 1e-3, (lhs, rhs)
    gt0, gx0, _ = oc.crop_backward(d["x"][:4], d["theta"][:4], osz, d["gy"][:4], None, 0.0, 1)
    assert G.rel_max(gx[:4], gx0) <= 2e-6 and G.rel_max(gt[:4], gt0) <= GRAD_TOL


# ------------------------------------------------------------------------------------------------ frames without grad
# ------------------------------------------------------------------------------------------------ several crops per frame
@pytest.mark.parametrize("rows", [0, 8, 16, 1000])         # frame rows per CTA: automatic, shortest, two passes per warp, whole frame
@pytest.mark.parametrize("shape,k", [((3, 24, 24, 9, 9), 4), ((3, 96, 128, 75, 75), 4), ((1, 40, 64, 5, 33), 2), ((4, 64, 48, 6, 70), 8),
                                     ((3, 32, 16, 1, 5), 4), ((3, 20, 12, 7, 1), 2)])
def test_kframe_backward_hard_boxes(G, shape, k, rows):
    """Several crops per frame with gx (stn_kframe.cu: warp-owned frame rows, inverse row map): frames whose crops all step by
    >= 2 frame pixels (taken), frames with an up-sampling / mirrored / rotated / zero-scale crop (declined: the general gx role
    in the same launch), boxes hanging out of the frame on every side, crops sharing frame rows -- against the oracle."""
    from loans_b200 import _lib
    c, h, w, oh, ow = shape
    rng = np.random.default_rng(sum(shape) + k)
    # frames 0..: all down-sampling boxes (taken); then the hard boxes of the band tests, k per frame (mostly declined)
    easy = []
    for _ in range(3 * k):
        s = rng.uniform(0.55, 1.2, 2) * np.array([max(2.2 * (ow - 1) / (w - 1), 0.3), max(2.2 * (oh - 1) / (h - 1), 0.3)])
        t = rng.uniform(-0.7, 0.7, 2)
        easy.append([[s[0], 0, t[0]], [0, s[1], t[1]]])
    theta = np.concatenate([np.array(easy, np.float32), BAND_THETAS[:(len(BAND_THETAS) // k) * k]])
    x = rng.random((len(theta) // k, c, h, w), dtype=np.float32)
    try:
        _lib.kframe_rows(rows)
        _lib.band_backward(True)                      # wherever it applies (the automatic rule takes it from 16 frames)
        _full_check(G, x, theta, (oh, ow), 0.0, k, seed=13)
        assert _lib.last_kernel() == "stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel"
    finally:
        _lib.band_backward(None)
        _lib.kframe_rows(0)


def test_kframe_backward_at_cfg4_size(G):
    """BASELINE config 4 with the automatic dispatch (mask 0, 16 jittered boxes per frame): 24 of its 128 frames at full frame and
    crop size, every frame against the C oracle; run-to-run bit reproducibility."""
    from loans_b200 import _lib
    wl = W.WORKLOADS["cfg4"]
    d = W.make_inputs(wl, batch=24, rotate=False, with_ggrid=True)
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    gt, gx, ggo = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, k)
    assert _lib.last_kernel() == "stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel"
    gt2, gx2, _ = G.crop_bwd(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, k)
    assert np.array_equal(gx, gx2) and np.array_equal(gt, gt2)
    gt0, gx0, gg0 = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0, k)
    assert np.array_equal(ggo, gg0)
    per_frame = np.abs(gx - gx0).reshape(24, -1).max(axis=1) / np.abs(gx0).reshape(24, -1).max(axis=1)
    assert per_frame.max() <= 2e-6, per_frame.max()
    sc = np.abs(gt0).reshape(24 * k, -1).max(axis=1)[:, None, None]
    assert (np.abs(gt - gt0) <= GRAD_TOL * sc).all()


@pytest.mark.parametrize("name,batch,bf16", [("cfg1", None, False), ("cfg2", 16, False), ("cfg3", 6, True), ("cfg4", 3, False), ("cfg5", 24, False)])
def test_theta_gradient_without_gx_table_kernel(G, name, batch, bf16):
    """gx == NULL with the rotation terms masked (every LoANs call): the table-driven theta kernel -- any number of crops per
    frame, mirrored and up-sampling boxes included -- against the oracle; the per-pixel grid gradient stays bit-exact."""
    from loans_b200 import _lib
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True, with_ggrid=True)
    d["theta"][::5, :, 2] += 0.9                      # hanging out of the frame
    d["theta"][1::7, 0, 0] *= -1.0                    # mirrored
    d["theta"][2::9, :, :2] *= 0.2                    # up-sampling
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    gy = G.bf16_round(d["gy"]) if bf16 else d["gy"]
    n0 = _lib.launch_count()
    gt, gx, ggo = G.crop_bwd(d["x"], d["theta"], osz, gy, d["ggrid"], 0.0, k, bf16=bf16, need_gx=False)
    assert _lib.launch_count() - n0 == 1 and gx is None
    gt0, _, gg0 = oc.crop_backward(d["x"], d["theta"], osz, gy, d["ggrid"], 0.0, k)
    assert np.array_equal(ggo, gg0)
    assert np.abs(gt - gt0).max() <= GRAD_TOL * max(1.0, np.abs(gt0).max())
    try:                                              # and the same through the general theta-only kernel
        _lib.band_backward(False)
        gt1, _, ggo1 = G.crop_bwd(d["x"], d["theta"], osz, gy, d["ggrid"], 0.0, k, bf16=bf16, need_gx=False)
    finally:
        _lib.band_backward(None)
    assert np.array_equal(ggo1, gg0) and G.rel_max(gt1, gt) <= 1e-5


def test_theta_gradient_without_gx_hard_boxes(G):
    rng = np.random.default_rng(4)
    for shape in ((3, 24, 24, 9, 9), (1, 8, 8, 16, 16), (3, 20, 12, 1, 5), (4, 13, 9, 6, 1), (3, 33, 45, 20, 30)):
        c, h, w, oh, ow = shape
        x = rng.random((len(BAND_THETAS), c, h, w), dtype=np.float32)
        gy = rng.standard_normal((len(BAND_THETAS), c, oh, ow)).astype(np.float32)
        gt, _, ggo = G.crop_bwd(x, BAND_THETAS, (oh, ow), gy, None, 0.0, 1, need_gx=False)
        gt0, _, gg0 = oc.crop_backward(x, BAND_THETAS, (oh, ow), gy, None, 0.0, 1)
        assert np.array_equal(ggo, gg0)
        assert np.abs(gt - gt0).max() <= GRAD_TOL * max(1.0, np.abs(gt0).max())


@pytest.mark.parametrize("name,batch,mask", [("cfg1", None, 0.0), ("cfg2", 8, 1.0)])
def test_reference_gpu_kernels_agree_with_oracle_and_cuda_path(G, name, batch, mask):
    """A second witness for the conventions the oracle restates (align-corners pixel map, zero padding, grid channel
    order, gradient scaling): cuDNN's cudnnSpatialTfGridGenerator / cudnnSpatialTfSampler, forward and backward -- the
    kernels chainer 4.1.0 itself runs for F.spatial_transformer_grid / _sampler on a GPU --
    driven through baseline/cudnn_stn.py on the same inputs.  cuDNN's arithmetic is not the numpy path's (fused
    multiply-adds, atomics for gx), so the bar is north_star's gradient tolerance, not bit-exactness."""
    import torch
    from baseline.cudnn_stn import CudnnStn
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, seed=5, batch=batch)
    x, theta, gy = d["x"], d["theta"], d["gy"]
    b, c, h, w = x.shape
    osz = (wl.out_h, wl.out_w)
    try:
        stn = CudnnStn(b, c, h, w, osz[0], osz[1], G.DEV)
    except RuntimeError as e:                              # no libcudnn on this box: nothing to compare with
        pytest.skip(str(e))
    stn.set_stream(torch.cuda.current_stream().cuda_stream)
    xd, td, gyd = G.dev(x), G.dev(theta), G.dev(gy)
    f32 = dict(dtype=torch.float32, device=G.DEV)
    y_c, grid2 = torch.empty((b, c) + osz, **f32), torch.empty((b,) + osz + (2,), **f32)
    grid_c, dgrid2 = torch.empty((b, 2) + osz, **f32), torch.empty((b,) + osz + (2,), **f32)
    gx_c, gt_c = torch.empty((b, c, h, w), **f32), torch.empty((b, 2, 3), **f32)
    stn.forward(xd, td, mask, y_c, grid2, grid_c)
    stn.backward(xd, None, mask, gyd, grid2, gx_c, dgrid2, gt_c)
    torch.cuda.synchronize()
    y, grid = G.crop_fwd(x, theta, osz, mask, 1)
    gt, gx, _ = G.crop_bwd(x, theta, osz, gy, None, mask, 1)
    y0, grid0 = oc.crop_forward(x, theta, osz, mask, 1)
    gt0, gx0, _ = oc.crop_backward(x, theta, osz, gy, None, mask, 1)
    for what, ours, orc, ref in (("grid", grid, grid0, grid_c), ("y", y, y0, y_c), ("gx", gx, gx0, gx_c), ("gtheta", gt, gt0, gt_c)):
        ref = ref.cpu().numpy()
        assert G.rel_max(orc, ref) <= GRAD_TOL, (what, "oracle vs cuDNN", G.rel_max(orc, ref))
        assert G.rel_max(ours, ref) <= GRAD_TOL, (what, "CUDA path vs cuDNN", G.rel_max(ours, ref))


# ------------------------------------------------------------------------------------------------ the benched dispatch, full size
@pytest.mark.parametrize("name,need_gx,kernel", [("cfg2", True, "stn_bwd_band_kernel/row"), ("cfg2", False, "stn_bwd_theta_tab_kernel"),
                                                 ("cfg5", True, "stn_bwd_band_kernel/row"), ("cfg5", False, "stn_bwd_theta_tab_kernel"),
                                                 ("cfg3", True, "stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel")])
def test_benched_dispatch_at_full_size_against_the_oracle(G, name, need_gx, kernel):
    """bench.py's headline (cfg2, batch 64) and its `configs` block (cfg5 batch 1024 -- two CTAs per crop --, cfg3) with the
    AUTOMATIC dispatch rule, as LoANs ships the path (mask 0): which kernel ran is asserted, and EVERY frame is compared with
    the C oracle (cfg3: the first 48 of its 256 frames go through the same launch geometry rule, crops x CTAs >= 200 x passes)."""
    from loans_b200 import _lib
    wl = W.WORKLOADS[name]
    batch = 48 if name == "cfg3" else wl.batch
    d = W.make_inputs(wl, batch=batch, rotate=False, with_ggrid=True)
    osz = (wl.out_h, wl.out_w)
    bf16 = wl.out_dtype == "bf16"
    gy = G.bf16_round(d["gy"]) if bf16 else d["gy"]
    y, grid = G.crop_fwd(d["x"], d["theta"], osz, 0.0, 1, bf16=bf16)
    assert _lib.last_kernel() == "stn_fwd_kernel"
    gt, gx, ggo = G.crop_bwd(d["x"], d["theta"], osz, gy, d["ggrid"], 0.0, 1, bf16=bf16, need_gx=need_gx)
    assert _lib.last_kernel() == kernel
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    gt0, gx0, gg0 = oc.crop_backward(d["x"], d["theta"], osz, gy, d["ggrid"], 0.0)
    assert np.array_equal(grid, grid0)
    assert np.array_equal(y, G.bf16_round(y0) if bf16 else y0)
    assert np.array_equal(ggo, gg0)
    sc = np.abs(gt0).reshape(batch, -1).max(axis=1)[:, None, None]
    assert (np.abs(gt - gt0) <= GRAD_TOL * sc).all()                       # per crop, against that crop's own largest entry
    if need_gx:
        per_frame = np.abs(gx - gx0).reshape(batch, -1).max(axis=1) / np.abs(gx0).reshape(batch, -1).max(axis=1)
        assert per_frame.max() <= 2e-6, per_frame.max()


def test_kframe_backward_random_shapes_against_the_general_kernel(G):
    """Random shapes (1 / 3 / 4 channels, odd crop sizes up to 128 columns, 1 ... 6 crops per frame, frames down to 8 x 8) and random
    upright boxes -- up- and down-sampling, mirrored, far outside the frame -- through the row-owner gx path (forced) and through
    the general kernel: gx within the summation-order bar, gtheta within the gradient bar, every element of gx written."""
    from loans_b200 import _lib
    rng = np.random.default_rng(2024)
    for case in range(40):
        c = int(rng.choice([1, 3, 4]))
        h, w = int(rng.integers(2, 40)) * 4, int(rng.integers(2, 40)) * 4
        oh, ow = int(rng.integers(1, 40)), int(rng.integers(1, 129))
        k = int(rng.integers(1, 7))
        frames = int(rng.integers(1, 5))
        n = frames * k
        theta = np.zeros((n, 2, 3), np.float32)
        theta[:, 0, 0] = rng.uniform(0.05, 1.4, n) * rng.choice([1, 1, 1, -1], n)
        theta[:, 1, 1] = rng.uniform(0.05, 1.4, n)
        theta[:, :, 2] = rng.uniform(-1.2, 1.2, (n, 2))
        if case % 3 == 0:                                  # frames whose crops all step by >= 2 pixels: the path's own kernel
            theta[:, 0, 0] = rng.uniform(2.2, 4.0, n) * max(ow - 1, 1) / (w - 1)
            theta[:, 1, 1] = rng.uniform(2.2, 4.0, n) * max(oh - 1, 1) / (h - 1)
        x = rng.random((frames, c, h, w), dtype=np.float32)
        gy = rng.standard_normal((n, c, oh, ow), dtype=np.float32)
        gg = rng.standard_normal((n, 2, oh, ow), dtype=np.float32)
        try:
            _lib.force_general(True)
            gt0, gx0, ggo0 = G.crop_bwd(x, theta, (oh, ow), gy, gg, 0.0, k)
        finally:
            _lib.force_general(False)
        try:
            _lib.band_backward(True)
            _lib.lib().loans_stn_configure(14, 1)          # LOANS_STN_CFG_KFRAME_SINGLE: also with one crop per frame
            gt1, gx1, ggo1 = G.crop_bwd(x, theta, (oh, ow), gy, gg, 0.0, k)
            assert _lib.last_kernel() == "stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel", (case, _lib.last_kernel())
        finally:
            _lib.lib().loans_stn_configure(14, 0)
            _lib.band_backward(None)
        assert not np.isnan(gx1).any(), case
        assert np.array_equal(ggo0, ggo1), case
        assert np.abs(gx1 - gx0).max() <= 2e-6 * max(np.abs(gx0).max(), 1e-30), (case, c, h, w, oh, ow, k)
        assert np.abs(gt1 - gt0).max() <= GRAD_TOL * max(1.0, np.abs(gt0).max()), case


@pytest.mark.parametrize("shape", [(16, 3, 1080, 1920, 75, 75), (2, 3, 2304, 4096, 33, 128), (1, 1, 4, 32768, 3, 9)])
def test_large_frames(G, shape):
    """HD and larger frames (the reference extracts 1920x1080 video frames before the loader shrinks them): whatever kernels the
    automatic dispatch picks at these row widths -- row buffers of 23 ... 48 KB per warp, or none that fit -- against the C oracle."""
    from loans_b200 import _lib
    b, c, h, w, oh, ow = shape
    rng = np.random.default_rng(b + h)
    x = rng.random((b, c, h, w), dtype=np.float32)
    theta = W.make_theta(rng, b, rotate=False)
    theta[:, 0, 1] = theta[:, 1, 0] = 0.0
    gy = rng.standard_normal((b, c, oh, ow), dtype=np.float32)
    y, grid = G.crop_fwd(x, theta, (oh, ow), 0.0, 1)
    gt, gx, _ = G.crop_bwd(x, theta, (oh, ow), gy, None, 0.0, 1)
    kern = _lib.last_kernel()
    y0, grid0 = oc.crop_forward(x, theta, (oh, ow), 0.0)
    gt0, gx0, _ = oc.crop_backward(x, theta, (oh, ow), gy, None, 0.0)
    assert np.array_equal(y, y0) and np.array_equal(grid, grid0), kern
    assert G.rel_max(gx, gx0) <= 2e-6, (kern, G.rel_max(gx, gx0))
    assert np.abs(gt - gt0).max() <= GRAD_TOL * max(1.0, np.abs(gt0).max()), kern
