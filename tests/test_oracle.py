"""CPU tests of the oracle itself: golden vectors, known answers, independent cross-checks.

The oracle (oracle/stn_numpy.py, oracle/stn_oracle.c) is the checker every GPU parity test leans on, so
it is pinned here first -- against the reference's own rotation-dropout file (golden vectors), against
analytic answers derivable from the reference (SURVEY.md 8c KATs), and against torch's independent
affine_grid / grid_sample.
"""
import os

import numpy as np
import pytest

from loans_b200 import workloads as W
from oracle import stn_c as oc
from oracle import stn_numpy as on

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class _SeededGlobalStream:
    """numpy's legacy global stream, seeded like the golden generator seeded it."""

    def __init__(self, seed):
        self.state = np.random.RandomState(int(seed))

    def rand(self, n):
        return self.state.rand(n)


# ----------------------------------------------------------------------------- a1 pinned to the reference
def test_rotation_dropout_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "rotation_dropout.npz"))
    n = int(g["n_cases"])
    assert n == 40
    seen_flags = set()
    for i in range(n):
        p = "c%03d_" % i
        train, ratio, seed = g[p + "meta"]
        train = bool(train)
        m = on.rotation_dropout_mask_value(ratio, train, _SeededGlobalStream(seed))
        seen_flags.add((train, float(m)))
        for impl in (on, oc):
            y = impl.rotation_dropout_forward(g[p + "theta"], m)
            assert y.dtype == np.float32
            assert np.array_equal(y, g[p + "y"]), (i, impl.__name__)
            if train:
                gt = impl.rotation_dropout_backward(g[p + "gy"], m)
                assert np.array_equal(gt, g[p + "gtheta"]), (i, impl.__name__)
    # both Bernoulli outcomes and the test-mode scaling were exercised
    assert (True, 0.0) in seen_flags and (True, 1.0) in seen_flags and (False, 0.25) in seen_flags


def test_rotation_dropout_ratio_zero_is_separable_in_both_modes():
    # LoANs always calls rotation_dropout(..., ratio=0.0) (sheep/sheep_localizer.py:61,169)
    theta = W.make_theta(np.random.default_rng(0), 5, rotate=True)
    for train in (True, False):
        m = on.rotation_dropout_mask_value(0.0, train, _SeededGlobalStream(7))
        y = on.rotation_dropout_forward(theta, m)
        assert np.all(y[:, 0, 1] == 0) and np.all(y[:, 1, 0] == 0)
        assert np.array_equal(y[:, [0, 1], [0, 1]], theta[:, [0, 1], [0, 1]])
        assert np.array_equal(y[:, :, 2], theta[:, :, 2])


def test_rotation_dropout_type_check():
    with pytest.raises(AssertionError):
        on.rotation_dropout_forward(np.zeros((2, 3, 2), np.float32), 1.0)


# ----------------------------------------------------------------------------- known answers
def _identity(b):
    return np.tile(np.array([[1, 0, 0], [0, 1, 0]], np.float32), (b, 1, 1))


@pytest.mark.parametrize("impl", [on, oc])
def test_identity_theta_reproduces_frame(impl):
    x = np.random.default_rng(1).random((2, 3, 11, 17), dtype=np.float32)
    y, grid = impl.crop_forward(x, _identity(2), (11, 17))
    assert np.array_equal(y, x)
    assert grid[0, 0, 0, 0] == -1 and grid[0, 0, 0, -1] == 1 and grid[0, 1, 0, 0] == -1 and grid[0, 1, -1, 0] == 1


@pytest.mark.parametrize("impl", [on, oc])
def test_constant_frame_and_affine_ramp(impl):
    h, w = 24, 31
    theta = np.array([[[0.6, 0.1, 0.05], [-0.08, 0.7, -0.1]]], np.float32)
    const = np.full((1, 1, h, w), 0.375, np.float32)
    y, grid = impl.crop_forward(const, theta, (9, 13))
    assert np.allclose(y, 0.375, rtol=0, atol=1e-7)
    jj, ii = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    ramp = (0.01 * jj + 0.02 * ii + 0.1).astype(np.float32)[None, None]
    y, grid = impl.crop_forward(ramp, theta, (9, 13))
    u = (grid[0, 0].astype(np.float64) + 1) * (w - 1) / 2
    v = (grid[0, 1].astype(np.float64) + 1) * (h - 1) / 2
    assert np.allclose(y[0, 0], 0.01 * u + 0.02 * v + 0.1, rtol=0, atol=2e-6)


@pytest.mark.parametrize("impl", [on, oc])
def test_box_outside_frame_is_zero_and_cuts_gradients(impl):
    x = np.random.default_rng(2).random((1, 3, 12, 12), dtype=np.float32) + 1.0
    theta = np.array([[[0.3, 0, 3.0], [0, 0.3, 0.0]]], np.float32)       # > 1 px right of the frame
    y, _ = impl.crop_forward(x, theta, (6, 6))
    assert np.all(y == 0)
    gy = np.ones_like(y)
    gtheta, gx, ggrid = impl.crop_backward(x, theta, (6, 6), gy)
    assert np.all(gx == 0) and np.all(ggrid[:, 0] == 0) and np.all(gtheta[:, 0] == 0)


@pytest.mark.parametrize("impl", [on, oc])
def test_loans_initial_theta_and_bbox_convention(impl):
    # sheep/sheep_localizer.py:30-33 bias [0.8,0,0,0,0.8,0]; :84-91 corner convention
    theta = np.tile(np.array([[0.8, 0, 0], [0, 0.8, 0]], np.float32), (3, 1, 1))
    grid = impl.grid_forward(theta, (75, 75))
    top, left = grid[:, 1, 0, 0], grid[:, 0, 0, 0]
    bottom, right = grid[:, 1, -1, -1], grid[:, 0, -1, -1]
    f = np.float32(0.8)
    assert np.all(top == -f) and np.all(left == -f) and np.all(bottom == f) and np.all(right == f)


def test_border_band_reads_zero_padding():
    # a sample between the last pixel and one pixel outside interpolates towards 0
    x = np.ones((1, 1, 4, 4), np.float32)
    grid = np.zeros((1, 2, 1, 1), np.float32)
    grid[0, 0] = 1.0 + (2.0 / 3.0) * 0.25            # u = 3.25 in unpadded pixels
    y = on.sampler_forward(x, grid)
    assert np.allclose(y, 0.75, atol=1e-6)
    assert np.array_equal(y, oc.sampler_forward(x, grid))


# ----------------------------------------------------------------------------- C restatement == numpy restatement
CASES = [
    # b, k, c, h, w, oh, ow, mask
    (3, 1, 3, 19, 23, 8, 11, 1.0),
    (2, 1, 1, 7, 5, 16, 12, 1.0),        # up-sampling, C=1
    (4, 1, 3, 32, 32, 10, 10, 0.0),      # separable (LoANs' case)
    (2, 4, 3, 20, 28, 6, 9, 0.0),        # K crops per frame
    (2, 3, 2, 15, 15, 5, 4, 0.5),        # test-mode scaling of the rotation terms, K=3
    (1, 1, 3, 6, 6, 1, 7, 1.0),          # oH == 1
]


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_is_bitwise_the_numpy_oracle(case):
    b, k, c, h, w, oh, ow, mask = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.random((b, c, h, w), dtype=np.float32)
    theta = W.make_theta(rng, b * k, rotate=True)
    theta[::3, :, 2] += 0.7                                   # some boxes partly outside
    gy = rng.standard_normal((b * k, c, oh, ow), dtype=np.float32)
    gg = rng.standard_normal((b * k, 2, oh, ow), dtype=np.float32)
    y1, g1 = on.crop_forward(x, theta, (oh, ow), mask, k)
    y2, g2 = oc.crop_forward(x, theta, (oh, ow), mask, k)
    assert np.array_equal(g1, g2) and np.array_equal(y1, y2)
    t1, gx1, gg1 = on.crop_backward(x, theta, (oh, ow), gy, gg, mask, k)
    t2, gx2, gg2 = oc.crop_backward(x, theta, (oh, ow), gy, gg, mask, k)
    assert np.array_equal(gg1, gg2) and np.array_equal(gx1, gx2)
    # gtheta goes through an sgemm over oH*oW terms in numpy and a float64 sum in C
    assert np.allclose(t1, t2, rtol=0, atol=2e-6 * max(1.0, np.abs(t2).max()))


def test_c_oracle_unfused_entry_points():
    rng = np.random.default_rng(5)
    x = rng.random((2, 3, 9, 14), dtype=np.float32)
    grid = rng.uniform(-1.3, 1.3, (2, 2, 5, 6)).astype(np.float32)      # arbitrary, not affine
    gy = rng.standard_normal((2, 3, 5, 6), dtype=np.float32)
    assert np.array_equal(on.sampler_forward(x, grid), oc.sampler_forward(x, grid))
    gx1, gg1 = on.sampler_backward(x, grid, gy)
    gx2, gg2 = oc.sampler_backward(x, grid, gy)
    assert np.array_equal(gx1, gx2) and np.array_equal(gg1, gg2)
    assert np.allclose(on.grid_backward(gg1), oc.grid_backward(gg1), rtol=0, atol=1e-5)


# ----------------------------------------------------------------------------- independent implementation (torch)
@pytest.mark.parametrize("smooth", [True, False])
def test_against_torch_affine_grid_and_grid_sample(smooth):
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    b, c, h, w, oh, ow = 4, 3, 40, 56, 17, 13
    x = W.make_frames(rng, b, c, h, w, smooth=smooth)
    theta = W.make_theta(rng, b, rotate=True)
    theta[0, 0, 2] = 0.55
    gy = rng.standard_normal((b, c, oh, ow), dtype=np.float32)
    y, grid = on.crop_forward(x, theta, (oh, ow))
    gtheta, gx, _ = on.crop_backward(x, theta, (oh, ow), gy)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tt = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    g = F.affine_grid(tt, (b, c, oh, ow), align_corners=True)
    yt = F.grid_sample(xt, g, mode="bilinear", padding_mode="zeros", align_corners=True)
    yt.backward(torch.tensor(gy, dtype=torch.float64))
    rel = lambda a, r: np.abs(a - r).max() / np.abs(r).max()            # noqa: E731
    assert rel(grid, g.detach().numpy().transpose(0, 3, 1, 2)) < 2e-7
    assert rel(y, yt.detach().numpy()) < (5e-6 if smooth else 1e-5)
    assert rel(gx, xt.grad.numpy()) < 1e-5
    assert rel(gtheta, tt.grad.numpy()) < (1e-5 if smooth else 1e-4)


def _forward_f64(x, theta, oh, ow):
    """The operator as the reference's call sites define it (sheep/sheep_localizer.py:62-63; SURVEY.md 8a: align-corners pixel
    map, zero padding), written independently of the oracle's statement sequence and in float64: for finite differences."""
    b, c, h, w = x.shape
    xs, ys = np.linspace(-1, 1, ow), np.linspace(-1, 1, oh)
    out = np.zeros((b, c, oh, ow))
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (1, 1), (1, 1)))
    for n in range(b):
        t = theta[n].astype(np.float64)
        gx_ = t[0, 0] * xs[None, :] + t[0, 1] * ys[:, None] + t[0, 2]
        gy_ = t[1, 0] * xs[None, :] + t[1, 1] * ys[:, None] + t[1, 2]
        u = np.clip((gx_ + 1) * (w - 1) / 2 + 1, 0, w + 1)
        v = np.clip((gy_ + 1) * (h - 1) / 2 + 1, 0, h + 1)
        u0 = np.clip(np.floor(u).astype(int), 0, w)
        v0 = np.clip(np.floor(v).astype(int), 0, h)
        fu, fv = u - u0, v - v0
        out[n] = (xp[n][:, v0, u0] * (1 - fu) * (1 - fv) + xp[n][:, v0, u0 + 1] * fu * (1 - fv) +
                  xp[n][:, v0 + 1, u0] * (1 - fu) * fv + xp[n][:, v0 + 1, u0 + 1] * fu * fv)
    return out


@pytest.mark.parametrize("impl", [on, oc])
def test_backward_is_the_derivative_of_the_forward(impl):
    """What chainer's own tests of these two functions check (gradient_check.check_backward): the analytic gradients against
    central differences of the forward -- here of an independent float64 forward, on smooth frames with every sample strictly
    inside the frame (the operator is piecewise smooth: kinks sit on pixel centres and on the border), for theta and, the
    operator being linear in x, as the exact adjoint identity <J d, gy> = <d, gx>."""
    rng = np.random.default_rng(77)
    b, c, h, w, oh, ow = 3, 3, 36, 44, 9, 11
    x = W.make_frames(rng, b, c, h, w, smooth=True)
    theta = np.zeros((b, 2, 3), np.float32)
    theta[:, 0, 0] = rng.uniform(0.4, 0.7, b)
    theta[:, 1, 1] = rng.uniform(0.4, 0.7, b)
    theta[:, 0, 1] = rng.uniform(-0.1, 0.1, b)
    theta[:, 1, 0] = rng.uniform(-0.1, 0.1, b)
    theta[:, :, 2] = rng.uniform(-0.15, 0.15, (b, 2))
    gy = rng.standard_normal((b, c, oh, ow)).astype(np.float32)
    gtheta, gx, _ = impl.crop_backward(x, theta, (oh, ow), gy)
    # theta: central differences of sum(gy * forward), step small against the distance to the next pixel centre for most samples
    eps = 1e-6
    num = np.zeros((b, 2, 3))
    for n in range(b):
        for idx in np.ndindex(2, 3):
            tp, tm = theta.astype(np.float64).copy(), theta.astype(np.float64).copy()
            tp[(n,) + idx] += eps
            tm[(n,) + idx] -= eps
            num[(n,) + idx] = ((_forward_f64(x[n:n + 1], tp[n:n + 1], oh, ow) - _forward_f64(x[n:n + 1], tm[n:n + 1], oh, ow)) * gy[n]).sum() / (2 * eps)
    assert np.abs(gtheta - num).max() <= 1e-4 * np.abs(num).max()
    # x: the forward is linear in x
    d = rng.standard_normal(x.shape)
    lhs = (_forward_f64(d, theta, oh, ow) * gy).sum()
    rhs = (d * gx).sum()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    # and the float64 forward itself agrees with the oracle's forward
    y, _ = impl.crop_forward(x, theta, (oh, ow))
    assert np.abs(y - _forward_f64(x, theta, oh, ow)).max() <= 1e-5


@pytest.mark.parametrize("impl", [on, oc])
def test_gradients_on_affine_ramp_are_analytic(impl):
    # x = a*col + b*row + c  =>  y = a*u + b*v + c  =>  dy/du = a, dy/dv = b wherever the box is inside
    h, w, oh, ow = 30, 40, 7, 9
    a, b_, c0 = 0.015625, -0.03125, 2.0                       # exactly representable slopes
    jj, ii = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    x = (a * jj + b_ * ii + c0).astype(np.float32)[None, None]
    theta = np.array([[[0.6, 0.15, 0.1], [-0.1, 0.5, -0.05]]], np.float32)
    gy = np.random.default_rng(4).standard_normal((1, 1, oh, ow), dtype=np.float32)
    gtheta, gx, ggrid = impl.crop_backward(x, theta, (oh, ow), gy)
    assert np.allclose(ggrid[0, 0], gy[0, 0] * a * (w - 1) / 2, rtol=0, atol=2e-5)
    assert np.allclose(ggrid[0, 1], gy[0, 0] * b_ * (h - 1) / 2, rtol=0, atol=2e-5)
    want = impl.grid_backward(np.stack([gy[0] * a * (w - 1) / 2, gy[0] * b_ * (h - 1) / 2], axis=1))
    assert np.allclose(gtheta, want, rtol=0, atol=2e-4)
    # bilinear weights sum to one: every gy lands somewhere in gx exactly once
    assert abs(float(gx.astype(np.float64).sum()) - float(gy.astype(np.float64).sum())) < 1e-4


# ----------------------------------------------------------------------------- committed fixtures
def test_stn_small_golden_replay():
    g = np.load(os.path.join(GOLDEN, "stn_small.npz"))
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        osz = tuple(int(v) for v in g[p + "out_size"])
        mask = float(g[p + "mask"])
        for impl in (on, oc):
            y, grid = impl.crop_forward(g[p + "x"], g[p + "theta"], osz, mask)
            assert np.array_equal(grid, g[p + "grid"]) and np.array_equal(y, g[p + "y"])
            gt, gx, gg = impl.crop_backward(g[p + "x"], g[p + "theta"], osz, g[p + "gy"], g[p + "ggrid_up"], mask)
            assert np.array_equal(gx, g[p + "gx"]) and np.array_equal(gg, g[p + "ggrid"])
            assert np.allclose(gt, g[p + "gtheta"], rtol=0, atol=2e-6 * max(1.0, np.abs(g[p + "gtheta"]).max()))


def test_shard_invariance():
    # the path shards by batch with no exchange: concat of shards == full batch
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=6, rotate=True)
    osz = (wl.out_h, wl.out_w)
    y, grid = oc.crop_forward(d["x"], d["theta"], osz)
    gt, gx, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"])
    for lo, hi in ((0, 3), (3, 6)):
        ys, gs = oc.crop_forward(d["x"][lo:hi], d["theta"][lo:hi], osz)
        gts, gxs, _ = oc.crop_backward(d["x"][lo:hi], d["theta"][lo:hi], osz, d["gy"][lo:hi])
        assert np.array_equal(ys, y[lo:hi]) and np.array_equal(gs, grid[lo:hi])
        assert np.array_equal(gts, gt[lo:hi]) and np.array_equal(gxs, gx[lo:hi])


# ------------------------------------------------------------------------------------------------ f1: prepare_images
def test_prepare_images_matches_the_pil_golden():
    """oracle/stn_numpy.prepare_images vs fixtures produced by the literal resnet.prepare statement sequence run
    through the real PIL (tests/golden/make_golden.py), quantisation boundaries included."""
    g = np.load(os.path.join(GOLDEN, "prepare_images.npz"))
    for i in range(int(g["n_cases"])):
        x = g["c%d_x" % i]
        out = on.prepare_images(x * 255)
        assert out.dtype == np.float32 and np.array_equal(out, g["c%d_out" % i])


def test_prepare_images_semantics():
    x = np.zeros((1, 3, 2, 2), np.float32)
    x[0, 0] = 1.0                                    # pure red frame
    out = on.prepare_images(x * 255)
    assert np.allclose(out[0, 2], 255 - 123.152) and np.allclose(out[0, 0], -103.063) and np.allclose(out[0, 1], -115.903)
    y = np.full((1, 3, 1, 1), 0.999, np.float32)     # 254.745 truncates to 254, it does not round to 255
    assert on.prepare_images(y * 255)[0, 0, 0, 0] == np.float32(254) - np.float32(103.063)


def test_grayscale_epilogue_semantics():
    """sheep/sheep_localizer.py:65-68: channel 0 is taken as b, channel 2 as r; the backward is the transpose."""
    rois = np.zeros((1, 3, 1, 2), np.float32)
    rois[0, 2, 0, 0] = 1.0                           # "r" = channel 2
    rois[0, 0, 0, 1] = 1.0                           # "b" = channel 0
    g = on.grayscale_forward(rois)
    assert g.shape == (1, 1, 1, 2) and g[0, 0, 0, 0] == np.float32(0.299) and g[0, 0, 0, 1] == np.float32(0.114)
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2, 3, 4, 5)).astype(np.float32)
    gg = rng.standard_normal((2, 1, 4, 5)).astype(np.float32)
    lhs = float((on.grayscale_forward(a).astype(np.float64) * gg).sum())
    rhs = float((a.astype(np.float64) * on.grayscale_backward(gg)).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
