"""GPU tests of the reference-facing operators: the three drop-in functions, their unfused kernels, autograd
wiring, and the reference's rotation-dropout golden vectors replayed on the device."""
import os

import numpy as np
import pytest

from loans_b200 import workloads as W
from oracle import stn_c as oc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available()
    return torch


def _t(T, a, grad=False):
    t = T.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.requires_grad_() if grad else t


def test_rotation_dropout_reference_golden_on_device(T):
    import loans_b200
    from loans_b200.functions import RotationDropout, rotation_dropout
    g = np.load(os.path.join(GOLDEN, "rotation_dropout.npz"))
    for i in range(int(g["n_cases"])):
        p = "c%03d_" % i
        train, ratio, seed = g[p + "meta"]
        with loans_b200.using_config("train", bool(train)):
            np.random.seed(int(seed))                 # the reference draws from numpy's global stream
            theta = _t(T, g[p + "theta"], grad=True)
            node = RotationDropout(ratio)
            y = node(theta)
            assert np.array_equal(y.detach().cpu().numpy(), g[p + "y"]), i
            if train:
                y.backward(_t(T, g[p + "gy"]))
                assert np.array_equal(theta.grad.cpu().numpy(), g[p + "gtheta"]), i
            else:
                with pytest.raises(AttributeError):     # reference :47-48 after a test-mode forward
                    y.backward(_t(T, g[p + "gy"]))
            np.random.seed(int(seed))
            assert np.array_equal(rotation_dropout(theta.detach(), ratio=ratio).cpu().numpy(), g[p + "y"])


def test_unfused_grid_and_sampler_kernels(T):
    from tests import gpu_util as G
    from loans_b200 import _lib
    rng = np.random.default_rng(8)
    b, c, h, w, oh, ow = 5, 3, 37, 29, 11, 14
    x = rng.random((b, c, h, w), dtype=np.float32)
    theta = W.make_theta(rng, b)
    L = _lib.lib()
    td = G.dev(theta)
    grid = T.empty((b, 2, oh, ow), device="cuda")
    _lib.check(L.loans_stn_grid_fwd(G.ptr(td), G.ptr(grid), b, oh, ow, G.stream()), "grid_fwd")
    grid0 = oc.grid_forward(theta, (oh, ow))
    assert np.array_equal(grid.cpu().numpy(), grid0)
    # arbitrary (non-affine) grid, reaching outside the frame
    garb = rng.uniform(-1.25, 1.25, (b, 2, oh, ow)).astype(np.float32)
    gy = rng.standard_normal((b, c, oh, ow), dtype=np.float32)
    xd, gd, gyd = G.dev(x), G.dev(garb), G.dev(gy)
    y = T.empty((b, c, oh, ow), device="cuda")
    _lib.check(L.loans_stn_sampler_fwd(G.ptr(xd), G.ptr(gd), G.ptr(y), b, 1, c, h, w, oh, ow, 0, G.stream()), "sampler_fwd")
    assert np.array_equal(y.cpu().numpy(), oc.sampler_forward(x, garb))
    gx = T.full((b, c, h, w), float("nan"), device="cuda")
    gg = T.empty((b, 2, oh, ow), device="cuda")
    _lib.check(L.loans_stn_sampler_bwd(G.ptr(xd), G.ptr(gd), G.ptr(gyd), G.ptr(gx), G.ptr(gg), b, 1, c, h, w, oh, ow, 0,
                                       G.stream()), "sampler_bwd")
    gx0, gg0 = oc.sampler_backward(x, garb, gy)
    assert np.array_equal(gg.cpu().numpy(), gg0)
    assert G.rel_max(gx.cpu().numpy(), gx0) <= 1e-5          # float atomics: order-dependent last bits
    gth = T.empty((b, 2, 3), device="cuda")
    _lib.check(L.loans_stn_grid_bwd(G.ptr(gg), G.ptr(gth), b, oh, ow, G.stream()), "grid_bwd")
    assert G.rel_max(gth.cpu().numpy(), oc.grid_backward(gg0)) <= 1e-5


def test_sampler_bwd_warp_aggregation_on_upsampling(T):
    # heavy up-sampling: many crop pixels of one warp hit the same source pixel
    from tests import gpu_util as G
    from loans_b200 import _lib
    rng = np.random.default_rng(9)
    b, c, h, w, oh, ow = 2, 3, 6, 5, 40, 48
    x = rng.random((b, c, h, w), dtype=np.float32)
    theta = np.tile(np.array([[0.4, 0.05, 0.1], [-0.03, 0.5, 0]], np.float32), (b, 1, 1))
    grid = oc.grid_forward(theta, (oh, ow))
    gy = rng.standard_normal((b, c, oh, ow), dtype=np.float32)
    xd, gd, gyd = G.dev(x), G.dev(grid), G.dev(gy)
    gx = T.full((b, c, h, w), float("nan"), device="cuda")
    _lib.check(_lib.lib().loans_stn_sampler_bwd(G.ptr(xd), G.ptr(gd), G.ptr(gyd), G.ptr(gx), None, b, 1, c, h, w, oh, ow, 0,
                                                G.stream()), "sampler_bwd")
    gx0, _ = oc.sampler_backward(x, grid, gy)
    assert G.rel_max(gx.cpu().numpy(), gx0) <= 1e-5


def test_drop_in_three_call_sequence_matches_fused_and_oracle(T):
    """The reference's call sequence (sheep/sheep_localizer.py:61-63) plus the corner regularisers' use of the
    grid (common/utils.py:152-157), through autograd."""
    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler, stn_crop
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=6, rotate=True)
    osz = (wl.out_h, wl.out_w)
    gy = _t(T, d["gy"])
    images = _t(T, d["x"])                                   # raw array: never requires grad in LoANs

    def corner_loss(points):
        return (points[:, :, 0, 0].sum() * 0.5 + points[:, :, 0, -1].sum() * 0.25 - points[:, :, -1, 0].sum())

    # (1) as the reference writes it
    theta1 = _t(T, d["theta"].reshape(-1, 6), grad=True)
    tp = rotation_dropout(theta1.reshape(-1, 2, 3), ratio=0.0)
    points = spatial_transformer_grid(tp, osz)
    rois = spatial_transformer_sampler(images, points)
    ((rois * gy).sum() + corner_loss(points)).backward()
    # (2) fused
    theta2 = _t(T, d["theta"], grad=True)
    rois2, points2 = stn_crop(images, theta2, osz, ratio=0.0)
    ((rois2 * gy).sum() + corner_loss(points2)).backward()
    assert T.equal(rois, rois2) and T.equal(points, points2)
    g1 = theta1.grad.reshape(-1, 2, 3).cpu().numpy()
    g2 = theta2.grad.cpu().numpy()
    # (3) oracle
    gg = np.zeros((6, 2) + osz, np.float32)
    gg[:, :, 0, 0] = 0.5
    gg[:, :, 0, -1] = 0.25
    gg[:, :, -1, 0] = -1.0
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    gt0, _, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], gg, 0.0)
    assert np.array_equal(rois.detach().cpu().numpy(), y0) and np.array_equal(points.detach().cpu().numpy(), grid0)
    sc = np.abs(gt0).max()
    assert np.abs(g1 - gt0).max() <= 1e-4 * sc and np.abs(g2 - gt0).max() <= 1e-4 * sc
    assert np.all(g1[:, 0, 1] == 0) and np.all(g1[:, 1, 0] == 0)           # ratio 0.0 cuts the rotation gradient


def test_sampler_falls_back_to_explicit_grid_when_grid_was_modified(T):
    from loans_b200.functions import spatial_transformer_grid, spatial_transformer_sampler
    rng = np.random.default_rng(2)
    x = rng.random((3, 3, 20, 20), dtype=np.float32)
    theta = W.make_theta(rng, 3)
    xt = _t(T, x, grad=True)
    grid = spatial_transformer_grid(_t(T, theta), (7, 7))
    grid2 = grid * 0.5                                                      # a different tensor: no origin note
    y = spatial_transformer_sampler(xt, grid2)
    g0 = oc.grid_forward(theta, (7, 7)) * np.float32(0.5)
    assert np.array_equal(y.detach().cpu().numpy(), oc.sampler_forward(x, g0))
    gy = rng.standard_normal((3, 3, 7, 7), dtype=np.float32)
    y.backward(_t(T, gy))
    gx0, _ = oc.sampler_backward(x, g0, gy)
    assert np.abs(xt.grad.cpu().numpy() - gx0).max() <= 1e-5 * np.abs(gx0).max()
    grid.mul_(0.5)                                                          # in-place edit: version counter moved
    y3 = spatial_transformer_sampler(xt.detach(), grid)
    assert np.array_equal(y3.cpu().numpy(), oc.sampler_forward(x, g0))


def test_gx_only_when_frames_require_grad(T):
    from loans_b200 import _lib
    from loans_b200.functions import stn_crop
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=2)
    th = _t(T, d["theta"], grad=True)
    xg = _t(T, d["x"], grad=True)
    rois, _ = stn_crop(xg, th, (75, 75))
    rois.backward(_t(T, d["gy"]))
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], (75, 75), d["gy"])
    assert np.abs(xg.grad.cpu().numpy() - gx0).max() <= 2e-6 * np.abs(gx0).max()
    n0 = _lib.launch_count()
    rois, _ = stn_crop(_t(T, d["x"]), th, (75, 75))
    rois.backward(_t(T, d["gy"]))
    assert _lib.launch_count() - n0 == 2                     # one fused kernel per direction


def test_type_checks_and_cpu_tensors_are_refused(T):
    from loans_b200.functions import InvalidType, spatial_transformer_grid, spatial_transformer_sampler, stn_crop
    with pytest.raises(RuntimeError):
        stn_crop(T.zeros(1, 3, 8, 8), T.zeros(1, 2, 3), (4, 4))
    x = T.zeros(2, 3, 8, 8, device="cuda")
    with pytest.raises(InvalidType):
        spatial_transformer_grid(T.zeros(2, 3, 2, device="cuda"), (4, 4))
    with pytest.raises(InvalidType):
        spatial_transformer_grid(T.zeros(2, 2, 3, device="cuda", dtype=T.float64), (4, 4))
    with pytest.raises(InvalidType):
        spatial_transformer_sampler(x, T.zeros(3, 2, 4, 4, device="cuda"))
    with pytest.raises(InvalidType):
        spatial_transformer_sampler(x, T.zeros(2, 3, 4, 4, device="cuda"))
    with pytest.raises(ValueError):
        spatial_transformer_sampler(x, T.zeros(2, 2, 4, 4, device="cuda"), use_cudnn=True)


def test_host_buffer_pipeline_matches_oracle(T):
    """HostCropPipeline: pinned host in / out, three streams; four different batches in flight through two buffer sets."""
    from loans_b200.pipeline import HostCropPipeline
    wl = W.WORKLOADS["cfg1"]
    osz = (wl.out_h, wl.out_w)
    pipe = HostCropPipeline(3, 3, wl.height, wl.width, osz, need_gx=True, depth=2)
    batches, outs = [], []
    for i in range(4):
        d = W.make_inputs(wl, seed=100 + i, batch=3)
        batches.append(d)
        h = {k: T.from_numpy(d[k]).pin_memory() for k in ("x", "theta", "gy")}
        o = {"y": T.empty((3, 3) + osz).pin_memory(), "grid": T.empty((3, 2) + osz).pin_memory(),
             "gtheta": T.empty((3, 2, 3)).pin_memory(), "gx": T.empty((3, 3, wl.height, wl.width)).pin_memory()}
        outs.append((h, o))
        pipe.submit(h["x"], h["theta"], h["gy"], o, mask01=0.0)
    pipe.drain()
    for d, (h, o) in zip(batches, outs):
        y0, g0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
        gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], None, 0.0)
        assert np.array_equal(o["y"].numpy(), y0) and np.array_equal(o["grid"].numpy(), g0)
        assert np.abs(o["gx"].numpy() - gx0).max() <= 2e-6 * np.abs(gx0).max()
        assert np.abs(o["gtheta"].numpy() - gt0).max() <= 1e-4 * np.abs(gt0).max()


# ------------------------------------------------------------------------------------------------ f1: prepare_images
def test_prepare_images_on_device_matches_oracle_and_pil_golden(T):
    """SheepLocalizer.prepare_images as one kernel: bit-exact against the oracle and the PIL-produced fixtures, both as
    the drop-in for the method (input already * 255) and with the multiply folded in; vector and scalar paths."""
    import os
    import torch
    from loans_b200.functions import prepare_images
    from oracle import stn_numpy as on
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "prepare_images.npz"))
    for i in range(int(g["n_cases"])):
        x = g["c%d_x" % i]
        xd = torch.from_numpy(x).cuda()
        assert np.array_equal(prepare_images(xd * 255).cpu().numpy(), g["c%d_out" % i])
        assert np.array_equal(prepare_images(xd, scale=255).cpu().numpy(), g["c%d_out" % i])
    rng = np.random.default_rng(5)
    for shp in ((4, 3, 224, 224), (2, 3, 75, 75), (3, 3, 33, 17), (1, 3, 512, 512)):
        x = rng.random(shp, dtype=np.float32)
        ref = on.prepare_images(x * 255)
        xd = torch.from_numpy(x).cuda()
        out = prepare_images(xd, scale=255)
        assert out.dtype == torch.float32 and not out.requires_grad
        assert np.array_equal(out.cpu().numpy(), ref)
        # an unaligned view: the scalar path
        big = torch.zeros(x.size + 1, dtype=torch.float32, device="cuda")
        view = big[1:].view(shp)
        view.copy_(xd)
        assert np.array_equal(prepare_images(view, scale=255).cpu().numpy(), ref)
    with pytest.raises(Exception):
        prepare_images(torch.zeros((1, 4, 8, 8), device="cuda"))
    with pytest.raises(Exception):
        prepare_images(torch.zeros((1, 3, 8, 8)))                 # CPU tensor: refused, no fallback


# ------------------------------------------------------------------------------------------------ f2: corner points
@pytest.mark.parametrize("name,batch,mask", [("cfg1", None, 0.0), ("cfg2", 8, 1.0), ("cfg3", 3, 0.0), ("cfg4", 2, 0.0)])
def test_corner_points_replace_the_dense_grid(T, name, batch, mask):
    """points='corners': (N,2,2,2), bit-identical to the dense grid's corner elements; a gradient arriving on them gives
    the gtheta the reference would get from the dense grid with that gradient at its corners; LoANs' consumers index it
    exactly as they index the dense grid."""
    from loans_b200.functions import stn_crop
    from oracle import stn_numpy as on
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True)
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    rng = np.random.default_rng(3)
    n = d["theta"].shape[0]
    gc = rng.standard_normal((n, 2, 2, 2)).astype(np.float32)
    x, th = _t(T, d["x"], grad=True), _t(T, d["theta"], grad=True)
    rois, corners = stn_crop(x, th, osz, mask01=mask, crops_per_frame=k, points="corners")
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, mask, k)
    assert corners.shape == (n, 2, 2, 2)
    assert np.array_equal(rois.detach().cpu().numpy(), y0)
    assert np.array_equal(corners.detach().cpu().numpy(), on.grid_corners(grid0))
    # the reference's consumers, written as they are there, on either array
    for pts, (hh, ww) in ((corners.detach().cpu().numpy(), (2, 2)), (grid0, osz)):
        tl = pts[:, 0, 0, 0], pts[:, 1, 0, 0]
        tr = pts[:, 0, 0, ww - 1]
        bl = pts[:, 1, hh - 1, 0]
        br = pts[:, 0, -1, -1], pts[:, 1, -1, -1]
        if pts.shape[2] == 2:
            got = (tl, tr, bl, br)
        else:
            want = (tl, tr, bl, br)
    for a, b in zip(got, want):
        for u, v in zip(np.atleast_2d(a), np.atleast_2d(b)):
            assert np.array_equal(u, v)
    T.autograd.backward([rois, corners], [_t(T, d["gy"]), _t(T, gc)])
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], on.corners_to_dense_ggrid(gc, *osz), mask, k)
    gt, gx = th.grad.cpu().numpy(), x.grad.cpu().numpy()
    assert np.abs(gt - gt0).max() <= 1e-4 * max(1.0, np.abs(gt0).max())
    assert np.abs(gx - gx0).max() <= 2e-6 * max(1.0, np.abs(gx0).max())


def test_corner_points_of_one_row_crops(T):
    from loans_b200.functions import stn_crop
    from oracle import stn_numpy as on
    rng = np.random.default_rng(8)
    for osz in ((1, 5), (6, 1), (1, 1)):
        x = rng.random((3, 3, 16, 20), dtype=np.float32)
        theta = W.make_theta(rng, 3)
        gy = rng.standard_normal((3, 3) + osz).astype(np.float32)
        gc = rng.standard_normal((3, 2, 2, 2)).astype(np.float32)
        xt, tt = _t(T, x), _t(T, theta, grad=True)
        rois, corners = stn_crop(xt, tt, osz, mask01=1.0, points="corners")
        _, grid0 = oc.crop_forward(x, theta, osz, 1.0, 1)
        assert np.array_equal(corners.detach().cpu().numpy(), on.grid_corners(grid0))
        T.autograd.backward([rois, corners], [_t(T, gy), _t(T, gc)])
        gt0, _, _ = oc.crop_backward(x, theta, osz, gy, on.corners_to_dense_ggrid(gc, *osz), 1.0, 1)
        assert np.abs(tt.grad.cpu().numpy() - gt0).max() <= 1e-4 * max(1.0, np.abs(gt0).max())


# ------------------------------------------------------------------------------------------------ f3: grayscale epilogue
@pytest.mark.parametrize("name,batch,mask,bf16", [("cfg1", None, 0.0, False), ("cfg2", 8, 1.0, False), ("cfg3", 4, 0.0, True),
                                                   ("cfg3", 3, 0.0, False), ("cfg4", 2, 0.0, False)])
def test_grayscale_epilogue_fused(T, name, batch, mask, bf16):
    """transform_rois_to_grayscale (sheep/sheep_localizer.py:65-68) inside the crop kernels: rois (N,1,oH,oW) bit-exact
    against the oracle's sampler followed by the reference's three statements; gradients equal to the oracle's backward fed
    coef * ggray per channel (general kernel, band kernel at cfg3, K = 16 at cfg4)."""
    from loans_b200.functions import stn_crop
    from oracle import stn_numpy as on
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=True)
    osz = (wl.out_h, wl.out_w)
    k = wl.crops_per_frame
    n = d["theta"].shape[0]
    rng = np.random.default_rng(12)
    gg = rng.standard_normal((n, 1) + osz).astype(np.float32)
    dt = T.bfloat16 if bf16 else T.float32
    x, th = _t(T, d["x"], grad=True), _t(T, d["theta"], grad=True)
    rois, points = stn_crop(x, th, osz, mask01=mask, crops_per_frame=k, out_dtype=dt, grayscale=True)
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, mask, k)
    gray0 = on.grayscale_forward(y0)
    assert rois.shape == (n, 1) + osz
    got = rois.detach().float().cpu().numpy()
    if bf16:
        ref = T.from_numpy(gray0).to(T.bfloat16).float().numpy()
        gg = T.from_numpy(gg).to(T.bfloat16).float().numpy()
    else:
        ref = gray0
    assert np.array_equal(got, ref)
    assert np.array_equal(points.detach().cpu().numpy(), grid0)
    rois.backward(_t(T, gg).to(dt))
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, on.grayscale_backward(gg), None, mask, k)
    assert np.abs(th.grad.cpu().numpy() - gt0).max() <= 1e-4 * max(1.0, np.abs(gt0).max())
    assert np.abs(x.grad.cpu().numpy() - gx0).max() <= 2e-6 * max(1.0, np.abs(gx0).max())


def test_grayscale_needs_three_channels(T):
    from loans_b200.functions import stn_crop
    x = T.zeros((1, 4, 8, 8), device="cuda")
    th = T.tensor([[[1.0, 0, 0], [0, 1.0, 0]]], device="cuda")
    with pytest.raises(Exception):
        stn_crop(x, th, (4, 4), grayscale=True)


def test_grayscale_epilogue_on_degenerate_transforms(T):
    """the per-frame-pixel gather fallback of the gx role (singular / wildly up-sampling transforms) behind the epilogue"""
    from loans_b200.functions import stn_crop
    from oracle import stn_numpy as on
    thetas = np.array([[[0.5, 0.25, 0], [1.0, 0.5, 0.1]], [[0, 0, 0.2], [0, 0, -0.3]], [[0.05, 0.01, 0.3], [-0.01, 0.06, -0.2]],
                       [[0.8, 0, 0], [0, 0.8, 0]], [[0.004, 0, 0.1], [0, 0.003, 0.2]]], np.float32)
    rng = np.random.default_rng(21)
    for shp, osz in (((5, 3, 24, 24), (9, 9)), ((5, 3, 16, 20), (12, 7))):
        x = rng.random(shp, dtype=np.float32)
        gg = rng.standard_normal((5, 1) + osz).astype(np.float32)
        for mask in (1.0, 0.0):
            xt, tt = _t(T, x, grad=True), _t(T, thetas, grad=True)
            rois, _ = stn_crop(xt, tt, osz, mask01=mask, grayscale=True)
            y0, _ = oc.crop_forward(x, thetas, osz, mask, 1)
            assert np.array_equal(rois.detach().cpu().numpy(), on.grayscale_forward(y0))
            rois.backward(_t(T, gg))
            gt0, gx0, _ = oc.crop_backward(x, thetas, osz, on.grayscale_backward(gg), None, mask, 1)
            assert np.abs(tt.grad.cpu().numpy() - gt0).max() <= 1e-4 * max(1.0, np.abs(gt0).max())
            assert np.abs(xt.grad.cpu().numpy() - gx0).max() <= 2e-6 * max(1.0, np.abs(gx0).max())


# ------------------------------------------------------------------------------------------------ round 2: deferral
def _corner_loss(points):
    return points[:, :, 0, 0].sum() * 0.5 + points[:, :, 0, -1].sum() * 0.25 - points[:, :, -1, 0].sum()


@pytest.mark.parametrize("grad_x", [False, True])
def test_three_reference_calls_are_two_launches_and_bitwise_the_fused_call(T, grad_x):
    """sheep/sheep_localizer.py:61-63 written out, with the corner regularisers' use of `points`
    (common/utils.py:152-157): ONE forward and ONE backward kernel, the ones stn_crop launches, same bits."""
    from loans_b200 import _lib
    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler, stn_crop
    wl = W.WORKLOADS["cfg2"]
    d = W.make_inputs(wl, batch=64, rotate=True)             # 64 crops: the automatic rule takes row bands when gx is wanted
    osz = (wl.out_h, wl.out_w)
    gy = _t(T, d["gy"])

    def run(three_calls):
        images = _t(T, d["x"], grad=grad_x)
        theta = _t(T, d["theta"].reshape(-1, 6), grad=True)
        n0 = _lib.launch_count()
        if three_calls:
            tp = rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0)
            points = spatial_transformer_grid(tp, osz)
            rois = spatial_transformer_sampler(images, points)
        else:
            rois, points = stn_crop(images, theta.reshape(-1, 2, 3), osz, ratio=0.0)
        fwd_kernel = _lib.last_kernel()
        n1 = _lib.launch_count()
        loss = (rois * gy).sum() + _corner_loss(points)      # torch's own kernels from here on: not counted
        loss.backward()
        T.cuda.synchronize()
        return (rois.detach(), points.detach(), theta.grad, images.grad, n1 - n0, _lib.launch_count() - n1, fwd_kernel,
                _lib.last_kernel())

    a, b = run(True), run(False)
    assert a[4] == 1 and a[5] == 1 and b[4] == 1 and b[5] == 1
    assert a[6] == b[6] == "stn_fwd_kernel"
    assert a[7] == b[7] == ("stn_bwd_band_kernel/row" if grad_x else "stn_bwd_theta_tab_kernel")
    for u, v in zip(a[:4], b[:4]):
        assert (u is None and v is None) or T.equal(u, v)
    gg = np.zeros((64, 2) + osz, np.float32)
    gg[:, :, 0, 0], gg[:, :, 0, -1], gg[:, :, -1, 0] = 0.5, 0.25, -1.0
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], gg, 0.0)
    assert np.array_equal(a[0].cpu().numpy(), y0) and np.array_equal(a[1].cpu().numpy(), grid0)
    assert np.abs(a[2].cpu().numpy().reshape(-1, 2, 3) - gt0).max() <= 1e-4 * np.abs(gt0).max()
    if grad_x:
        assert np.abs(a[3].cpu().numpy() - gx0).max() <= 2e-6 * np.abs(gx0).max()


def test_materialised_intermediates_give_the_same_numbers(T):
    """Looking at the masked theta or at the grid between the calls materialises them with their own kernels; the sampler then
    takes the gradient through `points` (hooks see it).  Same crops, same grid, gradients within the bars."""
    import loans_b200
    from loans_b200 import _lib
    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=8, rotate=True)
    osz = (wl.out_h, wl.out_w)
    gy = _t(T, d["gy"])
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    gg = np.zeros((8, 2) + osz, np.float32)
    gg[:, :, 0, 0], gg[:, :, 0, -1], gg[:, :, -1, 0] = 0.5, 0.25, -1.0
    gt0, gx0, ggrid0 = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], gg, 0.0)
    for mode in ("touch_theta", "touch_grid", "eager"):
        images = _t(T, d["x"], grad=True)
        theta = _t(T, d["theta"], grad=True)
        seen = []
        with loans_b200.using_config("defer", mode != "eager"):
            tp = rotation_dropout(theta, ratio=0.0)
            if mode == "touch_theta":
                assert float(tp.detach()[:, 0, 1].abs().max()) == 0.0
            points = spatial_transformer_grid(tp, osz)
            if mode == "touch_grid":
                points.register_hook(lambda g: seen.append(g.detach().clone()))
            rois = spatial_transformer_sampler(images, points)
        if mode == "touch_grid":
            assert _lib.last_kernel() == "stn_fwd_kernel"              # sampled from theta in registers, grid not re-read
        ((rois * gy).sum() + _corner_loss(points)).backward()
        assert np.array_equal(rois.detach().cpu().numpy(), y0), mode
        assert np.array_equal(points.detach().cpu().numpy(), grid0), mode
        assert np.abs(theta.grad.cpu().numpy() - gt0).max() <= 1e-4 * np.abs(gt0).max(), mode
        assert np.abs(images.grad.cpu().numpy() - gx0).max() <= 1e-5 * np.abs(gx0).max(), mode
        if mode == "touch_grid":
            assert len(seen) == 1 and np.array_equal(seen[0].cpu().numpy(), ggrid0 + gg)   # per-pixel grid gradient + the corners'


def test_grid_data_edit_between_the_calls_is_sampled_as_edited(T):
    from loans_b200.functions import spatial_transformer_grid, spatial_transformer_sampler
    rng = np.random.default_rng(4)
    x = rng.random((3, 3, 20, 20), dtype=np.float32)
    theta = W.make_theta(rng, 3)
    g0 = oc.grid_forward(theta, (7, 7)) * np.float32(0.5)
    for edit in ("data", "inplace"):
        grid = spatial_transformer_grid(_t(T, theta), (7, 7))
        if edit == "data":
            grid.data[...] *= 0.5                              # does not move the version counter
        else:
            grid.mul_(0.5)
        y = spatial_transformer_sampler(_t(T, x), grid)
        assert np.array_equal(y.cpu().numpy(), oc.sampler_forward(x, g0)), edit
    th = _t(T, theta)
    grid = spatial_transformer_grid(th, (7, 7))
    th.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        spatial_transformer_sampler(_t(T, x), grid)


def test_upright_hint_takes_the_axis_aligned_kernels_on_a_pre_masked_theta(T):
    """mask01 = 1 with a theta whose rotation terms are already zero (the materialised output of rotation_dropout): with
    LOANS_STN_FLAG_UPRIGHT the backward takes the band / table kernels and gives the bits of the mask01 = 0 call; a rotated
    crop among them is found on the device and handled by the general roles inside the same launch."""
    from tests import gpu_util as G
    from loans_b200 import _lib
    wl = W.WORKLOADS["cfg2"]
    d = W.make_inputs(wl, batch=64, rotate=True)
    masked = d["theta"].copy()
    masked[:, 0, 1] = 0
    masked[:, 1, 0] = 0
    osz = (wl.out_h, wl.out_w)
    L = _lib.lib()
    xd, gyd = G.dev(d["x"]), G.dev(d["gy"])
    n = 64

    def bwd(theta, mask, flags, need_gx=True):
        td = G.dev(theta)
        gt = T.empty((n, 2, 3), device="cuda")
        gx = T.empty_like(xd) if need_gx else None
        _lib.check(L.loans_stn_crop_bwd_ex(G.ptr(xd), G.ptr(td), float(mask), G.ptr(gyd), None, None, G.ptr(gt), G.ptr(gx), None, flags,
                                           n, 1, 3, wl.height, wl.width, osz[0], osz[1], _lib.F32, G.stream()), "bwd_ex")
        T.cuda.synchronize()
        return gt.cpu().numpy(), (None if gx is None else gx.cpu().numpy()), _lib.last_kernel()

    for need_gx, kernel in ((True, "stn_bwd_band_kernel/row"), (False, "stn_bwd_theta_tab_kernel")):
        gt_a, gx_a, k_a = bwd(d["theta"], 0.0, 0, need_gx)
        gt_b, gx_b, k_b = bwd(masked, 1.0, _lib.FLAG_UPRIGHT, need_gx)
        gt_c, gx_c, k_c = bwd(masked, 1.0, 0, need_gx)
        assert k_a == k_b == kernel and k_c in ("stn_bwd_kernel", "stn_bwd_theta_kernel")
        if need_gx:
            assert np.array_equal(gx_a, gx_b) and G.rel_max(gx_c, gx_a) <= 2e-6
        # d/d(theta01), d/d(theta10) are not masked when mask01 = 1: compare the four entries both calls define alike
        keep = np.array([[1, 0, 1], [0, 1, 1]], bool)
        assert np.array_equal(gt_a[:, keep], gt_b[:, keep]) and G.rel_max(gt_c, gt_b) <= 1e-4
        # a wrong hint: some crops ARE rotated -- still the oracle's numbers
        mixed = masked.copy()
        mixed[::3] = d["theta"][::3]
        gt_m, gx_m, k_m = bwd(mixed, 1.0, _lib.FLAG_UPRIGHT, need_gx)
        assert k_m == kernel
        gt0, gx0, _ = oc.crop_backward(d["x"], mixed, osz, d["gy"], None, 1.0)
        assert G.rel_max(gt_m, gt0) <= 1e-4
        if need_gx:
            assert G.rel_max(gx_m, gx0) <= 2e-6


def test_same_kernel_on_a_second_device_in_one_process(T):
    """The opt-in to > 48 KiB of dynamic shared memory is per device: the backward must work on cuda:1 after cuda:0."""
    if T.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from loans_b200.functions import stn_crop
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=4)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], (75, 75), d["gy"], None, 0.0)
    for dev in ("cuda:0", "cuda:1"):
        x = T.from_numpy(d["x"]).to(dev).requires_grad_()
        th = T.from_numpy(d["theta"]).to(dev).requires_grad_()
        rois, _ = stn_crop(x, th, (75, 75), ratio=0.0)
        rois.backward(T.from_numpy(d["gy"]).to(dev))
        T.cuda.synchronize(dev)
        assert np.abs(x.grad.cpu().numpy() - gx0).max() <= 2e-6 * np.abs(gx0).max(), dev
        assert np.abs(th.grad.cpu().numpy() - gt0).max() <= 1e-4 * np.abs(gt0).max(), dev


# ------------------------------------------------------------------------------------------------ round 2: conv-ready crops
@pytest.mark.parametrize("name,batch,mask,need_gx", [("cfg3", 5, 0.0, True), ("cfg3", 5, 0.0, False), ("cfg2", 64, 0.0, True),
                                                    ("cfg2", 9, 1.0, True), ("cfg1", 4, 1.0, False), ("cfg5", 12, 0.0, True)])
def test_channels_last_bf16_crops(T, name, batch, mask, need_gx):
    """LOANS_STN_FLAG_NHWC4: crops and their gradient as (N, oH, oW, 4) bf16 -- the values of the bf16
    NCHW crops, permuted, with a zero fourth channel; the backward reads gy in that layout and gives the bits of the NCHW
    call (same bf16 values in, same kernels).  Oracle: fp32 oracle -> bf16 round-to-nearest-even -> permute."""
    from tests import gpu_util as G
    from loans_b200 import _lib
    from loans_b200.functions import stn_crop
    wl = W.WORKLOADS[name]
    d = W.make_inputs(wl, batch=batch, rotate=mask != 0.0)
    osz = (wl.out_h, wl.out_w)
    n = batch
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, mask)
    x = _t(T, d["x"], grad=need_gx)
    th = _t(T, d["theta"], grad=True)
    rois, points = stn_crop(x, th, osz, mask01=mask, out_dtype=T.bfloat16, layout="nhwc4")
    assert tuple(rois.shape) == (n, osz[0], osz[1], 4) and rois.dtype == T.bfloat16
    got = rois.detach().float().cpu().numpy()
    assert np.array_equal(got[..., :3], np.transpose(G.bf16_round(y0), (0, 2, 3, 1)))
    assert np.all(got[..., 3] == 0) and not np.signbit(got[..., 3]).any()
    assert np.array_equal(points.detach().cpu().numpy(), grid0)
    # backward: gy arrives channels-last (the padding channel carries garbage on purpose: it must be ignored)
    gy_r = G.bf16_round(d["gy"])
    gy_cl = np.concatenate([np.transpose(gy_r, (0, 2, 3, 1)), np.full((n,) + osz + (1,), 123.0, np.float32)], axis=3)
    rois.backward(_t(T, gy_cl).to(T.bfloat16))
    k_nhwc = _lib.last_kernel()
    x2 = _t(T, d["x"], grad=need_gx)
    th2 = _t(T, d["theta"], grad=True)
    rois2, _ = stn_crop(x2, th2, osz, mask01=mask, out_dtype=T.bfloat16)
    rois2.backward(_t(T, gy_r).to(T.bfloat16))
    assert _lib.last_kernel() == k_nhwc                      # same dispatch as the planar call
    assert T.equal(th.grad, th2.grad)
    if need_gx:
        assert T.equal(x.grad, x2.grad)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, gy_r, None, mask)
    assert np.abs(th.grad.cpu().numpy() - gt0).max() <= 1e-4 * np.abs(gt0).max()
    if need_gx:
        assert np.abs(x.grad.cpu().numpy() - gx0).max() <= 2e-6 * np.abs(gx0).max()


def test_channels_last_needs_three_channel_bf16(T):
    from loans_b200.functions import InvalidType, stn_crop
    x = T.zeros(2, 3, 8, 8, device="cuda")
    th = T.zeros(2, 2, 3, device="cuda")
    with pytest.raises(InvalidType):
        stn_crop(x, th, (4, 4), layout="nhwc4")                                    # float32 crops
    with pytest.raises(InvalidType):
        stn_crop(T.zeros(2, 1, 8, 8, device="cuda"), th, (4, 4), out_dtype=T.bfloat16, layout="nhwc4")
    with pytest.raises(InvalidType):
        stn_crop(x, th, (4, 4), out_dtype=T.bfloat16, layout="nhwc4", grayscale=True)


# ------------------------------------------------------------------------------------------------ round 2: loader frame path
def test_frame_ingest_matches_pil_fixtures_and_oracle(T):
    """uint8 HWC frames -> float32 NCHW in [0,1] with the loader's LANCZOS resize (reference
    common/datasets/image_dataset.py:16-28,98) on the device, bit for bit what the real PIL produced (fixtures) and what
    the oracle restatement gives at BASELINE's frame sizes."""
    from loans_b200 import _lib
    from loans_b200.functions import FrameIngest, ingest_frames
    from oracle import ingest_numpy as ig
    g = np.load(os.path.join(GOLDEN, "ingest_lanczos.npz"))
    for i in range(int(g["n_cases"])):
        p = "c%02d_" % i
        out = ingest_frames(T.from_numpy(g[p + "frames"]).cuda(), tuple(int(v) for v in g[p + "size"]))
        assert np.array_equal(out.cpu().numpy(), g[p + "out_u8"].astype(np.float32) / 255), i
    rng = np.random.default_rng(3)
    for (b, h, w, size) in [(3, 384, 512, (224, 224)), (2, 512, 512, (224, 224)), (2, 224, 224, None), (2, 200, 300, (512, 512)),
                            (1, 1, 1, (4, 4)), (2, 75, 75, (75, 224)),
                            # vertical pass alone on the HWC frame (word-aligned rows, and not), conversion alone with an odd pixel count
                            (2, 96, 224, (50, 224)), (2, 96, 75, (50, 75)), (2, 75, 75, None),
                            # crop widths that are no multiple of four, odd frame widths (byte-wise staging), a 1080p frame (53 taps per
                            # output pixel: coefficient words read from the table, > 48 KiB of staged rows), a single output pixel
                            (1, 270, 480, (75, 75)), (2, 97, 301, (64, 70)), (1, 1080, 1920, (224, 224)), (1, 33, 47, (1, 1)),
                            (9, 40, 52, (24, 20))]:
        f = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
        n0 = _lib.launch_count()
        op = FrameIngest(b, (h, w), size)
        out = op(T.from_numpy(f).cuda())
        assert np.array_equal(out.cpu().numpy(), ig.ingest(f, size)), (b, h, w, size)
        both = size is not None and size[0] != h and size[1] != w
        assert _lib.launch_count() - n0 == (2 if both else 1)
        # a prepared ingest is graph-capturable: kernels only
        fd = T.from_numpy(f).cuda()
        od = T.empty_like(out)
        gr = T.cuda.CUDAGraph()
        with T.cuda.graph(gr):
            op(fd, out=od)
        gr.replay()
        T.cuda.synchronize()
        assert T.equal(od, out)
    with pytest.raises(Exception):
        ingest_frames(T.zeros(2, 8, 8, 3, device="cuda"), (4, 4))            # float frames are refused


def test_host_buffer_pipeline_with_decoded_uint8_frames(T):
    """HostCropPipeline(uint8_frames=True): the host uploads decoded (B,H,W,3) uint8 frames, `/ 255` and the NCHW layout happen
    on the device; frames take no gradient (what the LoANs step needs).  Same numbers as the oracle on x = u8 / 255."""
    from loans_b200.pipeline import HostCropPipeline
    wl = W.WORKLOADS["cfg1"]
    osz = (wl.out_h, wl.out_w)
    rng = np.random.default_rng(9)
    with HostCropPipeline(3, 3, wl.height, wl.width, osz, need_gx=False, depth=2, uint8_frames=True) as pipe:
        cases = []
        for i in range(3):
            d = W.make_inputs(wl, seed=200 + i, batch=3)
            u8 = rng.integers(0, 256, (3, wl.height, wl.width, 3), dtype=np.uint8)
            h = {"x": T.from_numpy(u8).pin_memory(), "theta": T.from_numpy(d["theta"]).pin_memory(), "gy": T.from_numpy(d["gy"]).pin_memory()}
            o = {"y": T.empty((3, 3) + osz).pin_memory(), "grid": T.empty((3, 2) + osz).pin_memory(), "gtheta": T.empty((3, 2, 3)).pin_memory()}
            cases.append((u8, d, h, o))
            pipe.submit(h["x"], h["theta"], h["gy"], o, mask01=0.0)
    for u8, d, h, o in cases:
        x = u8.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)          # reference image_dataset.py:98
        y0, g0 = oc.crop_forward(x, d["theta"], osz, 0.0)
        gt0, _, _ = oc.crop_backward(x, d["theta"], osz, d["gy"], None, 0.0)
        assert np.array_equal(o["y"].numpy(), y0) and np.array_equal(o["grid"].numpy(), g0)
        assert np.abs(o["gtheta"].numpy() - gt0).max() <= 1e-4 * np.abs(gt0).max()


def test_device_loader_reads_the_reference_formats(T, tmp_path):
    """SURVEY 8f rank 4, on-disk side: PNG frames named by `images.csv` / tab-separated `gt.csv` (reference
    common/datasets/image_dataset.py:48-98, :101-182) through loans_b200.datasets: frames bit for bit what the reference's
    per-sample sequence gives with the real PIL (decode -> tile -> resize_image -> / 255), boxes scaled as resize_bbox does."""
    from PIL import Image
    from loans_b200 import datasets as ds
    from oracle import ingest_numpy as ig
    rng = np.random.default_rng(21)
    root = str(tmp_path)
    specs = [("a.png", "RGB", (96, 128, 3)), ("b.png", "RGB", (96, 128, 3)), ("c.png", "L", (60, 80)), ("d.png", "RGBA", (96, 128, 4)),
             ("e.png", "RGB", (50, 50, 3))]
    raw = {}
    for name, mode, shape in specs:
        raw[name] = rng.integers(0, 256, shape, dtype=np.uint8)
        Image.fromarray(raw[name], mode).save(os.path.join(root, name))
    with open(os.path.join(root, "images.csv"), "w") as f:
        f.write("".join(n + "\n" for n, _, _ in specs))
    with open(os.path.join(root, "gt.csv"), "w") as f:
        f.write("a.png\t10\t20\t60\t100\nb.png\t0\t0\t96\t128\t5\t6\t50\t60\nc.png\t7\n")

    def reference(name, size):                                        # the reference's get_example, executed with the real PIL
        with Image.open(os.path.join(root, name)) as fh:
            image = np.asarray(fh, dtype=np.float32)
        if image.ndim == 2:
            image = image[:, :, None]
        image = image.transpose(2, 0, 1)
        if image.shape[0] == 1:
            image = np.tile(image, (3, 1, 1))
        hwc = np.asarray(Image.fromarray(image.transpose(1, 2, 0).astype('uint8')).convert('RGB'))
        if size is None:
            return hwc.transpose(2, 0, 1).astype(np.float32) / 255
        return ig.pil_reference(hwc[None], size)[0]

    size = (48, 56)
    d = ds.ImageDataset(os.path.join(root, "images.csv"), root=root, image_size=size)
    assert len(d) == 5
    batch = d.get_batch(range(5))                                      # three source sizes in one batch
    assert tuple(batch.shape) == (5, 3) + size and batch.is_cuda and batch.dtype == T.float32
    for i, (name, _, _) in enumerate(specs):
        assert np.array_equal(batch[i].cpu().numpy(), reference(name, size)), name
        assert np.array_equal(d[i].cpu().numpy(), reference(name, size)), name
    plain = ds.ImageDataset(["a.png", "e.png"], root=root)              # no resize: sizes differ -> a list
    out = plain.get_batch([0, 1])
    assert isinstance(out, list) and np.array_equal(out[1].cpu().numpy(), reference("e.png", None))
    lab = ds.LabeledImageDataset(os.path.join(root, "gt.csv"), root=root, image_size=size)
    frames, labels, scores = lab.get_batch([0, 1, 2])
    assert tuple(frames.shape) == (3, 3) + size and scores.shape == (3, 1)
    assert np.array_equal(frames[1].cpu().numpy(), reference("b.png", size))
    assert labels[0].dtype == np.int32 and labels[0].tolist() == [[5, 8, 30, 43]]          # 10*48/96, 20*56/128, 60/2, 100*.4375
    assert labels[1].tolist() == [[0, 0, 48, 56], [2, 2, 25, 26]] and labels[2].tolist() == [7]
    img, label, dummy = lab[0]
    assert np.array_equal(img.cpu().numpy(), reference("a.png", size)) and label.tolist() == labels[0].tolist() and dummy.shape == (1,)
    # the reference's iterator (train_sheep_localizer.py:113-116): loader threads decode, batches arrive on the device
    it = ds.MultithreadIterator(d, 2, repeat=False, shuffle=False, n_threads=3)
    got = [b for b in it]
    assert [tuple(b.shape) for b in got] == [(2, 3) + size, (2, 3) + size, (1, 3) + size] and it.epoch == 1
    assert np.array_equal(T.cat(got).cpu().numpy(), batch.cpu().numpy())
    it.finalize()


@pytest.mark.parametrize("variant", ["gray", "gray_bf16", "nhwc4", "bf16", "corners"])
def test_epilogues_through_the_row_owner_kernel(T, variant):
    """Several crops per frame with gx (stn_kframe.cu) behind every crop format of the `_ex` entry points: the grayscale epilogue
    (float32 and bf16 gy), bf16 crops planar and channels-last, and corner points with their upstream gradient -- forced on a
    small batch (the automatic rule takes the kernel from 16 frames), against the oracle."""
    from tests import gpu_util as G
    from loans_b200 import _lib
    from loans_b200.functions import stn_crop
    from oracle import stn_numpy as on
    wl = W.WORKLOADS["cfg4"]._replace(height=96, width=128, out_h=21, out_w=27)
    k = wl.crops_per_frame
    d = W.make_inputs(wl, batch=5, rotate=False)
    d["theta"][:, 0, 0] *= 1.3                                            # steps of >= 2 frame pixels: every frame is taken
    d["theta"][:, 1, 1] *= 1.3
    osz = (wl.out_h, wl.out_w)
    n = d["theta"].shape[0]
    rng = np.random.default_rng(3)
    x, th = _t(T, d["x"], grad=True), _t(T, d["theta"], grad=True)
    try:
        _lib.band_backward(True)
        if variant.startswith("gray"):
            dt = T.bfloat16 if variant == "gray_bf16" else T.float32
            gg = rng.standard_normal((n, 1) + osz).astype(np.float32)
            if dt == T.bfloat16:
                gg = G.bf16_round(gg)
            rois, _ = stn_crop(x, th, osz, mask01=0.0, crops_per_frame=k, out_dtype=dt, grayscale=True)
            rois.backward(_t(T, gg).to(dt))
            gy_eff, up = on.grayscale_backward(gg), None
        elif variant == "nhwc4":
            gy_r = G.bf16_round(d["gy"])
            gy_cl = np.concatenate([np.transpose(gy_r, (0, 2, 3, 1)), np.full((n,) + osz + (1,), -7.0, np.float32)], axis=3)
            rois, _ = stn_crop(x, th, osz, mask01=0.0, crops_per_frame=k, out_dtype=T.bfloat16, layout="nhwc4")
            rois.backward(_t(T, gy_cl).to(T.bfloat16))
            gy_eff, up = gy_r, None
        elif variant == "bf16":
            gy_r = G.bf16_round(d["gy"])
            rois, _ = stn_crop(x, th, osz, mask01=0.0, crops_per_frame=k, out_dtype=T.bfloat16)
            rois.backward(_t(T, gy_r).to(T.bfloat16))
            gy_eff, up = gy_r, None
        else:
            gc = rng.standard_normal((n, 2, 2, 2)).astype(np.float32)
            rois, pts = stn_crop(x, th, osz, mask01=0.0, crops_per_frame=k, points="corners")
            T.autograd.backward([rois, pts], [_t(T, d["gy"]), _t(T, gc)])
            up = np.zeros((n, 2) + osz, np.float32)
            up[:, :, 0, 0], up[:, :, 0, -1], up[:, :, -1, 0], up[:, :, -1, -1] = gc[:, :, 0, 0], gc[:, :, 0, 1], gc[:, :, 1, 0], gc[:, :, 1, 1]
            gy_eff = d["gy"]
        assert _lib.last_kernel() == "stn_bwd_theta_tab_kernel+stn_bwd_kframe_kernel"
    finally:
        _lib.band_backward(None)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, gy_eff, up, 0.0, k)
    assert np.abs(th.grad.cpu().numpy() - gt0).max() <= 1e-4 * max(1.0, np.abs(gt0).max())
    assert np.abs(x.grad.cpu().numpy() - gx0).max() <= 2e-6 * max(1.0, np.abs(gx0).max())


@pytest.mark.parametrize("uint8,need_gx,bf16", [(False, True, False), (True, False, False), (False, True, True)])
def test_host_buffer_pipeline_packed_transfers(T, uint8, need_gx, bf16):
    """HostCropPipeline.new_host_inputs / new_host_outputs: a step's inputs and results as views of one pinned buffer each -- one
    copy per direction -- give the bits of the per-tensor path (and of the oracle for crops and grid)."""
    from loans_b200.pipeline import HostCropPipeline
    wl = W.WORKLOADS["cfg1"]
    osz = (wl.out_h, wl.out_w)
    dt = T.bfloat16 if bf16 else T.float32
    rng = np.random.default_rng(17)
    res = {}
    for packed in (True, False):
        with HostCropPipeline(3, 3, wl.height, wl.width, osz, need_gx=need_gx, out_dtype=dt, depth=2, uint8_frames=uint8) as pipe:
            steps = []
            for i in range(3):
                d = W.make_inputs(wl, seed=300 + i, batch=3)
                xin = rng.integers(0, 256, (3, wl.height, wl.width, 3), dtype=np.uint8) if uint8 else d["x"]
                gy = T.from_numpy(d["gy"]).to(dt)
                if packed:
                    hin, out = pipe.new_host_inputs(), pipe.new_host_outputs()
                    hin["x"].copy_(T.from_numpy(xin)); hin["theta"].copy_(T.from_numpy(d["theta"])); hin["gy"].copy_(gy)
                    pipe.submit(None, None, None, out, mask01=0.0, inputs=hin)
                else:
                    hin = {"x": T.from_numpy(xin).pin_memory(), "theta": T.from_numpy(d["theta"]).pin_memory(), "gy": gy.pin_memory()}
                    out = {"y": T.empty((3, 3) + osz, dtype=dt).pin_memory(), "grid": T.empty((3, 2) + osz).pin_memory(),
                           "gtheta": T.empty((3, 2, 3)).pin_memory(), "gx": T.empty((3, 3, wl.height, wl.width)).pin_memory() if need_gx else None}
                    pipe.submit(hin["x"], hin["theta"], hin["gy"], out, mask01=0.0)
                steps.append((xin, d, hin, out))
            rng = np.random.default_rng(17) if packed else rng          # the same uint8 frames in both passes
        res[packed] = steps
    for (xin, d, _, a), (_, _, _, b) in zip(res[True], res[False]):
        for k in ("y", "grid", "gtheta") + (("gx",) if need_gx else ()):
            assert T.equal(a[k], b[k]), k
        x = xin.transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255) if uint8 else xin
        y0, g0 = oc.crop_forward(x, d["theta"], osz, 0.0)
        ref = T.from_numpy(y0).to(dt).float().numpy()
        assert np.array_equal(a["y"].float().numpy(), ref) and np.array_equal(a["grid"].numpy(), g0)
