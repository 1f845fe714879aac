"""CPU tests of the host logic that turns the reference's three calls (rotation_dropout -> spatial_transformer_grid ->
spatial_transformer_sampler, sheep/sheep_localizer.py:61-63) into one fused launch per direction: which C-ABI entry
points are called, in which order, with which mask / flags -- against a recording stand-in for the library (no compute
without a GPU; the numbers are checked by the -m gpu tests)."""
import contextlib

import numpy as np
import pytest
import torch

import loans_b200
from loans_b200 import _lib
from loans_b200.functions import rotation_droput as RD
from loans_b200.functions import spatial_transformer as ST


class _Recorder(object):
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if not name.startswith("loans_stn_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            return 0
        return fn

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture
def rec(monkeypatch):
    r = _Recorder()
    monkeypatch.setattr(_lib, "lib", lambda: r)
    for mod in (ST, RD):
        monkeypatch.setattr(mod, "_need_cuda", lambda *t: None)
        monkeypatch.setattr(mod, "_stream", lambda: 0)
        monkeypatch.setattr(mod, "_on_device", lambda t: contextlib.nullcontext())
    return r


def _inputs(n=3, grad_x=False):
    torch.manual_seed(0)
    x = torch.rand(n, 3, 16, 16, requires_grad=grad_x)
    theta = torch.rand(n, 6, requires_grad=True)
    return x, theta


def _corner_loss(points):
    return points[:, :, 0, 0].sum() * 0.5 + points[:, :, 0, -1].sum() * 0.25 - points[:, :, -1, 0].sum()


def test_three_calls_make_one_fused_call_per_direction(rec):
    x, theta = _inputs()
    tp = RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0)
    points = ST.spatial_transformer_grid(tp, (5, 7))
    assert isinstance(tp, ST.Deferred) and isinstance(points, ST.Deferred)
    assert tuple(points.shape) == (3, 2, 5, 7) and points.dtype == torch.float32 and points.requires_grad
    assert rec.calls == []                                   # nothing launched yet
    rois = ST.spatial_transformer_sampler(x, points)
    assert rec.names() == ["loans_stn_crop_fwd_ex"]
    name, a = rec.calls[0]
    assert a[2] == 0.0 and a[6] == 0                          # mask01 = the dropout draw at ratio 0.0, no flags
    assert a[4] is not None and a[5] is None                  # dense grid written by the same kernel, no corners
    assert not points.pending and tp.pending                  # points now IS the fused kernel's grid; the masked theta never existed
    assert tuple(rois.shape) == (3, 3, 5, 7)
    ((rois * 2.0).sum() + _corner_loss(points)).backward()
    assert rec.names() == ["loans_stn_crop_fwd_ex", "loans_stn_crop_bwd_ex"]
    name, a = rec.calls[1]
    assert a[2] == 0.0 and a[4] is not None and a[5] is None  # gradient on points arrives as the dense upstream ggrid
    assert a[7] is None                                       # frames take no gradient: gx == NULL
    assert theta.grad is not None and tuple(theta.grad.shape) == (3, 6)


def test_no_gradient_on_points_means_null_ggrid(rec):
    x, theta = _inputs(grad_x=True)
    points = ST.spatial_transformer_grid(RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0), (4, 4))
    rois = ST.spatial_transformer_sampler(x, points)
    rois.sum().backward()
    name, a = rec.calls[-1]
    assert name == "loans_stn_crop_bwd_ex" and a[4] is None and a[5] is None and a[7] is not None
    assert rec.names() == ["loans_stn_crop_fwd_ex", "loans_stn_crop_bwd_ex"]


def test_touching_the_masked_theta_materialises_it(rec):
    x, theta = _inputs()
    tp = RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0)
    _ = tp + 1.0                                              # any torch function on it
    assert rec.names() == ["loans_stn_rotation_dropout"] and not tp.pending
    points = ST.spatial_transformer_grid(tp, (4, 4))          # theta is now an ordinary tensor: mask 1
    ST.spatial_transformer_sampler(x, points)
    assert rec.names()[-1] == "loans_stn_crop_fwd_ex" and rec.calls[-1][1][2] == 1.0


def test_touching_the_grid_first_keeps_the_gradient_on_the_grid(rec):
    x, theta = _inputs(grad_x=True)
    points = ST.spatial_transformer_grid(RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0), (4, 4))
    seen = []
    points.register_hook(lambda g: seen.append(tuple(g.shape)))       # materialises: dropout + grid kernels
    assert rec.names() == ["loans_stn_rotation_dropout", "loans_stn_grid_fwd"]
    rois = ST.spatial_transformer_sampler(x, points)
    assert rec.names()[-1] == "loans_stn_crop_fwd_ex"                 # still sampled from theta in registers ...
    assert rec.calls[-1][1][2] == 1.0 and rec.calls[-1][1][4] is None  # ... the masked one, grid not rewritten
    rois.sum().backward()
    names = rec.names()
    i = names.index("loans_stn_crop_bwd_ex")
    a = rec.calls[i][1]
    assert a[8] is not None and (a[9] & _lib.FLAG_UPRIGHT)              # per-pixel grid gradient out, upright hint
    assert names[i + 1:] == ["loans_stn_grid_bwd", "loans_stn_rotation_dropout"]
    assert seen == [(3, 2, 4, 4)]                                       # the hook on points saw the sampler's gradient


def test_grid_edits_take_the_explicit_sampler(rec):
    x, theta = _inputs()
    th = theta.detach().reshape(-1, 2, 3)
    for edit in ("inplace", "data", "derived"):
        rec.calls.clear()
        grid = ST.spatial_transformer_grid(th, (4, 4))
        if edit == "inplace":
            grid.mul_(0.5)
        elif edit == "data":
            grid.data[...] *= 0.5                             # does not move the version counter: the wrapper voids the note
        else:
            grid = grid * 0.5
        ST.spatial_transformer_sampler(x, grid)
        assert rec.names() == ["loans_stn_grid_fwd", "loans_stn_sampler_fwd"], edit


def test_theta_modified_in_place_between_the_calls_is_refused(rec):
    x, theta = _inputs()
    th = theta.detach().reshape(-1, 2, 3).clone()
    grid = ST.spatial_transformer_grid(th, (4, 4))
    th.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        ST.spatial_transformer_sampler(x, grid)
    th2 = theta.detach().reshape(-1, 2, 3).clone()
    tp = RD.rotation_dropout(th2, ratio=0.0)
    th2.add_(1.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        tp.sum()


def test_eager_mode_is_three_nodes(rec):
    x, theta = _inputs()
    with loans_b200.using_config("defer", False):
        tp = RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0)
        points = ST.spatial_transformer_grid(tp, (4, 4))
        rois = ST.spatial_transformer_sampler(x, points)
    assert not isinstance(points, ST.Deferred)
    assert rec.names() == ["loans_stn_rotation_dropout", "loans_stn_grid_fwd", "loans_stn_sampler_fwd"]
    (rois.sum() + _corner_loss(points)).backward()
    assert rec.names()[3:] == ["loans_stn_sampler_bwd", "loans_stn_grid_bwd", "loans_stn_rotation_dropout"]


def test_test_mode_forward_fuses_and_backward_raises_like_the_reference(rec):
    x, theta = _inputs()
    with loans_b200.using_config("train", False):
        points = ST.spatial_transformer_grid(RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.25), (4, 4))
        rois = ST.spatial_transformer_sampler(x, points)
    assert rec.names() == ["loans_stn_crop_fwd_ex"] and rec.calls[0][1][2] == 0.25     # test mode scales by ratio (:33-35)
    with pytest.raises(AttributeError, match="mask"):                                 # reference :47-48
        rois.sum().backward()


def test_train_mode_draw_is_one_per_call(rec):
    x, theta = _inputs()
    np.random.seed(3)
    draws = set()
    for _ in range(24):
        rec.calls.clear()
        points = ST.spatial_transformer_grid(RD.rotation_dropout(theta.detach().reshape(-1, 2, 3), ratio=0.5), (4, 4))
        ST.spatial_transformer_sampler(x, points)
        draws.add(rec.calls[0][1][2])
    assert draws == {0.0, 1.0}


def test_sampling_twice_from_one_grid(rec):
    x, theta = _inputs()
    points = ST.spatial_transformer_grid(theta.reshape(-1, 2, 3), (4, 4))
    r1 = ST.spatial_transformer_sampler(x, points)
    r2 = ST.spatial_transformer_sampler(x * 2.0, points)      # points is now the first fused node's grid output
    assert rec.names() == ["loans_stn_crop_fwd_ex", "loans_stn_sampler_fwd"]
    (r1.sum() + r2.sum()).backward()
    assert theta.grad is not None


def test_metadata_does_not_materialise(rec):
    _, theta = _inputs()
    g = ST.spatial_transformer_grid(RD.rotation_dropout(theta.reshape(-1, 2, 3), ratio=0.0), [6, 5])
    assert (g.shape[2], g.shape[3], g.size(0), g.dim(), g.ndim, len(g), g.numel()) == (6, 5, 3, 4, 4, 3, 180)
    assert g.dtype == torch.float32 and g.device.type == "cpu" and g.is_cuda is False and "pending" in repr(g)
    assert rec.calls == []
