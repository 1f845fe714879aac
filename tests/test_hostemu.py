"""CPU emulation of the kernels' per-pixel device code (loans_b200/csrc/stn_math.cuh) against the oracle.

The header is compiled for the host with g++ -ffp-contract=off and driven pixel by pixel by
tests/hostemu/stn_hostemu.cpp.  This checks, without a GPU, the float32 rounding chain of the forward, the
per-pixel backward and -- the delicate part -- the inverse-mapping gather that produces gx, including
rotations, flips, singular and wildly scaled transforms.  It is a test harness, not a product path.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from loans_b200 import workloads as W
from oracle import stn_c as oc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "stn_hostemu.cpp")
SO = os.path.join(HERE, "hostemu", "libstn_hostemu.so")
_f = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def emu():
    hdr = os.path.join(HERE, "..", "loans_b200", "csrc", "stn_math.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                        "-Wl,--unresolved-symbols=ignore-all", "-o", SO, SRC], check=True)
    lib = ctypes.CDLL(SO)
    lib.emu_crop_fwd.argtypes = [_f, _f, ctypes.c_float, _f, _f] + [ctypes.c_int] * 7
    lib.emu_crop_bwd.argtypes = [_f, _f, ctypes.c_float, _f, _f, _f, _f, _f] + [ctypes.c_int] * 7
    lib.emu_crop_fwd.restype = lib.emu_crop_bwd.restype = None
    lib.emu_gx_scatter.argtypes = [_f, ctypes.c_float, _f, _f] + [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_longlong)]
    lib.emu_gx_scatter.restype = ctypes.c_longlong
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_f)


def run_emu(lib, x, theta, osz, gy, gg, mask, k):
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = osz
    y = np.empty((n, c, oh, ow), np.float32)
    grid = np.empty((n, 2, oh, ow), np.float32)
    lib.emu_crop_fwd(_p(x), _p(theta), mask, _p(y), _p(grid), n, k, c, h, w, oh, ow)
    gt = np.empty((n, 2, 3), np.float32)
    gx = np.full_like(x, np.nan)
    ggo = np.empty_like(grid)
    lib.emu_crop_bwd(_p(x), _p(theta), mask, _p(gy), _p(gg), _p(gt), _p(gx), _p(ggo), n, k, c, h, w, oh, ow)
    return y, grid, gt, gx, ggo


def run_scatter(lib, x, theta, osz, gy, mask, k, tile_rows, tile_cols):
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = osz
    gx = np.full_like(x, np.nan)
    stats = (ctypes.c_longlong * 4)()
    conflicts = lib.emu_gx_scatter(_p(theta), mask, _p(gy), _p(gx), n, k, c, h, w, oh, ow, tile_rows, tile_cols, stats)
    return gx, conflicts, list(stats)


def check(lib, x, theta, osz, mask=1.0, k=1, seed=0):
    rng = np.random.default_rng(seed)
    n, c = theta.shape[0], x.shape[1]
    gy = rng.standard_normal((n, c) + tuple(osz), dtype=np.float32)
    gg = rng.standard_normal((n, 2) + tuple(osz), dtype=np.float32)
    y, grid, gt, gx, ggo = run_emu(lib, x, theta, osz, gy, gg, mask, k)
    y0, grid0 = oc.crop_forward(x, theta, osz, mask, k)
    gt0, gx0, gg0 = oc.crop_backward(x, theta, osz, gy, gg, mask, k)
    assert np.array_equal(grid, grid0), "grid not bit-exact"
    assert np.array_equal(y, y0), "crop not bit-exact"
    assert np.array_equal(ggo, gg0), "ggrid not bit-exact"
    assert not np.isnan(gx).any()
    sc = max(1.0, float(np.abs(gx0).max()))
    assert np.abs(gx - gx0).max() <= 2e-6 * sc, ("gx", np.abs(gx - gx0).max(), sc)
    sc = max(1.0, float(np.abs(gt0).max()))
    assert np.abs(gt - gt0).max() <= 1e-4 * sc, ("gtheta", np.abs(gt - gt0).max(), sc)
    # the tile-scatter formulation the GPU actually runs for gx, on two tilings: complete and race free
    h, w = x.shape[2:]
    sc = max(1.0, float(np.abs(gx0).max()))
    for tr, tc in ((8, min(w, 56)), (max(1, h // 2), max(4, (w // 3) & ~3)), (h, w)):
        gxs, conflicts, stats = run_scatter(lib, x, theta, osz, gy, mask, k, tr, tc)
        assert conflicts == 0, ("same-phase write conflicts", conflicts, stats)
        assert not np.isnan(gxs).any()
        assert np.abs(gxs - gx0).max() <= 2e-6 * sc, ("gx scatter", tr, tc, np.abs(gxs - gx0).max(), sc, stats)


@pytest.mark.parametrize("wl,batch,mask", [("cfg2", 3, 1.0), ("cfg1", 2, 0.0), ("cfg3", 1, 0.0), ("cfg4", 1, 0.0)])
def test_workload_shapes(emu, wl, batch, mask):
    wl = W.WORKLOADS[wl]
    d = W.make_inputs(wl, batch=batch, rotate=True)
    check(emu, d["x"], d["theta"], (wl.out_h, wl.out_w), mask, wl.crops_per_frame)


def _theta(rows):
    return np.array(rows, np.float32).reshape(-1, 2, 3)


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
    "point_on_pixel": [[0, 0, 0], [0, 0, 0]],
    "upsample_8x": [[0.05, 0.01, 0.3], [-0.01, 0.06, -0.2]],
    "huge_scale": [[40.0, 3.0, 0.5], [-2.0, 55.0, 0.1]],
    "far_outside": [[0.5, 0, 7.0], [0, 0.5, -9.0]],
    "edge_exact": [[1.0, 0, 2.0 / 23.0], [0, 1.0, 0]],
    "tiny_rotation": [[0.8, 1e-7, 0], [-1e-7, 0.8, 0]],
    "half_outside": [[0.9, 0.1, 0.8], [0.05, 0.9, -0.7]],
}


@pytest.mark.parametrize("name", sorted(HARD_THETAS))
@pytest.mark.parametrize("shape", [(24, 24, 9, 9), (17, 31, 12, 7), (8, 8, 16, 16), (20, 12, 1, 5), (13, 9, 6, 1)])
def test_hard_transforms(emu, name, shape):
    h, w, oh, ow = shape
    rng = np.random.default_rng(abs(hash((name, shape))) % (2 ** 31))
    x = rng.random((1, 2, h, w), dtype=np.float32)
    check(emu, x, _theta([HARD_THETAS[name]]), (oh, ow), 1.0, 1, seed=3)


def test_random_transforms_many(emu):
    rng = np.random.default_rng(99)
    for it in range(60):
        h, w = int(rng.integers(2, 40)), int(rng.integers(2, 40))
        oh, ow = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        k = int(rng.integers(1, 4))
        b = int(rng.integers(1, 3))
        c = int(rng.integers(1, 5))
        x = rng.random((b, c, h, w), dtype=np.float32)
        theta = rng.uniform(-1.5, 1.5, (b * k, 2, 3)).astype(np.float32)
        if it % 3 == 0:
            theta[:, :, :2] *= rng.uniform(0.01, 0.3)
        mask = [1.0, 0.0, 0.5][it % 3]
        check(emu, x, theta, (oh, ow), mask, k, seed=it)


def test_golden_fixture(emu):
    g = np.load(os.path.join(HERE, "golden", "stn_small.npz"))
    for i in range(int(g["n_cases"])):
        p = "c%d_" % i
        osz = tuple(int(v) for v in g[p + "out_size"])
        y, grid, gt, gx, ggo = run_emu(emu, g[p + "x"], g[p + "theta"], osz, g[p + "gy"], g[p + "ggrid_up"],
                                       float(g[p + "mask"]), 1)
        assert np.array_equal(y, g[p + "y"]) and np.array_equal(grid, g[p + "grid"]) and np.array_equal(ggo, g[p + "ggrid"])
        assert np.abs(gx - g[p + "gx"]).max() <= 2e-6 * max(1.0, np.abs(g[p + "gx"]).max())
        assert np.abs(gt - g[p + "gtheta"]).max() <= 1e-4 * max(1.0, np.abs(g[p + "gtheta"]).max())


# ------------------------------------------------------------------------------------------------ band backward
def run_band(lib, x, theta, osz, gy, gg, mask, cs, cap, max_rows=1 << 20):
    b, c, h, w = x.shape
    n = theta.shape[0]
    oh, ow = osz
    lib.emu_band_bwd.argtypes = [_f, _f, ctypes.c_float, _f, _f, _f, _f, _f, ctypes.POINTER(ctypes.c_int)] + \
        [ctypes.c_int] * 9 + [ctypes.POINTER(ctypes.c_longlong)]
    lib.emu_band_bwd.restype = ctypes.c_longlong
    gt = np.full((n, 2, 3), np.nan, np.float32)
    gx = np.full_like(x, np.nan)
    ggo = np.full((n, 2, oh, ow), np.nan, np.float32)
    ok = np.zeros(n, np.int32)
    stats = (ctypes.c_longlong * 4)()
    conflicts = lib.emu_band_bwd(_p(x), _p(theta), mask, _p(gy), _p(gg), _p(gt), _p(gx), _p(ggo),
                                 ok.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n, c, h, w, oh, ow, cs, cap, max_rows, stats)
    return gt, gx, ggo, ok.astype(bool), conflicts, list(stats)


def check_band(lib, x, theta, osz, mask=0.0, seed=0, expect_ok=None, layouts=((1, 4), (2, 6, 2), (3, 16), (8, 19), (5, 7, 1), (8, 2, 1))):
    rng = np.random.default_rng(seed)
    n, c = theta.shape[0], x.shape[1]
    gy = rng.standard_normal((n, c) + tuple(osz), dtype=np.float32)
    gg = rng.standard_normal((n, 2) + tuple(osz), dtype=np.float32)
    gt0, gx0, gg0 = oc.crop_backward(x, theta, osz, gy, gg, mask, 1)
    oks = None
    for lay in layouts:
        cs, cap = lay[:2]
        max_rows = lay[2] if len(lay) > 2 else 1 << 20
        cs = min(cs, osz[0])
        gt, gx, ggo, ok, conflicts, stats = run_band(lib, x, theta, osz, gy, gg, mask, cs, cap, max_rows)
        assert conflicts == 0, ("same-phase write conflicts", conflicts, stats, cs, cap)
        assert stats[0] == 0, ("gx elements not written exactly once", stats, cs, cap)
        assert stats[3] == 0, ("tile slot out of range", stats, cs, cap)
        if expect_ok is not None:
            assert list(ok) == list(expect_ok), (ok, expect_ok)
        if not ok.any():
            continue
        assert np.array_equal(ggo[ok], gg0[ok]), "ggrid not bit-exact"
        assert not np.isnan(gx[ok]).any()
        sc = max(1.0, float(np.abs(gx0).max()))
        assert np.abs(gx[ok] - gx0[ok]).max() <= 2e-6 * sc, ("gx", np.abs(gx[ok] - gx0[ok]).max(), sc, cs, cap)
        sc = max(1.0, float(np.abs(gt0).max()))
        assert np.abs(gt[ok] - gt0[ok]).max() <= 1e-4 * sc, ("gtheta", np.abs(gt[ok] - gt0[ok]).max(), sc)
        oks = ok
    return oks


@pytest.mark.parametrize("wl,batch", [("cfg2", 6), ("cfg1", 4), ("cfg3", 2), ("cfg5", 4)])
def test_band_backward_on_workload_shapes(emu, wl, batch):
    wl = W.WORKLOADS[wl]
    d = W.make_inputs(wl, batch=batch, rotate=True)
    ok = check_band(emu, d["x"], d["theta"], (wl.out_h, wl.out_w), 0.0, layouts=((8, 16), (8, 18), (4, 12), (1, 6), (8, 2, 1)))
    assert ok is not None and ok.all()            # every synthetic LoANs crop is a down-sampling, upright box


BAND_THETAS = {
    "identity": [[1, 0, 0], [0, 1, 0]],
    "loans_init": [[0.8, 0, 0], [0, 0.8, 0]],
    "shifted": [[0.6, 0, 0.3], [0, 0.7, -0.25]],
    "half_outside": [[0.9, 0, 0.8], [0, 0.9, -0.7]],
    "far_outside": [[0.5, 0, 7.0], [0, 0.5, -9.0]],
    "just_outside_top": [[0.5, 0, 0.0], [0, 0.5, -1.6]],
    "bigger_than_frame": [[1.7, 0, 0.1], [0, 2.5, -0.2]],
    "huge_scale": [[40.0, 0, 0.5], [0, 55.0, 0.1]],
    "anisotropic": [[0.9, 0, 0.0], [0, 0.15, 0.1]],
    "upsample_2x": [[0.2, 0, 0.3], [0, 0.25, -0.2]],
    "upsample_8x": [[0.05, 0, 0.3], [0, 0.06, -0.2]],
    "edge_exact": [[1.0, 0, 2.0 / 23.0], [0, 1.0, 0]],
    "flip_x": [[-0.8, 0, 0.1], [0, 0.7, 0]],                   # not taken by the band path
    "zero_x_only": [[0, 0, 0.2], [0, 0.8, 0]],                 # not taken
    "rot_masked": [[0.7, 0.4, 0], [-0.3, 0.6, 0.1]],           # rotation terms masked away by mask01 == 0
}


@pytest.mark.parametrize("name", sorted(BAND_THETAS))
@pytest.mark.parametrize("shape", [(24, 24, 9, 9), (17, 32, 12, 7), (8, 8, 16, 16), (20, 12, 1, 5), (13, 8, 6, 1),
                                   (64, 48, 5, 33), (40, 40, 37, 3)])
def test_band_backward_hard_transforms(emu, name, shape):
    h, w, oh, ow = shape
    rng = np.random.default_rng(abs(hash((name, shape))) % (2 ** 31))
    x = rng.random((1, 3, h, w), dtype=np.float32)
    check_band(emu, x, _theta([BAND_THETAS[name]]), (oh, ow), 0.0, seed=3)


def test_band_backward_random_boxes(emu):
    rng = np.random.default_rng(7)
    taken = 0
    for it in range(120):
        h, w = int(rng.integers(2, 48)), int(rng.integers(1, 12)) * 4
        oh, ow = int(rng.integers(1, 28)), int(rng.integers(1, 28))
        b = int(rng.integers(1, 3))
        c = int(rng.integers(1, 5))
        x = rng.random((b, c, h, w), dtype=np.float32)
        theta = np.zeros((b, 2, 3), np.float32)
        theta[:, 0, 0] = np.exp(rng.uniform(np.log(0.03), np.log(4.0), b))
        theta[:, 1, 1] = np.exp(rng.uniform(np.log(0.03), np.log(4.0), b))
        theta[:, :, 2] = rng.uniform(-1.5, 1.5, (b, 2))
        theta[:, 0, 1] = rng.uniform(-1, 1, b)         # masked away
        theta[:, 1, 0] = rng.uniform(-1, 1, b)
        cs, cap, mr = int(rng.integers(1, 9)), int(rng.integers(2, 24)), int(rng.integers(1, 12))
        ok = check_band(emu, x, theta, (oh, ow), 0.0, seed=it, layouts=((cs, cap, mr),))
        taken += 0 if ok is None else int(ok.sum())
    assert taken > 60


def test_band_backward_declines_rotations(emu):
    x = np.random.default_rng(0).random((1, 3, 16, 16), dtype=np.float32)
    check_band(emu, x, _theta([[[0.7, 0.1, 0], [0, 0.7, 0]]]), (8, 8), 1.0, expect_ok=[False], layouts=((2, 8),))
    check_band(emu, x, _theta([[[0.7, 0.1, 0], [0, 0.7, 0]]]), (8, 8), 0.0, expect_ok=[True], layouts=((2, 8),))
    check_band(emu, x, _theta([[[np.nan, 0, 0], [0, 0.7, 0]]]), (8, 8), 0.0, expect_ok=[False], layouts=((2, 8),))


# ------------------------------------------------------------------------------------------------ several crops per frame
def run_kframe(lib, theta, osz, gy, x_shape, k):
    b, c, h, w = x_shape
    n = theta.shape[0]
    oh, ow = osz
    lib.emu_kframe_gx.argtypes = [_f, ctypes.c_float, _f, _f, ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 7 + [ctypes.POINTER(ctypes.c_longlong)]
    lib.emu_kframe_gx.restype = ctypes.c_longlong
    gx = np.full((b, c, h, w), np.nan, np.float32)
    ok = np.zeros(b, np.int32)
    stats = (ctypes.c_longlong * 2)()
    conflicts = lib.emu_kframe_gx(_p(theta), 0.0, _p(gy), _p(gx), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n, k, c, h, w, oh, ow, stats)
    return gx, ok.astype(bool), conflicts, stats[0], stats[1]


@pytest.mark.parametrize("shape,k", [((3, 40, 48, 9, 9), 4), ((3, 96, 128, 25, 31), 3), ((1, 33, 20, 5, 7), 2), ((3, 224, 224, 75, 75), 2)])
def test_kframe_plan_near_the_step_threshold(emu, shape, k):
    """The row-owner gx kernel (stn_kframe.cu) rests on one assumption: a crop it takes (make_kf_crop: steps of >= ~2 frame pixels)
    puts at most one crop row on a frame row and at most one crop column of a row on a frame column.  Random boxes whose steps
    straddle the threshold (1.7 ... 2.6 frame pixels per crop pixel), hanging out of the frame on every side: no inverse-map slot
    written twice, no buffer address written twice within a crop row, gx of the taken frames equal to the oracle's."""
    c, h, w, oh, ow = shape
    rng = np.random.default_rng(sum(shape) * 7 + k)
    frames = 48
    n = frames * k
    step = rng.uniform(1.7, 2.6, (n, 2))
    near = rng.uniform(2.1, 2.4, (n, 2))                                # just above (or on) the threshold: the delicate side
    every_other = (np.arange(n) // k) % 2 == 0
    step[every_other] = near[every_other]
    theta = np.zeros((n, 2, 3), np.float32)
    theta[:, 0, 0] = step[:, 0] * max(ow - 1, 1) / (w - 1)
    theta[:, 1, 1] = step[:, 1] * max(oh - 1, 1) / (h - 1)
    theta[:, :, 2] = rng.uniform(-0.9, 0.9, (n, 2))
    theta[::(11 * k)] *= np.float32(1.7)                                # some clearly down-sampling ones
    x = rng.random((frames, c, h, w), dtype=np.float32)
    gy = rng.standard_normal((n, c, oh, ow), dtype=np.float32)
    gx, ok, conflicts, dup_rows, taken = run_kframe(emu, theta, (oh, ow), gy, x.shape, k)
    assert conflicts == 0 and dup_rows == 0
    assert 0 < taken < frames or k == 1                                 # both verdicts occur
    _, gx0, _ = oc.crop_backward(x, theta, (oh, ow), gy, None, 0.0, k)
    assert ok.sum() == taken
    sc = np.abs(gx0[ok]).max()
    assert np.abs(gx[ok] - gx0[ok]).max() <= 2e-6 * sc


def test_kframe_plan_on_the_assessor_feed(emu):
    """BASELINE config 4's box distribution (16 jittered boxes per frame) at a reduced frame size: every frame is taken."""
    wl = W.WORKLOADS["cfg4"]._replace(height=160, width=160, out_h=25, out_w=25)
    d = W.make_inputs(wl, batch=6, rotate=False)
    gx, ok, conflicts, dup_rows, taken = run_kframe(emu, d["theta"], (25, 25), d["gy"], d["x"].shape, wl.crops_per_frame)
    assert conflicts == 0 and dup_rows == 0 and taken == 6
    _, gx0, _ = oc.crop_backward(d["x"], d["theta"], (25, 25), d["gy"], None, 0.0, wl.crops_per_frame)
    assert np.abs(gx - gx0).max() <= 2e-6 * np.abs(gx0).max()
