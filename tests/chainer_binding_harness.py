"""Runs loans_b200/chainer_compat.py -- install() and the reference's three operator calls written as the reference writes
them (sheep/sheep_localizer.py:60-63) -- on cuda:0 against the chainer / cupy STAND-IN of tests/chainer_standin (the real
packages are not installable here), and prints one JSON line of results.  Own process: the stand-in never enters the
test session's sys.modules.  Test infrastructure; started by tests/test_gpu_chainer_binding.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "chainer_standin")]

import numpy as np  # noqa: E402
import chainer  # noqa: E402
import cupy  # noqa: E402

assert getattr(chainer, "__standin__", False) and getattr(cupy, "__standin__", False)
import loans_b200.chainer_compat as stn  # noqa: E402
from loans_b200 import _lib  # noqa: E402
from loans_b200 import workloads as W  # noqa: E402
from oracle import stn_c as oc  # noqa: E402

assert stn.HAVE_CHAINER


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / max(float(np.abs(b).max()), 1e-30))


def main():
    wl = W.WORKLOADS["cfg1"]
    d = W.make_inputs(wl, batch=8, rotate=True, with_ggrid=True)
    osz = (wl.out_h, wl.out_w)
    y0, grid0 = oc.crop_forward(d["x"], d["theta"], osz, 0.0)
    gt0, gx0, _ = oc.crop_backward(d["x"], d["theta"], osz, d["gy"], d["ggrid"], 0.0)
    out = {}
    for fuse in ("full", "sampler", "off"):
        for frames_need_grad in (False, True):
            stn.install(fuse=fuse)
            import chainer.functions as F
            from functions.rotation_droput import rotation_dropout      # (sic) the reference's import, sheep_localizer.py:12

            class Localizer(object):
                out_size = osz
            self = Localizer()
            images = cupy.asarray(d["x"])                               # a raw array, as every LoANs caller passes it
            if frames_need_grad:
                images = chainer.Variable(images)
            h = chainer.Variable(cupy.asarray(d["theta"].reshape(-1, 6)))   # what Linear(512, 6) hands over
            n0 = _lib.launch_count()
            # ---- reference sheep/sheep_localizer.py:61-63, verbatim
            transform_params = rotation_dropout(F.reshape(h, (-1, 2, 3)), ratio=0.0)
            points = F.spatial_transformer_grid(transform_params, self.out_size)
            rois = F.spatial_transformer_sampler(images, points)
            # ----
            n1 = _lib.launch_count()
            rois.grad = cupy.asarray(d["gy"])
            points.grad = cupy.asarray(d["ggrid"])                      # what the corner regularisers send back
            chainer.backward_all([rois, points])
            n2 = _lib.launch_count()
            r = {"fwd_launches": n1 - n0, "bwd_launches": n2 - n1,
                 "rois_exact": bool(np.array_equal(rois.data.get(), y0)), "points_exact": bool(np.array_equal(points.data.get(), grid0)),
                 "gtheta_rel": rel(h.grad.get().reshape(-1, 2, 3), gt0)}
            if frames_need_grad:
                r["gx_rel"] = rel(images.grad.get(), gx0)
            # without a gradient on points: the sampler's backward alone
            h2 = chainer.Variable(cupy.asarray(d["theta"].reshape(-1, 6)))
            pts = F.spatial_transformer_grid(rotation_dropout(F.reshape(h2, (-1, 2, 3)), ratio=0.0), osz)
            ro = F.spatial_transformer_sampler(images, pts)
            ro.grad = cupy.asarray(d["gy"])
            n3 = _lib.launch_count()
            ro.backward()
            r["bwd_launches_no_points_grad"] = _lib.launch_count() - n3
            r["bwd_kernel"] = _lib.last_kernel()
            out["%s%s" % (fuse, "_gx" if frames_need_grad else "")] = r
    # test mode: ratio scales the rotation terms, backward raises like the reference (functions/rotation_droput.py:30-36,47-48)
    stn.install(fuse="full")
    import chainer.functions as F
    from functions.rotation_droput import rotation_dropout
    with chainer.using_config("train", False):
        h = chainer.Variable(cupy.asarray(d["theta"]))
        pts = F.spatial_transformer_grid(rotation_dropout(h, ratio=0.25), osz)
        ro = F.spatial_transformer_sampler(cupy.asarray(d["x"]), pts)
    y1, g1 = oc.crop_forward(d["x"], d["theta"], osz, 0.25)
    out["test_mode_exact"] = bool(np.array_equal(ro.data.get(), y1) and np.array_equal(pts.data.get(), g1))
    ro.grad = cupy.asarray(d["gy"])
    try:
        ro.backward()
        out["test_mode_backward_raises"] = False
    except AttributeError:
        out["test_mode_backward_raises"] = True
    # type checks and the no-CPU-fallback rule
    try:
        F.spatial_transformer_grid(chainer.Variable(cupy.asarray(d["theta"].astype(np.float64))), osz)
        out["float64_theta_refused"] = False
    except chainer.utils.type_check.InvalidType:
        out["float64_theta_refused"] = True
    try:
        F.spatial_transformer_sampler(d["x"], F.spatial_transformer_grid(chainer.Variable(cupy.asarray(d["theta"])), osz))
        out["numpy_frames_refused"] = False
    except RuntimeError:
        out["numpy_frames_refused"] = True
    # prepare_images patch
    class FakeLocalizer(object):
        pass
    stn.patch_localizers(FakeLocalizer)
    from oracle import stn_numpy as on
    prep = FakeLocalizer().prepare_images(chainer.Variable(cupy.asarray(d["x"] * 255)))
    out["prepare_images_exact"] = bool(np.array_equal(prep.data.get(), on.prepare_images(d["x"] * 255)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
