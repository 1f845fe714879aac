"""Regenerates the committed golden fixtures.  Run in the build container:  python tests/golden/make_golden.py

1. ``rotation_dropout.npz`` -- produced by IMPORTING THE REFERENCE'S OWN FILE
   ``/root/reference/functions/rotation_droput.py`` and calling its ``rotation_dropout`` /
   ``RotationDropout.backward``.  That file only needs five names from Chainer
   (``function.Function``, ``configuration.config``, ``cuda.get_array_module``, ``type_check.expect``,
   and the ``Function.__call__ -> forward`` protocol); Chainer itself is not installable here, so a
   throw-away stand-in package providing exactly those names is put on ``sys.path`` for the import.  The
   arithmetic that lands in the fixture is the reference's, executed unmodified.  This pins oracle
   row a1.

2. ``stn_small.npz`` -- small seeded cases of the grid / sampler / composite, produced by the numpy
   restatement ``oracle/stn_numpy.py`` and cross-checked here against torch's independent
   ``affine_grid`` / ``grid_sample(align_corners=True, padding_mode='zeros')`` before being written.
   Chainer 4.1.0 (where this arithmetic lives) is absent, so these vectors are ORACLE-DERIVED: they
   guard the oracle and the CUDA path against drift, they do not pin them to Chainer.

3. ``prepare_images.npz`` -- SheepLocalizer.prepare_images (reference sheep/sheep_localizer.py:72-82): the literal
   statement sequence of chainer 4.1.0's ``resnet.prepare(image, size=None)`` (uint8 cast, PIL ``Image.fromarray`` /
   ``convert('RGB')``, float32, BGR flip, mean subtraction) executed with the real PIL on seeded frames, including
   values on the quantisation boundaries.  Chainer is absent, so the statement sequence and the mean constants are
   restated from the published source (parity unpinned for them); PIL's part is executed, not restated.

4. ``ingest_lanczos.npz`` -- the loader's frame path (reference common/datasets/image_dataset.py:16-28, :98): the reference's
   own statement sequence (``Image.fromarray(...).convert('RGB').resize(..., Image.LANCZOS)``, float32 CHW, ``/ 255``) EXECUTED
   with the real PIL on seeded uint8 frames -- down-sampling, up-sampling, one axis only, no resampling.  Pins
   oracle/ingest_numpy.py and the CUDA kernels (bit for bit).

/root/reference is read at generation time only; nothing under tests/ reads it at test time.
"""
import importlib.util
import os
import sys
import tempfile
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE_FILE = "/root/reference/functions/rotation_droput.py"


def _import_reference_rotation_dropout():
    stub = tempfile.mkdtemp(prefix="chainer_standin_")
    pkg = os.path.join(stub, "chainer")
    os.makedirs(os.path.join(pkg, "utils"))
    files = {
        "__init__.py": "from chainer import configuration, cuda, function, utils\n",
        "configuration.py": "class _Cfg:\n    train = True\nconfig = _Cfg()\n",
        "cuda.py": "import numpy\ndef get_array_module(*args):\n    return numpy\n",
        "function.py": textwrap.dedent("""
            class Function(object):
                def retain_inputs(self, indexes):
                    self._retained = indexes
                def __call__(self, *inputs):
                    outs = self.forward(tuple(inputs))
                    return outs[0]
            """),
        "utils/__init__.py": "from chainer.utils import type_check\n",
        "utils/type_check.py": "def expect(*conds):\n    pass\n",
    }
    for name, body in files.items():
        with open(os.path.join(pkg, name), "w") as f:
            f.write(body)
    sys.path.insert(0, stub)
    spec = importlib.util.spec_from_file_location("reference_rotation_droput", REFERENCE_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import chainer
    return mod, chainer.configuration.config


def make_rotation_dropout():
    mod, config = _import_reference_rotation_dropout()
    rng = np.random.default_rng(20181017)
    cases = {}
    idx = 0
    for train in (True, False):
        for ratio in (0.0, 0.25, 0.5, 0.75, 1.0):
            for seed in (0, 1, 2, 3):
                b = int(rng.integers(1, 9))
                theta = rng.standard_normal((b, 2, 3)).astype(np.float32)
                gy = rng.standard_normal((b, 2, 3)).astype(np.float32)
                config.train = train
                np.random.seed(seed)                     # the reference draws from numpy's global stream
                node = mod.RotationDropout(ratio)
                y = node(theta)
                if train:
                    gtheta = node.backward((theta,), (gy,))[0]
                else:
                    # reference :47-48 reads self.mask, which a test-mode forward never sets
                    try:
                        node.backward((theta,), (gy,))
                        raise SystemExit("reference unexpectedly supports backward after test-mode forward")
                    except AttributeError:
                        gtheta = np.zeros_like(gy)
                # the functional front end must agree with the node
                np.random.seed(seed)
                y2 = mod.rotation_dropout(theta, ratio=ratio)
                assert np.array_equal(y, y2)
                p = "c%03d_" % idx
                cases[p + "theta"] = theta
                cases[p + "gy"] = gy
                cases[p + "y"] = np.asarray(y, np.float32)
                cases[p + "gtheta"] = np.asarray(gtheta, np.float32)
                cases[p + "meta"] = np.array([float(train), ratio, float(seed)], np.float64)
                idx += 1
    cases["n_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "rotation_dropout.npz"), **cases)
    print("rotation_dropout.npz:", idx, "cases from", REFERENCE_FILE)


def _torch_check(x, theta, osz, y, grid, gy, gx, gtheta):
    import torch
    import torch.nn.functional as F
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    tt = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    g = F.affine_grid(tt, (x.shape[0], x.shape[1]) + tuple(osz), align_corners=True)
    yt = F.grid_sample(xt, g, mode="bilinear", padding_mode="zeros", align_corners=True)
    yt.backward(torch.tensor(gy, dtype=torch.float64))
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))          # noqa: E731
    errs = (rel(grid, g.detach().numpy().transpose(0, 3, 1, 2)), rel(y, yt.detach().numpy()),
            rel(gx, xt.grad.numpy()), rel(gtheta, tt.grad.numpy()))
    return errs


def make_stn_small():
    from oracle import stn_numpy as on
    from loans_b200 import workloads as W
    rng = np.random.default_rng(424242)
    shapes = [  # (B, C, H, W, oH, oW, mask_value, smooth)
        (2, 3, 16, 20, 7, 9, 1.0, True),
        (3, 1, 9, 9, 9, 9, 0.0, True),
        (2, 3, 33, 18, 12, 5, 0.5, False),
        (1, 2, 5, 7, 16, 16, 1.0, False),     # up-sampling
        (4, 3, 24, 24, 8, 8, 0.0, False),
    ]
    out = {"n_cases": np.array(len(shapes))}
    for idx, (b, c, h, w, oh, ow, mask, smooth) in enumerate(shapes):
        x = W.make_frames(rng, b, c, h, w, smooth=smooth)
        theta = W.make_theta(rng, b, rotate=True)
        if idx == 3:
            theta[:, :, :2] *= 0.3
        gy = rng.standard_normal((b, c, oh, ow), dtype=np.float32)
        gg_up = rng.standard_normal((b, 2, oh, ow), dtype=np.float32)
        y, grid = on.crop_forward(x, theta, (oh, ow), mask)
        gtheta, gx, ggrid = on.crop_backward(x, theta, (oh, ow), gy, gg_up, mask)
        # independent cross-check (no upstream grid gradient, masked theta fed directly)
        th_m = on.rotation_dropout_forward(theta, np.float32(mask))
        gt0, gx0, _ = on.crop_backward(x, th_m, (oh, ow), gy, None, 1.0)
        errs = _torch_check(x, th_m, (oh, ow), y, grid, gy, gx0, gt0)
        print("stn_small case %d: rel err vs torch fp64  grid %.1e  y %.1e  gx %.1e  gtheta %.1e" % ((idx,) + errs))
        # grid/gx are smooth in theta -> tight; y and gtheta on noise frames move with 1-ulp coordinate shifts
        assert errs[0] < 1e-6 and errs[1] < 5e-5 and errs[2] < 5e-5 and errs[3] < 5e-4, errs
        p = "c%d_" % idx
        out.update({p + "x": x, p + "theta": theta, p + "gy": gy, p + "ggrid_up": gg_up,
                    p + "mask": np.array(mask, np.float32), p + "out_size": np.array([oh, ow]),
                    p + "y": y, p + "grid": grid, p + "gtheta": gtheta, p + "gx": gx, p + "ggrid": ggrid})
    np.savez_compressed(os.path.join(HERE, "stn_small.npz"), **out)
    print("stn_small.npz:", len(shapes), "cases (oracle-derived, torch cross-checked)")


def make_prepare_images():
    from PIL import Image
    rng = np.random.default_rng(4242)
    out = {}
    shapes = [(2, 3, 5, 7), (1, 3, 8, 8), (3, 3, 9, 12), (2, 3, 1, 3)]
    for i, shp in enumerate(shapes):
        x = rng.random(shp, dtype=np.float32)
        if i == 1:                                    # exact quantisation boundaries k / 255 and their float32 neighbours
            k = rng.integers(0, 256, shp).astype(np.float32)
            x = (k / np.float32(255)).astype(np.float32)
            x[..., ::2] = np.nextafter(x[..., ::2], np.float32(0))
            x[0, 0, 0, :4] = [0.0, 1.0, 0.5, 1.0 / 255]
        scaled = x * 255                              # images.copy() * 255, sheep/sheep_localizer.py:45
        res = []
        for image in scaled:
            im = image.transpose((1, 2, 0))
            im = Image.fromarray(im.astype(np.uint8))
            im = im.convert('RGB')
            im = np.asarray(im, dtype=np.float32)
            im = im[:, :, ::-1]
            im = im - np.array([103.063, 115.903, 123.152], dtype=np.float32)
            res.append(im.transpose((2, 0, 1)))
        out["c%d_x" % i] = x
        out["c%d_out" % i] = np.stack(res, axis=0)
    out["n_cases"] = np.int64(len(shapes))
    np.savez_compressed(os.path.join(HERE, "prepare_images.npz"), **out)
    print("prepare_images.npz:", len(shapes), "cases (literal resnet.prepare statement sequence through PIL)")


def make_ingest():
    from oracle import ingest_numpy as ig
    rng = np.random.default_rng(20181018)
    cases = {}
    shapes = [(37, 53, 20, 31), (16, 16, 40, 24), (24, 20, 24, 20), (90, 120, 75, 75), (24, 32, 56, 56), (51, 77, 51, 30),
              (9, 200, 3, 7), (300, 5, 10, 5), (64, 48, 32, 48)]
    for i, (h, w, oh, ow) in enumerate(shapes):
        f = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        f[0, ::3, ::5] = 255                                    # saturated / black pixels next to noise: clipping of the ringing
        f[0, 1::4, 2::7] = 0
        out = ig.pil_reference(f, (oh, ow))                      # the real PIL
        assert np.array_equal(out, ig.ingest(f, (oh, ow))), "oracle restatement differs from PIL at %s" % ((h, w, oh, ow),)
        out_u8 = np.rint(out * 255).astype(np.uint8)            # stored as the uint8 image PIL returned (float32 / 255 is
        assert np.array_equal(out_u8.astype(np.float32) / 255, out)   # re-applied by the tests: the reference's last statement)
        cases["c%02d_frames" % i] = f
        cases["c%02d_out_u8" % i] = out_u8
        cases["c%02d_size" % i] = np.array([oh, ow])
    cases["n_cases"] = np.array(len(shapes))
    import PIL
    cases["pil_version"] = np.array(PIL.__version__)
    np.savez_compressed(os.path.join(HERE, "ingest_lanczos.npz"), **cases)
    print("ingest_lanczos.npz: %d cases from PIL %s" % (len(shapes), PIL.__version__))


if __name__ == "__main__":
    make_rotation_dropout()
    make_stn_small()
    make_prepare_images()
    make_ingest()
