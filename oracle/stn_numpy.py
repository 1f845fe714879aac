"""CPU oracle for the LoANs STN crop path -- numpy restatement.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product (``loans_b200``) never does; it fails loudly when its CUDA library is missing.

What it restates, and where that lives in the reference (paths relative to /root/reference):

* ``rotation_dropout_*``      -> ``functions/rotation_droput.py:26-48`` (in tree).  PINNED: checked against
  golden vectors produced by importing that very file (``tests/golden/make_golden.py``).
* ``grid_forward/backward``   -> ``chainer.functions.spatial_transformer_grid`` as called at
  ``sheep/sheep_localizer.py:62,170``.
* ``sampler_forward/backward``-> ``chainer.functions.spatial_transformer_sampler`` as called at
  ``sheep/sheep_localizer.py:63,171``.

The last two live in the third-party dependency ``chainer==4.1.0`` (``requirements.txt:1``), which is
NOT vendored under /root/reference and is not installable here (no network, no wheel).  Their CPU
algorithm (``SpatialTransformerGrid._forward/_backward`` and ``SpatialTransformerSampler._forward/
_backward`` in ``chainer/functions/array/spatial_transformer_{grid,sampler}.py``) is restated below
statement by statement from the published source, keeping numpy's own dtype promotion (int32 - float32
-> float64 for the bilinear weights) and evaluation order.  **PARITY UNPINNED** for these two: the
reference ships no test, golden vector or fixture for them.  They are instead
cross-checked against an independent implementation (``torch.nn.functional.affine_grid`` /
``grid_sample(align_corners=True, padding_mode='zeros')``), against analytic known answers and against
float64 finite differences in ``tests/test_oracle.py``.
"""
import numpy as np

__all__ = [
    "rotation_dropout_mask_value", "rotation_dropout_forward", "rotation_dropout_backward",
    "grid_coords", "grid_forward", "grid_backward", "sampler_forward", "sampler_backward",
    "crop_forward", "crop_backward", "prepare_images", "RESNET_MEAN_BGR", "grid_corners", "corners_to_dense_ggrid", "grayscale_forward", "grayscale_backward",
]


# --------------------------------------------------------------------------- a1: rotation dropout
def rotation_dropout_mask_value(ratio, train, rng=None):
    """Scalar written into mask[:,0,1] and mask[:,1,0] (functions/rotation_droput.py:30-43).

    test mode (:30-36): the ratio itself.  train mode (:38-43): ONE Bernoulli draw for the whole
    batch, ``rand(1) < ratio`` -- so ``ratio`` is the probability of KEEPING the rotation terms.
    ``rng`` is anything with ``.rand(n)`` (numpy's global ``numpy.random`` in the reference).
    """
    if not train:
        return np.float32(ratio)
    rng = np.random if rng is None else rng
    return np.float32(bool(rng.rand(1)[0] < ratio))


def _rotation_mask(theta, value):
    mask = np.ones_like(theta)                       # :33 / :39
    mask[:, 0, 1] = value                            # :34 / :42
    mask[:, 1, 0] = value                            # :35 / :43
    return mask


def rotation_dropout_forward(theta, mask_value):
    """y = theta * mask (functions/rotation_droput.py:36,45).  theta (B,2,3) float."""
    theta = np.asarray(theta)
    assert theta.dtype.kind == "f" and theta.ndim == 3 and theta.shape[1:] == (2, 3)   # :16-24
    return theta * _rotation_mask(theta, mask_value)


def rotation_dropout_backward(gy, mask_value):
    """gtheta = gy * mask (functions/rotation_droput.py:47-48)."""
    gy = np.asarray(gy)
    return gy * _rotation_mask(gy, mask_value)


# --------------------------------------------------------------------------- a2: affine grid
def grid_coords(out_h, out_w):
    """(3, oH*oW) float32 homogeneous target coordinates; row 0 = x (width), row 1 = y, row 2 = 1."""
    ys, xs = np.meshgrid(
        np.linspace(-1, 1, out_h, dtype=np.float32),
        np.linspace(-1, 1, out_w, dtype=np.float32), indexing="ij", copy=False)
    coords = np.concatenate(
        [xs[None], ys[None], np.ones((1, out_h, out_w), dtype=np.float32)], axis=0)
    return coords.reshape(3, out_h * out_w)


def grid_forward(theta, output_shape):
    """theta (B,2,3) f32 -> grid (B,2,oH,oW) f32;  grid = theta . [xs; ys; 1]."""
    theta = np.asarray(theta)
    assert theta.dtype == np.float32 and theta.ndim == 3 and theta.shape[1:] == (2, 3)
    out_h, out_w = output_shape
    b = theta.shape[0]
    coords = grid_coords(out_h, out_w)
    return theta.dot(coords).reshape(b, 2, out_h, out_w)


def grid_backward(ggrid):
    """ggrid (B,2,oH,oW) -> gtheta (B,2,3);  gtheta = ggrid . coords^T."""
    ggrid = np.asarray(ggrid)
    b, _, out_h, out_w = ggrid.shape
    coords = grid_coords(out_h, out_w)
    return ggrid.reshape(b, 2, out_h * out_w).dot(coords.T).astype(ggrid.dtype, copy=False)


# --------------------------------------------------------------------------- a3/a4: bilinear sampler
def _sampler_common(x, grid):
    b, c, h, w = x.shape
    grid = grid.reshape(grid.shape[:2] + (-1,))
    u = grid[:, 0]
    v = grid[:, 1]
    # zero frame of one pixel so that out-of-image taps read 0
    x_pad = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="constant")
    # [-1,1] -> [0, size-1], then +1 for the padding.  Four separately rounded float32 operations.
    u = (u + 1) * (w - 1) / 2 + 1
    v = (v + 1) * (h - 1) / 2 + 1
    u_clipped = u.clip(0, w + 1)
    v_clipped = v.clip(0, h + 1)
    u0 = np.floor(u_clipped).astype(np.int32)
    u0 = u0.clip(0, w)
    u1 = u0 + 1
    v0 = np.floor(v_clipped).astype(np.int32)
    v0 = v0.clip(0, h)
    v1 = v0 + 1
    return x_pad, u, v, u_clipped, v_clipped, u0, u1, v0, v1


def _taps(x_pad, u0, u1, v0, v1):
    b = x_pad.shape[0]
    g = lambda vv, uu: np.concatenate(                                   # noqa: E731
        [np.expand_dims(x_pad[i, :, vv[i], uu[i]], axis=0) for i in range(b)], axis=0)
    return g(v0, u0), g(v0, u1), g(v1, u0), g(v1, u1)                  # each (B, N, C)


def sampler_forward(x, grid):
    """x (B,C,H,W) f32, grid (B,2,oH,oW) f32 -> y (B,C,oH,oW) f32.  Zero padding, align-corners."""
    x = np.asarray(x)
    grid = np.asarray(grid)
    assert x.dtype == np.float32 and grid.dtype == np.float32
    assert x.ndim == 4 and grid.ndim == 4 and grid.shape[1] == 2 and x.shape[0] == grid.shape[0]
    b, c, h, w = x.shape
    out_h, out_w = grid.shape[2:]
    x_pad, _, _, uc, vc, u0, u1, v0, v1 = _sampler_common(x, grid)
    # int32 - float32 promotes to float64 in numpy: exact differences, exact product, one rounding
    w1 = ((u1 - uc) * (v1 - vc)).astype(x_pad.dtype)
    w2 = ((uc - u0) * (v1 - vc)).astype(x_pad.dtype)
    w3 = ((u1 - uc) * (vc - v0)).astype(x_pad.dtype)
    w4 = ((uc - u0) * (vc - v0)).astype(x_pad.dtype)
    x1, x2, x3, x4 = _taps(x_pad, u0, u1, v0, v1)
    y = w1[:, :, None] * x1
    y += w2[:, :, None] * x2
    y += w3[:, :, None] * x3
    y += w4[:, :, None] * x4
    return np.ascontiguousarray(y.reshape(b, out_h, out_w, c).transpose(0, 3, 1, 2))


def sampler_backward(x, grid, gy):
    """-> (gx (B,C,H,W), ggrid (B,2,oH,oW)).  gx via unbuffered scatter-add (numpy.add.at)."""
    x = np.asarray(x)
    grid = np.asarray(grid)
    gy = np.asarray(gy)
    b, c, h, w = x.shape
    out_h, out_w = grid.shape[2:]
    x_pad, u, v, uc, vc, u0, u1, v0, v1 = _sampler_common(x, grid)
    wu0 = (uc - u0).astype(gy.dtype)
    wu1 = (u1 - uc).astype(gy.dtype)
    wv0 = (vc - v0).astype(gy.dtype)
    wv1 = (v1 - vc).astype(gy.dtype)

    x1, x2, x3, x4 = _taps(x_pad, u0, u1, v0, v1)
    gu = -wv1[:, :, None] * x1
    gu += wv1[:, :, None] * x2
    gu -= wv0[:, :, None] * x3
    gu += wv0[:, :, None] * x4
    gv = -wu1[:, :, None] * x1
    gv -= wu0[:, :, None] * x2
    gv += wu1[:, :, None] * x3
    gv += wu0[:, :, None] * x4
    gu = gu.reshape(b, out_h, out_w, c).transpose(0, 3, 1, 2)
    gv = gv.reshape(b, out_h, out_w, c).transpose(0, 3, 1, 2)
    gu = gu * gy
    gv = gv * gy
    gu = np.sum(gu, axis=1)
    gv = np.sum(gv, axis=1)
    # chain rule of the rescaling; gradient is cut where the UNclipped coordinate left the padded image
    u_r = u.reshape(gu.shape)
    v_r = v.reshape(gv.shape)
    gu = gu / 2. * (w - 1) * (u_r > 0) * (u_r < (w + 1))
    gv = gv / 2. * (h - 1) * (v_r > 0) * (v_r < (h + 1))
    ggrid = np.concatenate((gu[:, None], gv[:, None]), axis=1).astype(gy.dtype, copy=False)

    gx = np.zeros_like(x_pad)
    gyr = gy.reshape(b, c, -1)
    for i in range(b):
        np.add.at(gx[i], (slice(None), v0[i], u0[i]), gyr[i] * wu1[i] * wv1[i])
        np.add.at(gx[i], (slice(None), v0[i], u1[i]), gyr[i] * wu0[i] * wv1[i])
        np.add.at(gx[i], (slice(None), v1[i], u0[i]), gyr[i] * wu1[i] * wv0[i])
        np.add.at(gx[i], (slice(None), v1[i], u1[i]), gyr[i] * wu0[i] * wv0[i])
    gx = np.ascontiguousarray(gx[:, :, 1:-1, 1:-1])
    return gx, ggrid


# --------------------------------------------------------------------------- a5: the composite
def crop_forward(x, theta, output_shape, mask_value=1.0, crops_per_frame=1):
    """rotation_dropout -> grid -> sampler as wired at sheep/sheep_localizer.py:61-63.

    ``crops_per_frame`` K > 1 (BASELINE config 4) has no Chainer equivalent -- the reference sampler
    needs equal batch sizes -- and is defined as sampling ``repeat(x, K, axis=0)``.
    Returns (y, grid).
    """
    th = rotation_dropout_forward(np.asarray(theta, dtype=np.float32), np.float32(mask_value))
    grid = grid_forward(th, output_shape)
    xx = np.repeat(x, crops_per_frame, axis=0) if crops_per_frame > 1 else x
    return sampler_forward(xx, grid), grid


def crop_backward(x, theta, output_shape, gy, ggrid_upstream=None, mask_value=1.0, crops_per_frame=1):
    """Backward of ``crop_forward``: returns (gtheta (N,2,3), gx (B,C,H,W), ggrid_sampler (N,2,oH,oW))."""
    k = crops_per_frame
    th = rotation_dropout_forward(np.asarray(theta, dtype=np.float32), np.float32(mask_value))
    grid = grid_forward(th, output_shape)
    xx = np.repeat(x, k, axis=0) if k > 1 else x
    gx, ggrid = sampler_backward(xx, grid, gy)
    if k > 1:
        gx = gx.reshape((x.shape[0], k) + x.shape[1:]).sum(axis=1, dtype=np.float32)
    total = ggrid if ggrid_upstream is None else ggrid + np.asarray(ggrid_upstream, dtype=np.float32)
    gtheta = rotation_dropout_backward(grid_backward(total), np.float32(mask_value))
    return gtheta, gx, ggrid


# --------------------------------------------------------------------------- f1: prepare_images
RESNET_MEAN_BGR = np.array([103.063, 115.903, 123.152], dtype=np.float32)


def prepare_images(scaled_images):
    """SheepLocalizer.prepare_images (sheep/sheep_localizer.py:72-82) applied to what the caller hands it,
    ``images.copy() * 255`` (:45): per image ``chainer.links.model.vision.resnet.prepare(image, size=None)``, stacked.

    resnet.prepare lives in chainer 4.1.0 (third party, not in /root/reference; PARITY UNPINNED for the mean constants,
    restated from the published source): ndarray (3,H,W) -> transpose to (H,W,3) -> ``astype(numpy.uint8)`` ->
    ``Image.fromarray`` -> ``convert('RGB')`` (a no-op for an RGB array) -> ``numpy.asarray(image, float32)`` ->
    ``image[:, :, ::-1]`` (RGB -> BGR) -> ``image -= [103.063, 115.903, 123.152]`` -> transpose to (3,H,W).
    tests/test_oracle.py checks this restatement against the literal PIL chain.
    """
    scaled_images = np.asarray(scaled_images)
    assert scaled_images.ndim == 4 and scaled_images.shape[1] == 3
    out = []
    for image in scaled_images:                                   # F.separate(images, axis=0)  :77
        image = image.transpose((1, 2, 0)).astype(np.uint8)       # truncation toward zero
        image = np.asarray(image, dtype=np.float32)
        image = image[:, :, ::-1]
        image = image - RESNET_MEAN_BGR
        out.append(image.transpose((2, 0, 1)))
    return np.stack(out, axis=0)                                  # F.stack  :78


# --------------------------------------------------------------------------- f2: the grid's four corner points
def grid_corners(grid):
    """grid (N,2,oH,oW) -> (N,2,2,2): the points LoANs reads besides sampling -- [0,0], [0,W-1], [H-1,0] in
    LossCalculator.get_corners (common/utils.py:141-159) and [0,0], [-1,-1] in extract_corners
    (sheep/sheep_localizer.py:84-91)."""
    return np.ascontiguousarray(grid[:, :, [0, -1]][:, :, :, [0, -1]])


def corners_to_dense_ggrid(gcorners, out_h, out_w):
    """The gradient arriving on those four points as the dense upstream grid gradient the reference would see
    (zeros elsewhere; coinciding corners of a 1-row / 1-column crop add up, as get_item's backward does)."""
    n = gcorners.shape[0]
    gg = np.zeros((n, 2, out_h, out_w), np.float32)
    for ci, i in enumerate((0, out_h - 1)):
        for cj, j in enumerate((0, out_w - 1)):
            gg[:, :, i, j] += gcorners[:, :, ci, cj]
    return gg


# --------------------------------------------------------------------------- f3: the localizer's grayscale epilogue
def grayscale_forward(rois):
    """sheep/sheep_localizer.py:65-68: ``b, g, r = F.split_axis(rois, 3, axis=1); rois = 0.299 * r + 0.587 * g + 0.114 * b``
    (chainer: the Python constants are cast to the array dtype, float32 products, summed left to right)."""
    rois = np.asarray(rois)
    assert rois.shape[1] == 3, "rois are not in RGB, can not convert them to grayscale"      # :66
    b, g, r = rois[:, 0:1], rois[:, 1:2], rois[:, 2:3]
    return np.float32(0.299) * r + np.float32(0.587) * g + np.float32(0.114) * b


def grayscale_backward(ggray):
    """gradient w.r.t. the 3-channel rois: MulConstant's backward is ``value * gy``, Add's passes gy through."""
    ggray = np.asarray(ggray)
    return np.concatenate([np.float32(0.114) * ggray, np.float32(0.587) * ggray, np.float32(0.299) * ggray], axis=1)
