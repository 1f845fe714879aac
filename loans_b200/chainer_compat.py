"""Chainer 4.x binding of libloans_stn.so: the reference-side FFI a LoANs maintainer would add.

    import loans_b200.chainer_compat as stn
    stn.install()            # once, before sheep/sheep_localizer.py is imported

rebinds ``chainer.functions.spatial_transformer_grid`` / ``spatial_transformer_sampler`` to the functions below and
registers a ``functions.rotation_droput`` module exposing ``rotation_dropout`` / ``RotationDropout``, so that reference
``sheep/sheep_localizer.py``, ``sheep/sheep_updater.py`` and ``iou/iou_regressor.py`` run unchanged (they reach the three
operators only through those names: sheep_localizer.py:2,12,61-63,169-171).

How the three calls reach the fused kernels (``install(fuse=...)``):

``"full"`` (default)  ``rotation_dropout`` runs its 6-floats-per-crop kernel and notes (theta_in, mask value) on its output;
    ``spatial_transformer_grid`` allocates ``points`` WITHOUT launching and passes the note on; ``spatial_transformer_sampler``
    launches ``loans_stn_crop_fwd(theta_in, mask01=<the draw>)`` once -- it writes the crops and fills ``points`` -- and its
    backward is one ``loans_stn_crop_bwd`` with the same mask (the band / table kernels at LoANs' ratio 0.0), gradient
    straight to the un-masked theta.  A gradient arriving on ``points`` (corner regularisers) takes the grid node's own
    tiny backward.  Contract: ``points`` holds its values only after the sampler call that follows -- exactly the
    reference's call order; use another mode for code that reads the grid in between.
``"sampler"``  grid computed eagerly by the grid node; the sampler, handed that very array, still samples from theta in
    registers (the masked theta, ``LOANS_STN_FLAG_UPRIGHT`` when the dropout zeroed the rotation terms).  Assumes nobody
    edits the grid array in place between the two calls (cupy arrays carry no version counter).
``"off"``  three independent nodes (explicit-grid sampler): no assumption at all.

Chainer and cupy are NOT installed in the build container or on the GPU box (chainer==4.1.0 / cupy==4.1.0, reference
requirements.txt:1-3, are not installable offline).  The module is import-guarded; it is EXECUTED by
``tests/test_gpu_chainer_binding.py`` against a stand-in for the two packages (tests/chainer_standin: Variable, the v4
FunctionNode protocol, type_check, a cupy ndarray over torch memory), which proves the code paths and the numbers, not
compatibility with a real Chainer build -- treat it as experimental until it has run against chainer 4.x.
Everything here is pointer plumbing: ``cupy.ndarray.data.ptr`` in, ``cupy.cuda.get_current_stream().ptr`` as the stream,
outputs allocated from cupy's pool.
"""
import sys
import types

from loans_b200 import _lib

try:
    import chainer
    from chainer import cuda, function_node
    from chainer.utils import type_check
    HAVE_CHAINER = True
except ImportError:
    chainer = None
    HAVE_CHAINER = False

_STATE = {"fuse": "full"}


def _require():
    if not HAVE_CHAINER:
        raise ImportError("loans_b200.chainer_compat needs chainer (4.x) and cupy; use loans_b200.functions with torch "
                          "tensors otherwise")


def _ptr(a):
    return None if a is None else int(a.data.ptr)


def _stream():
    return int(cuda.cupy.cuda.get_current_stream().ptr)


def _gpu_only(*arrays):
    for a in arrays:
        if a is not None and not isinstance(a, cuda.ndarray):
            raise RuntimeError("loans_b200 runs on cupy arrays only (no CPU fallback): move the model with to_gpu()")


if HAVE_CHAINER:

    class RotationDropout(function_node.FunctionNode):
        """reference functions/rotation_droput.py:9-48 (old-style Function there; same semantics)."""

        def __init__(self, dropout_ratio):
            self.dropout_ratio = dropout_ratio

        def check_type_forward(self, in_types):
            type_check.expect(in_types.size() == 1)
            x_type = in_types[0]
            type_check.expect(x_type.dtype.kind == 'f', x_type.ndim == 3, x_type.shape[1] == 2, x_type.shape[2] == 3)

        def forward(self, inputs):
            x, = inputs
            _gpu_only(x)
            xp = cuda.cupy
            x = xp.ascontiguousarray(x, dtype=xp.float32)
            if not chainer.config.train:
                self.mask_value = None
                value = float(self.dropout_ratio)                      # :33-35
            else:
                if not hasattr(self, 'mask_value') or self.mask_value is None:
                    self.mask_value = float(bool(xp.random.rand(1) < self.dropout_ratio))     # :41, one draw per call
                value = self.mask_value
            self.value = value
            y = xp.empty_like(x)
            _lib.check(_lib.lib().loans_stn_rotation_dropout(_ptr(x), value, _ptr(y), x.shape[0], _stream()),
                       "loans_stn_rotation_dropout")
            return y,

        def backward(self, indexes, grad_outputs):
            if self.mask_value is None:
                raise AttributeError("'RotationDropout' object has no attribute 'mask'")    # :47-48 after a test-mode forward
            gy = cuda.cupy.ascontiguousarray(grad_outputs[0].data)
            gx = cuda.cupy.empty_like(gy)
            _lib.check(_lib.lib().loans_stn_rotation_dropout(_ptr(gy), self.mask_value, _ptr(gx), gy.shape[0], _stream()),
                       "loans_stn_rotation_dropout")
            return chainer.Variable(gx),

    def rotation_dropout(x, ratio=.5, **kwargs):
        node = RotationDropout(ratio)
        y, = node.apply((x,))
        x_var = node.inputs[0]
        x_arr = x_var.data
        if isinstance(x_arr, cuda.ndarray) and x_arr.dtype == cuda.cupy.float32:
            # note for the grid node: where this theta came from (checked again by pointer before it is trusted)
            y._stn_dropout = {"theta_in": x_var, "theta_in_ptr": _ptr(x_arr), "mask01": node.value,
                              "can_backprop": node.mask_value is not None, "array": y.data, "ptr": _ptr(y.data)}
        return y

    class SpatialTransformerGrid(function_node.FunctionNode):
        def __init__(self, output_shape, defer=False):
            self.output_shape = tuple(int(v) for v in output_shape)
            self.defer = defer

        def check_type_forward(self, in_types):
            type_check.expect(in_types.size() == 1)
            theta_type = in_types[0]
            type_check.expect(theta_type.dtype.char == 'f', theta_type.ndim == 3, theta_type.shape[1] == 2,
                              theta_type.shape[2] == 3)

        def forward(self, inputs):
            theta, = inputs
            _gpu_only(theta)
            xp = cuda.cupy
            theta = xp.ascontiguousarray(theta)
            oh, ow = self.output_shape
            grid = xp.empty((theta.shape[0], 2, oh, ow), dtype=xp.float32)
            if not self.defer:            # "full" mode: the sampler's fused kernel fills it
                _lib.check(_lib.lib().loans_stn_grid_fwd(_ptr(theta), _ptr(grid), theta.shape[0], oh, ow, _stream()),
                           "loans_stn_grid_fwd")
            return grid,

        def backward(self, indexes, grad_outputs):
            xp = cuda.cupy
            ggrid = xp.ascontiguousarray(grad_outputs[0].data)
            n, _, oh, ow = ggrid.shape
            gtheta = xp.empty((n, 2, 3), dtype=xp.float32)
            _lib.check(_lib.lib().loans_stn_grid_bwd(_ptr(ggrid), _ptr(gtheta), n, oh, ow, _stream()), "loans_stn_grid_bwd")
            return chainer.Variable(gtheta),

    class FusedSampler(function_node.FunctionNode):
        """sampler whose grid came straight from our grid node: inputs (x, theta); coordinates are recomputed from theta in
        registers (loans_stn_crop_fwd / _bwd), the gradient goes to theta directly.  ``mask01``: the rotation-dropout value
        folded into the kernel (theta is then the UN-masked theta); ``grid_out``: the grid node's deferred output, filled by
        the same launch; ``upright``: theta is already masked to axis-aligned boxes (LOANS_STN_FLAG_UPRIGHT)."""

        def __init__(self, output_shape, mask01=1.0, grid_out=None, upright=False, can_backprop=True):
            self.output_shape = tuple(int(v) for v in output_shape)
            self.mask01 = float(mask01)
            self.grid_out = grid_out
            self.upright = bool(upright)
            self.can_backprop = bool(can_backprop)

        def check_type_forward(self, in_types):
            type_check.expect(in_types.size() == 2)
            x_type, theta_type = in_types
            type_check.expect(x_type.dtype.char == 'f', theta_type.dtype.char == 'f', x_type.ndim == 4, theta_type.ndim == 3,
                              theta_type.shape[1] == 2, theta_type.shape[2] == 3, x_type.shape[0] == theta_type.shape[0])

        def forward(self, inputs):
            x, theta = inputs
            _gpu_only(x, theta)
            xp = cuda.cupy
            x, theta = xp.ascontiguousarray(x), xp.ascontiguousarray(theta)
            self.retain_inputs((0, 1))
            b, c, h, w = x.shape
            oh, ow = self.output_shape
            y = xp.empty((b, c, oh, ow), dtype=xp.float32)
            with cuda.get_device_from_array(x):
                _lib.check(_lib.lib().loans_stn_crop_fwd(_ptr(x), _ptr(theta), self.mask01, _ptr(y), _ptr(self.grid_out),
                                                         b, 1, c, h, w, oh, ow, _lib.F32, _stream()), "loans_stn_crop_fwd")
            self.grid_out = None
            return y,

        def backward(self, indexes, grad_outputs):
            if not self.can_backprop:
                raise AttributeError("'RotationDropout' object has no attribute 'mask'")    # as the reference, :47-48
            xp = cuda.cupy
            x, theta = (xp.ascontiguousarray(v.data) for v in self.get_retained_inputs())
            gy = xp.ascontiguousarray(grad_outputs[0].data)
            b, c, h, w = x.shape
            oh, ow = self.output_shape
            need_gx = 0 in indexes                      # LoANs passes the frames as a raw array: never needed there
            gx = xp.empty_like(x) if need_gx else None
            gtheta = xp.empty_like(theta)
            flags = _lib.FLAG_UPRIGHT if self.upright else 0
            with cuda.get_device_from_array(x):
                _lib.check(_lib.lib().loans_stn_crop_bwd_ex(_ptr(x), _ptr(theta), self.mask01, _ptr(gy), None, None, _ptr(gtheta),
                                                            _ptr(gx), None, flags, b, 1, c, h, w, oh, ow, _lib.F32, _stream()),
                           "loans_stn_crop_bwd_ex")
            return (chainer.Variable(gx) if need_gx else None), chainer.Variable(gtheta)

    class SpatialTransformerSampler(function_node.FunctionNode):
        """explicit-grid sampler (any grid): loans_stn_sampler_fwd / _bwd."""

        def check_type_forward(self, in_types):
            type_check.expect(2 == in_types.size())
            x_type, grid_type = in_types
            type_check.expect(x_type.dtype.char == 'f', grid_type.dtype.char == 'f', x_type.ndim == 4, grid_type.ndim == 4,
                              grid_type.shape[1] == 2, x_type.shape[0] == grid_type.shape[0])

        def forward(self, inputs):
            x, grid = inputs
            _gpu_only(x, grid)
            xp = cuda.cupy
            x, grid = xp.ascontiguousarray(x), xp.ascontiguousarray(grid)
            self.retain_inputs((0, 1))
            b, c, h, w = x.shape
            _, _, oh, ow = grid.shape
            y = xp.empty((b, c, oh, ow), dtype=xp.float32)
            _lib.check(_lib.lib().loans_stn_sampler_fwd(_ptr(x), _ptr(grid), _ptr(y), b, 1, c, h, w, oh, ow, _lib.F32,
                                                        _stream()), "loans_stn_sampler_fwd")
            return y,

        def backward(self, indexes, grad_outputs):
            xp = cuda.cupy
            x, grid = (xp.ascontiguousarray(v.data) for v in self.get_retained_inputs())
            gy = xp.ascontiguousarray(grad_outputs[0].data)
            b, c, h, w = x.shape
            _, _, oh, ow = grid.shape
            gx = xp.empty_like(x) if 0 in indexes else None
            ggrid = xp.empty_like(grid) if 1 in indexes else None
            _lib.check(_lib.lib().loans_stn_sampler_bwd(_ptr(x), _ptr(grid), _ptr(gy), _ptr(gx), _ptr(ggrid), b, 1, c, h, w,
                                                        oh, ow, _lib.F32, _stream()), "loans_stn_sampler_bwd")
            return (None if gx is None else chainer.Variable(gx)), (None if ggrid is None else chainer.Variable(ggrid))

    def _no_kwargs(kwargs):
        if 'use_cudnn' in kwargs:
            raise ValueError("The argument \"use_cudnn\" is not supported anymore. Use "
                             "chainer.using_config('use_cudnn', value) context where value can be `always`, `never`, or `auto`.")
        if kwargs:
            raise TypeError('unexpected keyword arguments: %s' % ', '.join(sorted(kwargs)))

    def spatial_transformer_grid(theta, output_shape, **kwargs):
        _no_kwargs(kwargs)
        theta = theta if isinstance(theta, chainer.Variable) else chainer.Variable(theta)
        fuse = _STATE["fuse"]
        drop = getattr(theta, '_stn_dropout', None)
        if drop is not None and not (theta.data is drop["array"] and _ptr(theta.data) == drop["ptr"]
                                     and _ptr(drop["theta_in"].data) == drop["theta_in_ptr"]):
            drop = None                                              # not (or no longer) the array our dropout node returned
        ok = isinstance(theta.data, cuda.ndarray) and theta.data.dtype == cuda.cupy.float32
        defer = fuse == "full" and ok
        grid, = SpatialTransformerGrid(output_shape, defer=defer).apply((theta,))
        if fuse != "off" and ok:
            # note for the sampler: which theta this grid is the image of
            grid._stn_origin = {"array": grid.data, "ptr": _ptr(grid.data), "pending": defer,
                                "theta": theta, "theta_ptr": _ptr(theta.data), "dropout": drop}
        return grid

    def spatial_transformer_sampler(x, grid, **kwargs):
        _no_kwargs(kwargs)
        origin = getattr(grid, '_stn_origin', None)
        if origin is not None and grid.data is origin["array"] and _ptr(grid.data) == origin["ptr"] and \
                _ptr(origin["theta"].data) == origin["theta_ptr"]:
            drop, shape = origin["dropout"], grid.shape[2:]
            if origin["pending"]:
                origin["pending"] = False                            # the launch below fills the grid array
                if drop is not None:                                 # un-masked theta + the draw: one kernel, mask folded in
                    node = FusedSampler(shape, mask01=drop["mask01"], grid_out=grid.data, can_backprop=drop["can_backprop"])
                    return node.apply((x, drop["theta_in"]))[0]
                return FusedSampler(shape, grid_out=grid.data).apply((x, origin["theta"]))[0]
            upright = drop is not None and drop["mask01"] == 0.0
            return FusedSampler(shape, upright=upright).apply((x, origin["theta"]))[0]
        return SpatialTransformerSampler().apply((x, grid))[0]


if HAVE_CHAINER:

    def prepare_images(self, images):
        """Drop-in for SheepLocalizer.prepare_images / Resnet50SheepLocalizer.prepare_images (reference
        sheep/sheep_localizer.py:72-82): same input (`images.copy() * 255`, a Variable on the GPU), same output
        (a Variable holding the BGR, mean-subtracted float32 batch on the same device), no host round trip."""
        x = images.data if isinstance(images, chainer.Variable) else images
        _gpu_only(x)
        x = cuda.cupy.ascontiguousarray(x, dtype=cuda.cupy.float32)
        out = cuda.cupy.empty_like(x)
        b, c, h, w = x.shape
        with cuda.get_device_from_array(x):
            _lib.check(_lib.lib().loans_stn_prepare_images(_ptr(x), 1.0, _ptr(out), b, c, h, w, _stream()),
                       "loans_stn_prepare_images")
        return chainer.Variable(out)

    def patch_localizers(*classes):
        """After `from sheep.sheep_localizer import SheepLocalizer, Resnet50SheepLocalizer`:
        `patch_localizers(SheepLocalizer, Resnet50SheepLocalizer)` replaces their prepare_images method."""
        for cls in classes:
            cls.prepare_images = prepare_images


def install(fuse="full"):
    """Rebind the three operator names the reference uses.  Call before importing sheep.sheep_localizer.
    ``fuse``: "full" (default) / "sampler" / "off", see the module docstring."""
    _require()
    if fuse not in ("full", "sampler", "off"):
        raise ValueError("fuse must be 'full', 'sampler' or 'off'")
    _STATE["fuse"] = fuse
    import chainer.functions as F
    F.spatial_transformer_grid = spatial_transformer_grid
    F.spatial_transformer_sampler = spatial_transformer_sampler
    F.array.spatial_transformer_grid.spatial_transformer_grid = spatial_transformer_grid
    F.array.spatial_transformer_sampler.spatial_transformer_sampler = spatial_transformer_sampler
    mod = types.ModuleType("functions.rotation_droput")        # sic: the typo is the reference's import path
    mod.rotation_dropout = rotation_dropout
    mod.RotationDropout = RotationDropout
    pkg = sys.modules.setdefault("functions", types.ModuleType("functions"))
    pkg.rotation_droput = mod
    sys.modules["functions.rotation_droput"] = mod
