"""``spatial_transformer_grid`` / ``spatial_transformer_sampler`` / fused ``stn_crop`` on CUDA tensors.

Front ends of libloans_stn.so with the signatures of ``chainer.functions.spatial_transformer_grid(theta,
output_shape)`` and ``chainer.functions.spatial_transformer_sampler(x, grid)`` as the reference calls them
(sheep/sheep_localizer.py:62-63, :170-171).  torch supplies device memory, the stream and the autograd
tape; every number is produced by the hand-written kernels behind the C ABI.  No CPU path.

The reference's three calls run as ONE kernel per direction
----------------------------------------------------------
    theta  = rotation_dropout(theta, ratio=0.0)          # sheep/sheep_localizer.py:61
    points = F.spatial_transformer_grid(theta, out_size)  # :62
    rois   = F.spatial_transformer_sampler(images, points) # :63

``rotation_dropout`` and ``spatial_transformer_grid`` return DEFERRED tensors (``Deferred``, a ``torch.Tensor``
wrapper subclass): shape, dtype and device are there, the values are not computed yet.  The sampler, handed a grid
nobody has looked at, launches the fused kernel ``loans_stn_crop_fwd(theta_in, mask01=<the dropout draw>)`` once: it
writes the crops AND the dense grid (which is what ``points`` then holds), and one autograd node with the two outputs
``(rois, points)`` sends the gradient of both -- the assessor's ``gy`` and whatever the corner regularisers put on
``points`` -- through ONE ``loans_stn_crop_bwd`` launch to the un-masked theta.  Exactly the launches of ``stn_crop``.

Anything else that touches a deferred tensor (any torch function, method or operator on it) first materialises it
with the unfused kernel of its own operator (``loans_stn_rotation_dropout`` / ``loans_stn_grid_fwd``), and the
sampler then sees an ordinary tensor: a grid that is still the unmodified output of our grid node is sampled by the
fused kernel from theta (its gradient travels through ``points`` like any other, so hooks and ``retain_grad`` on
``points`` see it); any other grid takes the explicit-grid sampler.  The fusion never guesses: it is taken only when
the deferred grid was never materialised, and it raises if theta was modified in place after it was handed to
``rotation_dropout`` / ``spatial_transformer_grid`` (the deferred operators would otherwise read the new values).
``loans_b200.config.defer = False`` switches the deferral off (three eager nodes).
"""
import torch

from loans_b200 import _lib
from loans_b200.configuration import config


class InvalidType(TypeError):
    """Stand-in for ``chainer.utils.type_check.InvalidType`` (raised by ``check_type_forward``)."""


_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _expect(cond, msg):
    if not cond:
        raise InvalidType(msg)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("loans_b200 operators run on CUDA tensors only (no CPU fallback); got a %s tensor"
                               % t.device.type)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_device(t):
    """Context: the tensor's device is the current CUDA device (the C ABI launches on the current device)."""
    return torch.cuda.device(t.device)


def _out_hw(output_shape):
    # a 2-sequence (H, W): tuple from argparse or list from the JSON log (reference evaluate.py:42,46)
    oh, ow = output_shape
    return int(oh), int(ow)


# ------------------------------------------------------------------------------------------ deferred tensors
def _meta_getters():
    T = torch.Tensor
    fns = {T.size, T.dim, T.numel, T.element_size, T.is_floating_point, T.__len__, T.stride, T.ndimension, T.nelement,
           T.is_contiguous, T.get_device, T.type}
    for name in ("shape", "dtype", "device", "ndim", "is_cuda", "requires_grad", "layout", "names", "is_sparse",
                 "is_quantized", "is_meta", "is_cpu"):
        d = getattr(T, name, None)
        if d is not None and hasattr(d, "__get__"):
            fns.add(d.__get__)
    return fns


def _alias_getters():
    # ways to obtain an alias whose in-place edits do NOT move the tensor's version counter: once one of these has been
    # used on a grid, "unmodified output of our grid node" can no longer be verified and the note for the sampler is dropped
    T = torch.Tensor
    fns = set()
    for name in ("data", "__cuda_array_interface__"):
        d = getattr(T, name, None)
        if d is not None and hasattr(d, "__get__"):
            fns.add(d.__get__)
        if d is not None and hasattr(d, "__set__"):
            fns.add(d.__set__)
    for name in ("__dlpack__", "untyped_storage", "set_", "numpy", "__array__"):
        if hasattr(T, name):
            fns.add(getattr(T, name))
    return fns


class Deferred(torch.Tensor):
    """A tensor whose values are computed on first use.

    Metadata (shape, dtype, device, ...) is answered from the wrapper; every other torch function, method or operator
    resolves it -- ``thunk()`` launches the operator's own kernel, its result replaces the wrapper in the call -- so the
    object behaves like the tensor the eager operator would have returned.  ``note`` is what the next operator of the
    chain needs to fuse instead (never inspected by anything else).
    """
    _META = None
    _ALIAS = None

    @staticmethod
    def __new__(cls, shape, dtype, device, requires_grad, thunk, note):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=bool(requires_grad))
        t._thunk = thunk
        t._real = None
        t._note = note
        return t

    def __init__(self, *a, **k):
        pass

    @property
    def pending(self):
        return self._real is None

    def resolve(self):
        """The ordinary tensor behind the wrapper (materialised now if it was still pending)."""
        if self._real is None:
            self._real = self._thunk()
            self._thunk = None
        return self._real

    def _bind(self, real):
        """The next operator of the chain produced the values itself (fused kernel): adopt them."""
        self._real = real
        self._thunk = None

    def __repr__(self):
        if self._real is None:
            return "Deferred(%s, shape=%s, pending)" % (self._note[0], tuple(self.shape))
        return repr(self._real)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if cls._META is None:
            cls._META, cls._ALIAS = _meta_getters(), _alias_getters()
        if func in cls._META:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        untracked = func in cls._ALIAS

        def unwrap(a):
            if isinstance(a, Deferred):
                real = a.resolve()
                if untracked and getattr(real, "_stn_origin", None) is not None:
                    real._stn_origin = None
                return real
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(v) for v in a)
            if isinstance(a, dict):
                return {k: unwrap(v) for k, v in a.items()}
            return a

        args, kwargs = unwrap(args), unwrap(kwargs)
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # safety net (a wrapper reached an ATen operator without passing __torch_function__): same rule, resolve and go on
        import torch.utils._pytree as pytree
        args, kwargs = pytree.tree_map_only(Deferred, lambda a: a.resolve(), (args, kwargs or {}))
        return func(*args, **kwargs)


def resolve(t):
    """``t`` as an ordinary tensor (materialising it if it is a pending ``Deferred``)."""
    return t.resolve() if isinstance(t, Deferred) else t


def _unchanged(t, version, ptr):
    return t._version == version and t.data_ptr() == ptr


def _modified_error(what):
    return RuntimeError(
        "%s was modified in place after it was handed to a deferred loans_b200 operator (rotation_dropout / "
        "spatial_transformer_grid compute on first use); clone() it before the in-place update, or set "
        "loans_b200.config.defer = False" % what)


# ------------------------------------------------------------------------------------------ fused a5
class _StnCrop(torch.autograd.Function):
    """One fused forward launch, one fused backward launch.  ``opt`` = dict(mask01, oh, ow, k, out_dtype, points, gray,
    nhwc, upright, can_backprop); inputs (x, theta) or, with ``via_grid``, (x, theta_detached, grid): the grid is then a
    real autograd input whose gradient is the per-pixel grid gradient (theta gets none from this node)."""

    @staticmethod
    def forward(ctx, x, theta, grid_in, opt):
        # points: 0 = crops only, 1 = crops + dense grid (N,2,oH,oW), 2 = crops + the grid's four corner points (N,2,2,2)
        x = x.contiguous()
        theta = theta.contiguous()
        b, c, h, w = x.shape
        n = theta.shape[0]
        oh, ow, k, out_dtype, points = opt["oh"], opt["ow"], opt["k"], opt["out_dtype"], opt["points"]
        gray, nhwc, mask01 = opt["gray"], opt["nhwc"], opt["mask01"]
        yshape = (n, oh, ow, 4) if nhwc else (n, 1 if gray else c, oh, ow)
        y = torch.empty(yshape, dtype=out_dtype, device=x.device)
        grid = (torch.empty((n, 2, 2, 2), dtype=torch.float32, device=x.device) if points == 2 else
                torch.empty((n, 2, oh, ow), dtype=torch.float32, device=x.device) if points == 1 else None)
        flags = (_lib.FLAG_GRAY if gray else 0) | (_lib.FLAG_NHWC4 if nhwc else 0)
        with _on_device(x):
            _lib.check(_lib.lib().loans_stn_crop_fwd_ex(_ptr(x), _ptr(theta), mask01, _ptr(y),
                                                        _ptr(grid) if points == 1 else None, _ptr(grid) if points == 2 else None,
                                                        flags, n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                       "loans_stn_crop_fwd_ex")
        ctx.save_for_backward(x, theta)
        ctx.opt = opt
        ctx.via_grid = grid_in is not None
        ctx.set_materialize_grads(False)          # no gradient on points -> NULL, not a dense zero array read back
        if grid is None:
            return y
        return y, grid

    @staticmethod
    def backward(ctx, gy, ggrid=None):
        x, theta = ctx.saved_tensors
        opt = ctx.opt
        oh, ow, k, out_dtype, points = opt["oh"], opt["ow"], opt["k"], opt["out_dtype"], opt["points"]
        gray, nhwc, mask01 = opt["gray"], opt["nhwc"], opt["mask01"]
        if not opt["can_backprop"]:
            # reference functions/rotation_droput.py:47-48 multiplies by self.mask, which a test-mode forward never creates
            raise AttributeError("'RotationDropout' object has no attribute 'mask' "
                                 "(backward after a test-mode forward, as in the reference)")
        b, c, h, w = x.shape
        n = theta.shape[0]
        need_gx = ctx.needs_input_grad[0]
        via_grid = ctx.via_grid
        gtheta = torch.empty_like(theta)
        gx = torch.empty_like(x) if need_gx else None
        yshape = (n, oh, ow, 4) if nhwc else (n, 1 if gray else c, oh, ow)
        if gy is None:                             # rois unused downstream: only the gradient on points arrives
            gy = torch.zeros(yshape, dtype=out_dtype, device=x.device)
        gy = gy.contiguous()
        if gy.dtype != out_dtype:
            gy = gy.to(out_dtype)
        if ggrid is not None:
            ggrid = ggrid.contiguous().float()
        ggrid_out = None
        if via_grid and ctx.needs_input_grad[2]:
            ggrid_out = torch.empty((n, 2, oh, ow), dtype=torch.float32, device=x.device)
        flags = (_lib.FLAG_GRAY if gray else 0) | (_lib.FLAG_NHWC4 if nhwc else 0) | (_lib.FLAG_UPRIGHT if opt["upright"] else 0)
        with _on_device(x):
            _lib.check(_lib.lib().loans_stn_crop_bwd_ex(_ptr(x), _ptr(theta), mask01, _ptr(gy),
                                                        _ptr(ggrid) if points == 1 else None, _ptr(ggrid) if points == 2 else None,
                                                        _ptr(gtheta), _ptr(gx), _ptr(ggrid_out), flags,
                                                        n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                       "loans_stn_crop_bwd_ex")
        if via_grid:
            return gx, None, ggrid_out, None
        return gx, gtheta, None, None


def _check_sampler_types(x, theta_or_grid, grid_like):
    _expect(x.dtype == torch.float32, "x.dtype.char == 'f' (got %s)" % x.dtype)
    _expect(theta_or_grid.dtype == torch.float32, "%s must be float32 (got %s)"
            % ("grid" if grid_like else "theta", theta_or_grid.dtype))
    _expect(x.dim() == 4, "x.ndim == 4 (got %d)" % x.dim())


def _crop_options(mask01, oh, ow, k=1, out_dtype=torch.float32, points=1, gray=False, nhwc=False, upright=False,
                  can_backprop=True):
    return {"mask01": float(mask01), "oh": int(oh), "ow": int(ow), "k": int(k), "out_dtype": out_dtype, "points": int(points),
            "gray": bool(gray), "nhwc": bool(nhwc), "upright": bool(upright), "can_backprop": bool(can_backprop)}


def stn_crop(x, theta, output_shape, ratio=None, crops_per_frame=1, out_dtype=torch.float32,
             return_grid=True, mask01=None, points="grid", grayscale=False, layout="nchw"):
    """rotation_dropout(theta, ratio) -> grid -> sampler in one kernel (reference sheep/sheep_localizer.py:61-63).

    x (B,C,H,W) float32 frames; theta (B*K,2,3) float32; ``ratio`` as in ``rotation_dropout`` (``None``: no
    dropout node in front of the grid); ``mask01`` overrides the drawn mask value (tests, data-parallel runs
    that broadcast one draw to all ranks).  Returns ``(rois, points)`` like the localizer does (``rois`` only
    if ``return_grid`` is false).  ``out_dtype`` torch.float32 or torch.bfloat16 (= bf16(fp32 result)).
    ``points="corners"``: ``points`` is the grid at its four corners only, shape (N,2,2,2) -- bit-identical to
    ``grid[:, :, [0,-1]][:, :, :, [0,-1]]`` and a valid ``points`` array for everything LoANs does with it besides
    sampling (regularisers, extract_corners, evaluator: they index ``[0,0]``, ``[0,width-1]``, ``[height-1,0]``,
    ``[-1,-1]`` with height and width taken from its shape); the dense grid is never written or read back.
    ``grayscale=True`` (3-channel frames): the localizer's ``transform_rois_to_grayscale`` epilogue
    (sheep/sheep_localizer.py:65-68) fused in; ``rois`` is (N,1,oH,oW) = ``0.299*ch2 + 0.587*ch1 + 0.114*ch0``.
    ``layout="nhwc4"`` (3 channels, bf16): ``rois`` (and the gradient handed back for them) are channels-last with the
    channel count padded to four, (N,oH,oW,4) -- the layout the assessor's first convolution consumes on tensor cores
    (reference common/net.py:15-25 is the consumer); the padding channel is written as zero.
    """
    from loans_b200.functions.rotation_droput import draw_mask_value
    x, theta = resolve(x), resolve(theta)
    _need_cuda(x, theta)
    _check_sampler_types(x, theta, False)
    _expect(theta.dim() == 3 and theta.shape[1] == 2 and theta.shape[2] == 3, "theta.shape == (N, 2, 3)")
    k = int(crops_per_frame)
    _expect(k >= 1 and theta.shape[0] == x.shape[0] * k,
            "theta.shape[0] == x.shape[0] * crops_per_frame (%d vs %d*%d)" % (theta.shape[0], x.shape[0], k))
    if out_dtype not in _DT:
        raise InvalidType("out_dtype must be torch.float32 or torch.bfloat16")
    oh, ow = _out_hw(output_shape)
    if mask01 is None:
        mask01 = 1.0 if ratio is None else draw_mask_value(ratio)
    if points not in ("grid", "corners"):
        raise ValueError("points must be 'grid' or 'corners'")
    if layout not in ("nchw", "nhwc4"):
        raise ValueError("layout must be 'nchw' or 'nhwc4'")
    mode = 0 if not return_grid else (2 if points == "corners" else 1)
    if grayscale:
        _expect(x.shape[1] == 3, "rois are not in RGB, can not convert them to grayscale (C == %d)" % x.shape[1])
    if layout == "nhwc4":
        _expect(x.shape[1] == 3 and out_dtype == torch.bfloat16 and not grayscale,
                "layout='nhwc4' needs 3-channel frames, out_dtype=torch.bfloat16 and no grayscale epilogue")
    opt = _crop_options(mask01, oh, ow, k, out_dtype, mode, grayscale, layout == "nhwc4")
    return _StnCrop.apply(x, theta, None, opt)


# ------------------------------------------------------------------------------------------ a2
class _Grid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, oh, ow):
        theta = theta.contiguous()
        n = theta.shape[0]
        grid = torch.empty((n, 2, oh, ow), dtype=torch.float32, device=theta.device)
        with _on_device(theta):
            _lib.check(_lib.lib().loans_stn_grid_fwd(_ptr(theta), _ptr(grid), n, oh, ow, _stream()), "loans_stn_grid_fwd")
        ctx.meta = (n, oh, ow)
        return grid

    @staticmethod
    def backward(ctx, ggrid):
        n, oh, ow = ctx.meta
        ggrid = ggrid.contiguous().float()
        gtheta = torch.empty((n, 2, 3), dtype=torch.float32, device=ggrid.device)
        with _on_device(ggrid):
            _lib.check(_lib.lib().loans_stn_grid_bwd(_ptr(ggrid), _ptr(gtheta), n, oh, ow, _stream()), "loans_stn_grid_bwd")
        return gtheta, None, None


def _reject_kwargs(kwargs):
    if "use_cudnn" in kwargs:
        raise ValueError("The argument \"use_cudnn\" is not supported anymore. "
                         "Use chainer.using_config('use_cudnn', value) context where value can be `always`, "
                         "`never`, or `auto`.")
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % ", ".join(sorted(kwargs)))


def _eager_grid(theta, oh, ow, upright, note):
    """The grid node, eagerly: launches loans_stn_grid_fwd.  ``note``: leave a note for the sampler on the result -- only
    for grids that live behind a ``Deferred`` wrapper, which sees (and voids the note on) every access that could edit
    the values without moving the version counter (``.data`` and friends)."""
    grid = _Grid.apply(theta, oh, ow)
    if note:
        # which theta this grid is the unmodified image of -- checked again (versions, pointers) before it is trusted
        grid._stn_origin = {"theta": theta, "theta_version": theta._version, "theta_ptr": theta.data_ptr(),
                            "grid_version": grid._version, "grid_ptr": grid.data_ptr(), "upright": bool(upright)}
    return grid


def spatial_transformer_grid(theta, output_shape, **kwargs):
    """theta (B,2,3) float32 -> grid (B,2,H,W): ``chainer.functions.spatial_transformer_grid``."""
    _reject_kwargs(kwargs)
    _need_cuda(theta)
    _expect(theta.dtype == torch.float32, "theta.dtype.char == 'f' (got %s)" % theta.dtype)
    _expect(theta.dim() == 3, "theta.ndim == 3 (got %d)" % theta.dim())
    _expect(theta.shape[1] == 2 and theta.shape[2] == 3, "theta.shape[1:] == (2, 3)")
    oh, ow = _out_hw(output_shape)
    _expect(oh >= 1 and ow >= 1, "output_shape must be positive (got %s)" % (tuple(output_shape),))
    # where theta comes from: the pending output of our rotation dropout (then theta_in and the draw are what the fused
    # kernel needs), or an ordinary tensor
    dropout = None
    if isinstance(theta, Deferred):
        if theta.pending and theta._note[0] == "rotation_dropout":
            dropout = theta
        else:
            theta = theta.resolve()
    if not config.defer:
        th = dropout.resolve() if dropout is not None else theta
        return _eager_grid(th, oh, ow, False, note=False)
    if dropout is not None:
        src = dict(dropout._note[1])                  # theta_in, its version / pointer, mask01, can_backprop
        src["dropout"] = dropout
    else:
        src = {"theta": theta, "theta_version": theta._version, "theta_ptr": theta.data_ptr(), "mask01": 1.0,
               "can_backprop": True, "dropout": None}
    n = theta.shape[0]

    def materialise():
        d = src["dropout"]
        if d is not None:
            return _eager_grid(d.resolve(), oh, ow, src["mask01"] == 0.0, note=True)
        if not _unchanged(src["theta"], src["theta_version"], src["theta_ptr"]):
            raise _modified_error("theta")
        return _eager_grid(src["theta"], oh, ow, False, note=True)

    return Deferred((n, 2, oh, ow), torch.float32, theta.device, src["theta"].requires_grad, materialise, ("grid", src, oh, ow))


# ------------------------------------------------------------------------------------------ a3/a4
class _SamplerExplicit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grid):
        x = x.contiguous()
        grid = grid.contiguous()
        b, c, h, w = x.shape
        n, _, oh, ow = grid.shape
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        with _on_device(x):
            _lib.check(_lib.lib().loans_stn_sampler_fwd(_ptr(x), _ptr(grid), _ptr(y), n, 1, c, h, w, oh, ow,
                                                        _lib.F32, _stream()), "loans_stn_sampler_fwd")
        ctx.save_for_backward(x, grid)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, grid = ctx.saved_tensors
        b, c, h, w = x.shape
        n, _, oh, ow = grid.shape
        gy = gy.contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        ggrid = torch.empty_like(grid) if ctx.needs_input_grad[1] else None
        with _on_device(x):
            _lib.check(_lib.lib().loans_stn_sampler_bwd(_ptr(x), _ptr(grid), _ptr(gy), _ptr(gx), _ptr(ggrid),
                                                        n, 1, c, h, w, oh, ow, _lib.F32, _stream()),
                       "loans_stn_sampler_bwd")
        return gx, ggrid


def spatial_transformer_sampler(x, grid, **kwargs):
    """x (B,C,H,W), grid (B,2,oH,oW) float32 -> (B,C,oH,oW): ``chainer.functions.spatial_transformer_sampler``."""
    _reject_kwargs(kwargs)
    x = resolve(x)
    _need_cuda(x, grid)
    _check_sampler_types(x, grid, True)
    _expect(grid.dim() == 4, "grid.ndim == 4 (got %d)" % grid.dim())
    _expect(grid.shape[1] == 2, "grid.shape[1] == 2")
    _expect(x.shape[0] == grid.shape[0], "x.shape[0] == grid.shape[0] (%d vs %d)" % (x.shape[0], grid.shape[0]))
    if isinstance(grid, Deferred):
        if grid.pending and grid._note[0] == "grid":
            # nobody has looked at the grid since our grid node promised it: ONE fused kernel writes the crops and the
            # grid; one autograd node, two outputs -- gy and the gradient on points share one backward launch
            _, src, oh, ow = grid._note
            if not _unchanged(src["theta"], src["theta_version"], src["theta_ptr"]):
                raise _modified_error("theta")
            opt = _crop_options(src["mask01"], oh, ow, points=1, can_backprop=src["can_backprop"])
            rois, points = _StnCrop.apply(x, src["theta"], None, opt)
            grid._bind(points)
            return rois
        grid = grid.resolve()
    origin = getattr(grid, "_stn_origin", None)
    if origin is not None and _unchanged(grid, origin["grid_version"], origin["grid_ptr"]) and \
            _unchanged(origin["theta"], origin["theta_version"], origin["theta_ptr"]) and origin["theta"].shape[0] == grid.shape[0]:
        # an unmodified grid of our grid node that something else materialised first: the coordinates are recomputed from
        # theta in registers (the grid is not read back), the gradient still travels through `grid` like any other
        opt = _crop_options(1.0, grid.shape[2], grid.shape[3], points=0, upright=origin["upright"])
        return _StnCrop.apply(x, origin["theta"].detach(), grid, opt)
    return _SamplerExplicit.apply(x, grid)
