"""``spatial_transformer_grid`` / ``spatial_transformer_sampler`` / fused ``stn_crop`` on CUDA tensors.

Front ends of libloans_stn.so with the signatures of ``chainer.functions.spatial_transformer_grid(theta,
output_shape)`` and ``chainer.functions.spatial_transformer_sampler(x, grid)`` as the reference calls them
(sheep/sheep_localizer.py:62-63, :170-171).  torch supplies device memory, the stream and the autograd
tape; every number is produced by the hand-written kernels behind the C ABI.  No CPU path.

Calling the two functions one after the other, as the reference does, still runs the fused sampler: the
grid node leaves a note on the tensor it returns (which theta it came from), and the sampler node, when
handed that very tensor unmodified, recomputes the coordinates from theta in registers instead of
reading the grid back, and sends its gradient straight to theta.
"""
import torch

from loans_b200 import _lib


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


def _out_hw(output_shape):
    # a 2-sequence (H, W): tuple from argparse or list from the JSON log (reference evaluate.py:42,46)
    oh, ow = output_shape
    return int(oh), int(ow)


# ------------------------------------------------------------------------------------------ fused a5
class _StnCrop(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, theta, mask01, oh, ow, k, out_dtype, points, gray=False):
        # points: 0 = crops only, 1 = crops + dense grid (N,2,oH,oW), 2 = crops + the grid's four corner points (N,2,2,2)
        x = x.contiguous()
        theta = theta.contiguous()
        b, c, h, w = x.shape
        n = theta.shape[0]
        y = torch.empty((n, 1 if gray else c, oh, ow), dtype=out_dtype, device=x.device)
        grid = None
        with torch.cuda.device(x.device):
            if gray:
                grid = (torch.empty((n, 2, 2, 2), dtype=torch.float32, device=x.device) if points == 2 else
                        torch.empty((n, 2, oh, ow), dtype=torch.float32, device=x.device) if points == 1 else None)
                _lib.check(_lib.lib().loans_stn_crop_fwd_ex(_ptr(x), _ptr(theta), mask01, _ptr(y),
                                                            _ptr(grid) if points == 1 else None, _ptr(grid) if points == 2 else None,
                                                            _lib.FLAG_GRAY, n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_fwd_ex")
            elif points == 2:
                grid = torch.empty((n, 2, 2, 2), dtype=torch.float32, device=x.device)
                _lib.check(_lib.lib().loans_stn_crop_fwd_corners(_ptr(x), _ptr(theta), mask01, _ptr(y), _ptr(grid),
                                                                 n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_fwd_corners")
            else:
                grid = torch.empty((n, 2, oh, ow), dtype=torch.float32, device=x.device) if points == 1 else None
                _lib.check(_lib.lib().loans_stn_crop_fwd(_ptr(x), _ptr(theta), mask01, _ptr(y), _ptr(grid),
                                                         n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_fwd")
        ctx.save_for_backward(x, theta)
        ctx.meta = (mask01, oh, ow, k, out_dtype, points, gray)
        if grid is None:
            return y
        return y, grid

    @staticmethod
    def backward(ctx, gy, ggrid=None):
        x, theta = ctx.saved_tensors
        mask01, oh, ow, k, out_dtype, points, gray = ctx.meta
        b, c, h, w = x.shape
        n = theta.shape[0]
        need_gx = ctx.needs_input_grad[0]
        gtheta = torch.empty_like(theta)
        gx = torch.empty_like(x) if need_gx else None
        if gy is None:
            gy = torch.zeros((n, 1 if gray else c, oh, ow), dtype=out_dtype, device=x.device)
        gy = gy.contiguous()
        if gy.dtype != out_dtype:
            gy = gy.to(out_dtype)
        if ggrid is not None:
            ggrid = ggrid.contiguous().float()
        with torch.cuda.device(x.device):
            if gray:
                _lib.check(_lib.lib().loans_stn_crop_bwd_ex(_ptr(x), _ptr(theta), mask01, _ptr(gy),
                                                            _ptr(ggrid) if points == 1 else None, _ptr(ggrid) if points == 2 else None,
                                                            _ptr(gtheta), _ptr(gx), None, _lib.FLAG_GRAY,
                                                            n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_bwd_ex")
            elif points == 2:
                _lib.check(_lib.lib().loans_stn_crop_bwd_corners(_ptr(x), _ptr(theta), mask01, _ptr(gy), _ptr(ggrid),
                                                                 _ptr(gtheta), _ptr(gx),
                                                                 n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_bwd_corners")
            else:
                _lib.check(_lib.lib().loans_stn_crop_bwd(_ptr(x), _ptr(theta), mask01, _ptr(gy), _ptr(ggrid),
                                                         _ptr(gtheta), _ptr(gx), None,
                                                         n, k, c, h, w, oh, ow, _DT[out_dtype], _stream()),
                           "loans_stn_crop_bwd")
        return gx, gtheta, None, None, None, None, None, None, None


def _check_sampler_types(x, theta_or_grid, grid_like):
    _expect(x.dtype == torch.float32, "x.dtype.char == 'f' (got %s)" % x.dtype)
    _expect(theta_or_grid.dtype == torch.float32, "%s must be float32 (got %s)"
            % ("grid" if grid_like else "theta", theta_or_grid.dtype))
    _expect(x.dim() == 4, "x.ndim == 4 (got %d)" % x.dim())


def stn_crop(x, theta, output_shape, ratio=None, crops_per_frame=1, out_dtype=torch.float32,
             return_grid=True, mask01=None, points="grid", grayscale=False):
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
    """
    from loans_b200.functions.rotation_droput import draw_mask_value
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
    mode = 0 if not return_grid else (2 if points == "corners" else 1)
    if grayscale:
        _expect(x.shape[1] == 3, "rois are not in RGB, can not convert them to grayscale (C == %d)" % x.shape[1])
    return _StnCrop.apply(x, theta, float(mask01), oh, ow, k, out_dtype, mode, bool(grayscale))


# ------------------------------------------------------------------------------------------ a2
class _Grid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, oh, ow):
        theta = theta.contiguous()
        n = theta.shape[0]
        grid = torch.empty((n, 2, oh, ow), dtype=torch.float32, device=theta.device)
        with torch.cuda.device(theta.device):
            _lib.check(_lib.lib().loans_stn_grid_fwd(_ptr(theta), _ptr(grid), n, oh, ow, _stream()), "loans_stn_grid_fwd")
        ctx.meta = (n, oh, ow)
        return grid

    @staticmethod
    def backward(ctx, ggrid):
        n, oh, ow = ctx.meta
        ggrid = ggrid.contiguous().float()
        gtheta = torch.empty((n, 2, 3), dtype=torch.float32, device=ggrid.device)
        with torch.cuda.device(ggrid.device):
            _lib.check(_lib.lib().loans_stn_grid_bwd(_ptr(ggrid), _ptr(gtheta), n, oh, ow, _stream()), "loans_stn_grid_bwd")
        return gtheta, None, None


def _reject_kwargs(kwargs):
    if "use_cudnn" in kwargs:
        raise ValueError("The argument \"use_cudnn\" is not supported anymore. "
                         "Use chainer.using_config('use_cudnn', value) context where value can be `always`, "
                         "`never`, or `auto`.")
    if kwargs:
        raise TypeError("unexpected keyword arguments: %s" % ", ".join(sorted(kwargs)))


def spatial_transformer_grid(theta, output_shape, **kwargs):
    """theta (B,2,3) float32 -> grid (B,2,H,W): ``chainer.functions.spatial_transformer_grid``."""
    _reject_kwargs(kwargs)
    _need_cuda(theta)
    _expect(theta.dtype == torch.float32, "theta.dtype.char == 'f' (got %s)" % theta.dtype)
    _expect(theta.dim() == 3, "theta.ndim == 3 (got %d)" % theta.dim())
    _expect(theta.shape[1] == 2 and theta.shape[2] == 3, "theta.shape[1:] == (2, 3)")
    oh, ow = _out_hw(output_shape)
    grid = _Grid.apply(theta, oh, ow)
    # note for the sampler node: where this grid came from (see module docstring)
    grid._stn_origin = (theta, grid._version)
    return grid


# ------------------------------------------------------------------------------------------ a3/a4
class _SamplerExplicit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grid):
        x = x.contiguous()
        grid = grid.contiguous()
        b, c, h, w = x.shape
        n, _, oh, ow = grid.shape
        y = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
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
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().loans_stn_sampler_bwd(_ptr(x), _ptr(grid), _ptr(gy), _ptr(gx), _ptr(ggrid),
                                                        n, 1, c, h, w, oh, ow, _lib.F32, _stream()),
                       "loans_stn_sampler_bwd")
        return gx, ggrid


def spatial_transformer_sampler(x, grid, **kwargs):
    """x (B,C,H,W), grid (B,2,oH,oW) float32 -> (B,C,oH,oW): ``chainer.functions.spatial_transformer_sampler``."""
    _reject_kwargs(kwargs)
    _need_cuda(x, grid)
    _check_sampler_types(x, grid, True)
    _expect(grid.dim() == 4, "grid.ndim == 4 (got %d)" % grid.dim())
    _expect(grid.shape[1] == 2, "grid.shape[1] == 2")
    _expect(x.shape[0] == grid.shape[0], "x.shape[0] == grid.shape[0] (%d vs %d)" % (x.shape[0], grid.shape[0]))
    origin = getattr(grid, "_stn_origin", None)
    if origin is not None and origin[1] == grid._version and origin[0].shape[0] == grid.shape[0]:
        theta = origin[0]
        return _StnCrop.apply(x, theta, 1.0, grid.shape[2], grid.shape[3], 1, torch.float32, False)
    return _SamplerExplicit.apply(x, grid)
