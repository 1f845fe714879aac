"""Stand-in for the handful of cupy names loans_b200/chainer_compat.py touches (tests only; see ../README.md)."""
import numpy
import torch

__standin__ = True
float32 = numpy.float32
float64 = numpy.float64
_NP2T = {numpy.dtype("float32"): torch.float32, numpy.dtype("float64"): torch.float64, numpy.dtype("float16"): torch.float16}
_T2NP = {v: k for k, v in _NP2T.items()}


class _MemoryPointer(object):
    def __init__(self, t):
        self.ptr = t.data_ptr()


class ndarray(object):
    def __init__(self, t):
        assert t.is_cuda
        self._t = t

    data = property(lambda self: _MemoryPointer(self._t))
    shape = property(lambda self: tuple(self._t.shape))
    dtype = property(lambda self: _T2NP[self._t.dtype])
    ndim = property(lambda self: self._t.dim())
    size = property(lambda self: self._t.numel())
    device = property(lambda self: cuda.Device(self._t.device.index))

    def copy(self):
        return ndarray(self._t.clone())

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return ndarray(self._t.reshape(shape))

    def get(self):
        return self._t.cpu().numpy()

    def _other(self, o):
        return o._t if isinstance(o, ndarray) else o

    def __add__(self, o):
        return ndarray(self._t + self._other(o))

    __radd__ = __add__

    def __mul__(self, o):
        return ndarray(self._t * self._other(o))

    __rmul__ = __mul__


def _tdtype(dtype):
    return _NP2T[numpy.dtype(dtype)]


def empty(shape, dtype=float32):
    return ndarray(torch.empty(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_tdtype(dtype), device="cuda"))


def empty_like(a):
    return ndarray(torch.empty_like(a._t))


def asarray(a, dtype=None):
    if isinstance(a, ndarray):
        return a if dtype is None or numpy.dtype(dtype) == a.dtype else ndarray(a._t.to(_tdtype(dtype)))
    t = torch.from_numpy(numpy.ascontiguousarray(a)).cuda()
    return ndarray(t if dtype is None else t.to(_tdtype(dtype)))


def ascontiguousarray(a, dtype=None):
    a = asarray(a, dtype)
    return a if a._t.is_contiguous() else ndarray(a._t.contiguous())


def asnumpy(a):
    return a.get() if isinstance(a, ndarray) else numpy.asarray(a)


class random(object):
    @staticmethod
    def rand(*shape):
        # the draw itself happens on the host in this stand-in (cupy would draw on the device from its own stream)
        return numpy.random.rand(*shape)


class _Stream(object):
    def __init__(self, ptr):
        self.ptr = ptr


class cuda(object):
    class Device(object):
        def __init__(self, index=0):
            self.id = 0 if index is None else int(index)

        def __enter__(self):
            self._prev = torch.cuda.current_device()
            torch.cuda.set_device(self.id)
            return self

        def __exit__(self, *exc):
            torch.cuda.set_device(self._prev)

    @staticmethod
    def get_current_stream():
        return _Stream(torch.cuda.current_stream().cuda_stream)
