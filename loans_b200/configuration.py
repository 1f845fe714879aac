"""Thread-local train/test switch, the stand-in for ``chainer.config.train``.

The reference's rotation dropout reads exactly one flag, ``chainer.configuration.config.train``
(reference functions/rotation_droput.py:30), set with ``chainer.using_config('train', False)`` at
sheep/sheep_localizer.py:103.  Same names here so call sites read the same.
"""
import contextlib
import threading


class _Config(threading.local):
    train = True
    # rotation_dropout / spatial_transformer_grid return deferred tensors so that the reference's three calls run as one
    # fused kernel per direction (loans_b200/functions/spatial_transformer.py); False: three eager operator nodes
    defer = True


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    if not hasattr(config, name):
        raise AttributeError("unknown configuration entry %r" % name)
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)
