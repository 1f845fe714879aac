import contextlib


class _Config(object):
    train = True
    use_cudnn = "auto"


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)
