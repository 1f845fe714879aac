from chainer.utils import type_check            # noqa: F401
