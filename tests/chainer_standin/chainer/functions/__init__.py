from chainer.functions import array                                                        # noqa: F401
from chainer.functions.array.reshape import reshape                                          # noqa: F401
from chainer.functions.array.spatial_transformer_grid import spatial_transformer_grid        # noqa: F401
from chainer.functions.array.spatial_transformer_sampler import spatial_transformer_sampler  # noqa: F401
