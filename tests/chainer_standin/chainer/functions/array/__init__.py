from chainer.functions.array import reshape, spatial_transformer_grid, spatial_transformer_sampler   # noqa: F401
