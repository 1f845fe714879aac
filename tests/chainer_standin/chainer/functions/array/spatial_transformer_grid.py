def spatial_transformer_grid(*args, **kwargs):
    raise NotImplementedError("stand-in: chainer's own spatial_transformer_grid is not restated here; "
                              "loans_b200.chainer_compat.install() rebinds this name")
