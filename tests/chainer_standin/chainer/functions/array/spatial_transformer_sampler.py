def spatial_transformer_sampler(*args, **kwargs):
    raise NotImplementedError("stand-in: chainer's own spatial_transformer_sampler is not restated here; "
                              "loans_b200.chainer_compat.install() rebinds this name")
