"""Operator front ends with the reference's names and signatures (torch tensors as the device arrays).

    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler

mirror ``functions.rotation_droput.rotation_dropout``, ``chainer.functions.spatial_transformer_grid`` and
``chainer.functions.spatial_transformer_sampler`` as used at reference sheep/sheep_localizer.py:61-63;
``stn_crop`` is the same three steps as one fused call; ``prepare_images`` mirrors ``SheepLocalizer.prepare_images``
(:72-82) on the device; ``ingest_frames`` is the loader's ``resize_image(frame, image_size) / 255``
(common/datasets/image_dataset.py:16-28, :98) for a batch of decoded uint8 frames on the device.
"""
from loans_b200.functions.rotation_droput import RotationDropout, rotation_dropout          # noqa: F401
from loans_b200.functions.spatial_transformer import (                                      # noqa: F401
    InvalidType, spatial_transformer_grid, spatial_transformer_sampler, stn_crop)
from loans_b200.functions.prepare import prepare_images                                     # noqa: F401
from loans_b200.functions.ingest import FrameIngest, ingest_frames                          # noqa: F401
