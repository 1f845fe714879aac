"""loans_b200 -- the STN crop stage of LoANs (Bartzi/loans), rebuilt B200-native.

One hot path: ``rotation_dropout -> spatial_transformer_grid -> spatial_transformer_sampler``, forward and
backward (reference sheep/sheep_localizer.py:61-63), as hand-written sm_100a CUDA kernels behind the C ABI
in ``include/loans_stn.h`` (``libloans_stn.so``).  See DESIGN.md.

    from loans_b200.functions import rotation_dropout, spatial_transformer_grid, spatial_transformer_sampler, stn_crop
"""
from loans_b200.configuration import config, using_config          # noqa: F401

__version__ = "0.1.0"
