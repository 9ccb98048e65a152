"""multimodal_alzheimer_b200 — B200-native training hot path of Liz490/multimodal_alzheimer.

`pkg/` mirrors the reference's package layout for the path (models, loss_functions, utils); everything below
it runs on hand-written sm_100a CUDA kernels reached through the C-ABI in include/adni_b200.h.
"""
from ._lib import LIB_PATH, AdniError, launch_count  # noqa: F401

__version__ = "0.1.0"
