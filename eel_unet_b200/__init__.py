"""eel_unet_b200 -- B200-native (sm_100a) implementation of the EEL-Unet training/inference hot path.

    from eel_unet_b200 import EELUnet, edge_BceDiceLoss      # drop-ins for the reference classes
    from eel_unet_b200 import edges                          # GPU Canny / Sobel edge maps
    from eel_unet_b200 import data                           # GPU Resize / ToTensor / Normalize of uint8 batches

Importing this package loads libeel.so; if it has not been built the import fails (no fallback).
"""
from . import _lib  # noqa: F401  (fails loudly when libeel.so is missing)
from . import data  # noqa: F401
from . import edges  # noqa: F401
from .loss import edge_BceDiceLoss  # noqa: F401
from .model import EELUnet  # noqa: F401
from .ops import weights_changed as invalidate_packed_weights  # noqa: F401  (call after updating weights through p.data)
from .unet import Unet  # noqa: F401

__all__ = ["EELUnet", "Unet", "edge_BceDiceLoss", "edges", "data", "invalidate_packed_weights"]
