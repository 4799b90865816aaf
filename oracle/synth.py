"""Synthetic inputs for the checkers: the generator that the bench and the tests share (eel_unet_b200/synth.py -- pure
numpy, no kernel, no reference code; it lives in the package so that the measured GPU arm of bench.py imports nothing
from oracle/).  Loaded BY FILE PATH, not through the package: importing ``eel_unet_b200`` maps libeel.so, and the
CPU reference arm of bench.py must not have the product's library in its process."""
import importlib.util as _ilu
import os as _os

_path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "eel_unet_b200", "synth.py")
_spec = _ilu.spec_from_file_location("_eel_synth_standalone", _path)
_mod = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_mod)

IMAGENET_MEAN, IMAGENET_STD = _mod.IMAGENET_MEAN, _mod.IMAGENET_STD
_blur5, _ellipse_mask = _mod._blur5, _mod._ellipse_mask
tooth_images, normalize, soften, batch = _mod.tooth_images, _mod.normalize, _mod.soften, _mod.batch
