"""Synthetic inputs for the checkers: re-export of the generator that the bench and the tests share
(eel_unet_b200/synth.py -- pure numpy, no kernel, no reference code; it lives in the package so that the measured GPU
arm of bench.py imports nothing from oracle/)."""
from eel_unet_b200.synth import *  # noqa: F401,F403
from eel_unet_b200.synth import IMAGENET_MEAN, IMAGENET_STD, _blur5, _ellipse_mask  # noqa: F401
