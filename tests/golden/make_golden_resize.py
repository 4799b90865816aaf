"""Generates tests/golden/resize_pil.npz with Pillow + torchvision themselves (the third-party code the reference's
dataset calls: data/ToothDataset.py:58-61, train.py:249-252).  Run in the build container:

    python tests/golden/make_golden_resize.py
"""
import os
import sys

import numpy as np
import PIL
import torch
from PIL import Image
from torchvision import transforms

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import synth

out = {"pillow_version": np.array(PIL.__version__)}
rng = np.random.default_rng(7)
cases = [("down2", (128, 160), (64, 80)), ("down_frac", (150, 201), (64, 96)), ("up", (40, 56), (96, 112)),
         ("square", (96, 96), (64, 64)), ("same_w", (90, 64), (64, 64)), ("tooth", (192, 256), (128, 128))]
for name, (h, w), (oh, ow) in cases:
    if name == "tooth":
        img = synth.tooth_images(1, h, w, seed=5)[0][0]
    else:
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mask = (rng.random((h, w)) > 0.5).astype(np.uint8) * 255
    out[name + "_img"] = img
    out[name + "_mask"] = mask
    out[name + "_size"] = np.array([oh, ow])
    out[name + "_img_resized"] = np.asarray(Image.fromarray(img, "RGB").resize((ow, oh), Image.BILINEAR))
    out[name + "_mask_resized"] = np.asarray(Image.fromarray(mask, "L").resize((ow, oh), Image.BILINEAR))
    tf = transforms.Compose([transforms.Resize((oh, ow)), transforms.ToTensor()])
    t = tf(Image.fromarray(img, "RGB"))
    t = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])(t)
    out[name + "_img_tensor"] = t.numpy()
    out[name + "_mask_tensor"] = tf(Image.fromarray(mask, "L")).numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "resize_pil.npz"), **out)
print("wrote resize_pil.npz with", len(out), "arrays")
