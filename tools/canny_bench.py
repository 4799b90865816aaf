#!/usr/bin/env python
"""BASELINE config 2: gray + Canny(100, 200) on a 64 x 512 x 512 x 3 uint8 batch (device resident and from host memory),
bit-exact against the numpy oracle on a sample; `--once` launches it exactly once (for ncu)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from eel_unet_b200 import edges, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--once", action="store_true")
a = ap.parse_args()
base, _ = synth.tooth_images(8, a.size, a.size, seed=3)
imgs = np.concatenate([base] * (a.batch // 8), 0)
d = torch.from_numpy(imgs).cuda()
out = edges.canny(d)
torch.cuda.synchronize()
if a.once:
    sys.exit(0)
from oracle import edge_np  # noqa: E402  (checker only)
ref = edge_np.canny_rgb(imgs[:3])
assert np.array_equal(out[:3].cpu().numpy(), ref), "canny mismatch"
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = edges.canny(d)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
px = imgs.shape[0] * a.size * a.size
print("canny %dx%dx%d: %.1f us device-resident, %.1f GB/s on 4 B/pixel (%.1f %% of 6534.8), edge pixels %.2f %%, bit-exact on 3 images"
      % (a.batch, a.size, a.size, ms * 1e3, 4 * px / ms / 1e6, 4 * px / ms / 1e6 / 65.348, 100.0 * float((out != 0).float().mean())))
