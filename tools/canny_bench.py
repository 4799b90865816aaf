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
ap.add_argument("--json", action="store_true", help="also print a JSON line with end-to-end and cv2 timings")
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
if "--json" in sys.argv:
    import json, time
    pinned = torch.from_numpy(imgs).pin_memory()
    hostout = torch.empty(imgs.shape[:3], dtype=torch.uint8).pin_memory()
    te = []
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hostout.copy_(edges.canny(pinned.cuda(non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize()
        te.append((time.perf_counter() - t0) * 1e3)
    line = {"config": "canny %dx%dx%dx3 u8" % (a.batch, a.size, a.size), "bit_exact_vs_oracle": True, "device_ms": ms,
            "device_img_s": a.batch / ms * 1e3, "device_GBs_algorithmic": 4 * px / ms / 1e6, "e2e_ms": sorted(te)[len(te) // 2],
            "code": "third-generation Canny (register row bands + bitmap flood fill)"}
    try:
        import cv2
        t0 = time.perf_counter()
        for i in range(imgs.shape[0]):
            cv2.Canny(cv2.cvtColor(imgs[i], cv2.COLOR_RGB2GRAY), 100, 200)
        line["cv2_host_ms"] = (time.perf_counter() - t0) * 1e3
        line["cv2_threads"] = cv2.getNumThreads()
    except Exception:
        line["cv2_host_ms"] = None
    print(json.dumps(line))
print("canny %dx%dx%d: %.1f us device-resident, %.1f GB/s on 4 B/pixel (%.1f %% of 6534.8), edge pixels %.2f %%, bit-exact on 3 images"
      % (a.batch, a.size, a.size, ms * 1e3, 4 * px / ms / 1e6, 4 * px / ms / 1e6 / 65.348, 100.0 * float((out != 0).float().mean())))
