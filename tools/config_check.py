#!/usr/bin/env python
"""Timings of the BASELINE.json configurations that are not the bench line (printed as JSON lines):

  config 2  Canny on a 64 x 512 x 512 x 3 uint8 batch (bit-exact against the numpy oracle on a sample), device-resident and
            end to end from host memory, next to cv2 on the host cores when cv2 is importable
  config 5  EELUnet bf16 inference at 1024^2, batch 32;  Unet bf16 training step at 512^2, batch 16

    python tools/config_check.py [--quick]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from eel_unet_b200 import EELUnet, Unet, edges, synth
from oracle import edge_np   # checker for the bit-exactness line only

dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------------------------------------------------------- config 2
N, S = (8, 256) if quick else (64, 512)
base = synth.tooth_images(8, S, S, seed=3)[0]
imgs = np.concatenate([base] * (N // 8), 0)
d = torch.from_numpy(imgs).to(dev)
out = edges.canny(d)
ref = np.stack([edge_np.canny_rgb(imgs[i]) for i in range(2)])
exact = bool((out[:2].cpu().numpy() == ref).all())
ms = timed(lambda: edges.canny(d))
pinned = torch.from_numpy(imgs).pin_memory()
hostout = torch.empty((N, S, S), dtype=torch.uint8).pin_memory()


def e2e():
    dd = pinned.to(dev, non_blocking=True)
    hostout.copy_(edges.canny(dd), non_blocking=True)
    torch.cuda.synchronize()


ms_e2e = timed(e2e)
line = {"config": "canny %dx%dx%dx3 u8" % (N, S, S), "bit_exact_vs_oracle": exact, "device_ms": ms, "device_img_s": N / ms * 1e3,
        "device_GBs_algorithmic": N * S * S * 4 / ms / 1e6, "e2e_ms": ms_e2e, "e2e_img_s": N / ms_e2e * 1e3}
try:
    import cv2
    t0 = time.perf_counter()
    for i in range(N):
        cv2.Canny(cv2.cvtColor(imgs[i], cv2.COLOR_RGB2GRAY), 100, 200)
    t = time.perf_counter() - t0
    line["cv2_host_ms"] = t * 1e3
    line["cv2_threads"] = cv2.getNumThreads()
except Exception as ex:  # cv2 is a third-party oracle, not a dependency
    line["cv2_host_ms"] = None
print(json.dumps(line), flush=True)

# ---------------------------------------------------------------- config 5
B, S5 = (2, 256) if quick else (32, 1024)
torch.manual_seed(0)
m = EELUnet(3, 1, precision="bf16").to(dev).eval()
x = torch.randn(B, 3, S5, S5, device=dev)
with torch.no_grad():
    ms = timed(lambda: m(x), iters=3, warm=1)
print(json.dumps({"config": "EELUnet bf16 inference %dx3x%dx%d" % (B, S5, S5), "ms": ms, "img_s": B / ms * 1e3,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
del m, x
torch.cuda.empty_cache()

B, S5 = (2, 128) if quick else (16, 512)
u = Unet(3, 1, precision="bf16").to(dev).train()
x = torch.randn(B, 3, S5, S5, device=dev)
y = (torch.rand(B, 1, S5, S5, device=dev) > 0.5).float()


def ustep():
    for p in u.parameters():
        p.grad = None
    torch.nn.functional.binary_cross_entropy_with_logits(u(x), y).backward()


ms = timed(ustep, iters=3, warm=1)
print(json.dumps({"config": "Unet bf16 fwd+bwd %dx3x%dx%d" % (B, S5, S5), "ms": ms, "img_s": B / ms * 1e3}), flush=True)
