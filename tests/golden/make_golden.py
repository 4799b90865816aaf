#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference (build container only; /root/reference must exist).

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are outputs of
the reference itself on seeded synthetic inputs:
  eelunet_train_2x128.npz  one fp64 train step of models/EELUnet.py + utils/Loss.py (ground truth) and the
                           reference's own fp32 deviations from it (its rounding-noise floor)
  eelunet_eval_2x128.npz   eval-mode forward (running statistics)
  loss_cases.npz           utils/Loss.py edge_BceDiceLoss on random probabilities (value + gradients)
  edges_*.npz              cv2 (the library the reference calls: AddCannyEdge.py:25-27, CannyEnhance.py:32-43,
                           Sobel.py:9-18, tools.py:145) on synthetic images of several sizes
Weights are not stored (105 MB): they are regenerated from torch.manual_seed(0) with the reference's
construction order; per-tensor checksums pin them.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import, synth  # noqa: E402


def model_fixtures():
    EELUnet, EdgeLoss, _ = ref_import.load()
    torch.manual_seed(0)
    m32 = EELUnet(3, 1)
    keys = list(m32.state_dict().keys())
    chk = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in m32.state_dict().values()])
    xs, ys, _ = synth.batch(2, 128, 128, 0)
    x, y = torch.from_numpy(xs), torch.from_numpy(ys)

    m64 = copy.deepcopy(m32).double().train()
    seg64, e64 = m64(x.double())
    l64 = EdgeLoss(1, 1)(e64, seg64, y.double())
    l64.backward()
    m32t = copy.deepcopy(m32).train()
    seg32, e32 = m32t(x)
    l32 = EdgeLoss(1, 1)(e32, seg32, y)
    l32.backward()

    out = {
        "keys": np.array(keys), "state_checksums": chk,
        "seg": seg64.detach().numpy().astype(np.float32), "loss": np.float64(l64.item()),
        "seg_f32_relerr": np.float64(((seg32.double() - seg64).norm() / seg64.norm()).item()),
        "loss_f32_abserr": np.float64(abs(l32.item() - l64.item())),
    }
    for k, (a, b) in enumerate(zip(e64, e32)):
        out["edge%d" % (5 - k)] = a.detach().numpy().astype(np.float32)
        out["edge%d_f32_relerr" % (5 - k)] = np.float64(((b.double() - a).norm() / a.norm()).item())
    names, gnorm, gsum, gerr32 = [], [], [], []
    small = {}
    for (n, p), (_, q) in zip(m64.named_parameters(), m32t.named_parameters()):
        names.append(n)
        gnorm.append(p.grad.norm().item())
        gsum.append(p.grad.sum().item())
        gerr32.append(((q.grad.double() - p.grad).norm() / (p.grad.norm() + 1e-300)).item())
        if p.numel() <= 1024:
            small["grad:" + n] = p.grad.numpy().astype(np.float64)
    out.update(grad_names=np.array(names), grad_norm=np.array(gnorm), grad_sum=np.array(gsum), grad_f32_relerr=np.array(gerr32))
    out.update(small)
    for k, v in m64.state_dict().items():
        if "running_" in k:
            out["stat:" + k] = v.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "eelunet_train_2x128.npz"), **out)

    me = copy.deepcopy(m32).double().eval()
    with torch.no_grad():
        sege, ee = me(x.double())
    ev = {"seg": sege.numpy().astype(np.float32)}
    for k, a in enumerate(ee):
        ev["edge%d" % (5 - k)] = a.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "eelunet_eval_2x128.npz"), **ev)

    # loss alone
    g = torch.Generator().manual_seed(7)
    n, h, w = 3, 64, 96
    t = (torch.rand(n, 1, h, w, generator=g) > 0.7).double()
    t[1] = torch.nn.functional.avg_pool2d(t[1:2], 3, 1, 1)[0]          # one non-binary sample
    preds = [torch.rand(n, 1, h // s, w // s, generator=g, dtype=torch.float64).clamp(1e-4, 1 - 1e-4).requires_grad_(True)
             for s in (1, 16, 8, 4, 2, 1)]
    loss = EdgeLoss(1, 1)(preds[1:], preds[0], t)
    loss.backward()
    lc = {"target": t.numpy(), "loss": np.float64(loss.item())}
    for i, p in enumerate(preds):
        lc["pred%d" % i] = p.detach().numpy()
        lc["grad%d" % i] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "loss_cases.npz"), **lc)


def edge_fixtures():
    import cv2

    cv2.setNumThreads(1)
    for (n, h, w) in [(2, 96, 128), (2, 37, 53), (1, 256, 256)]:
        imgs, masks = synth.tooth_images(n, h, w, seed=h * 7 + w)
        gray = np.stack([cv2.cvtColor(i, cv2.COLOR_RGB2GRAY) for i in imgs])
        canny = np.stack([cv2.Canny(g, 100, 200) for g in gray])
        sob = np.stack([cv2.convertScaleAbs(cv2.magnitude(cv2.Sobel(g, cv2.CV_64F, 1, 0, ksize=3), cv2.Sobel(g, cv2.CV_64F, 0, 1, ksize=3)))
                        for g in gray]).reshape(gray.shape)
        lap = np.stack([cv2.convertScaleAbs(cv2.Laplacian(g, cv2.CV_64F)) for g in gray]).reshape(gray.shape)
        enh = []
        for i, e in zip(imgs, canny):
            ov = np.zeros_like(i)
            ov[e != 0] = (255, 255, 255)
            enh.append(cv2.addWeighted(i, 1.0, ov, 0.2, 0))
        label = np.stack([cv2.Canny((m[0] * 255).astype(np.uint8), 100, 200) for m in masks])
        np.savez_compressed(os.path.join(HERE, "edges_%dx%dx%d.npz" % (n, h, w)), seed=h * 7 + w, gray=gray, canny=canny, sobel=sob,
                            laplacian=lap, enhance=np.stack(enh), label=label, cv2_version=cv2.__version__)


if __name__ == "__main__":
    model_fixtures()
    edge_fixtures()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
