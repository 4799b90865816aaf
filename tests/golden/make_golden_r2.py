#!/usr/bin/env python
"""Round-2 fixtures from the REAL reference (build container only; /root/reference must exist).

    python tests/golden/make_golden_r2.py

  metrics_eval.npz         the reference's own ``evaluate(model, dataloader, device)`` (evaluate.py:62-124) on seeded
                           probability / label batches fed through a stand-in model, plus ``seg2bnd`` (:25-41) and
                           ``boundary_f1_score`` (:43-60) per sample -- pins oracle/metrics_np.py and the GPU metrics
  unet_2x64x96.npz         models/Unet.py fp64 forward + BCE-with-logits backward -- pins oracle.unet_forward and the drop-in
  eelunet_train_8x256.npz  BASELINE config 1's shape (8 x 3 x 256 x 256): one fp64 train step of models/EELUnet.py +
                           utils/Loss.py, and the reference's own fp32 deviations from it
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


def metrics_fixture():
    ev = ref_import.load_evaluate()

    class Passthrough:
        """stand-in for the network: evaluate() only needs .name, .eval() and a call returning (seg_prob, edges)"""
        name = "eelunet"

        def eval(self):
            return self

        def __call__(self, x):
            return x, []

    from oracle import metrics_np

    batches = metrics_np.seeded_batches()
    loader = [(torch.from_numpy(s), torch.from_numpy(l)) for s, l in batches]
    out = {"metrics": np.array(ev.evaluate(Passthrough(), loader, torch.device("cpu")), dtype=np.float64)}
    bf, nb = [], 0
    for k, (s, l) in enumerate(batches):
        for i in range(s.shape[0]):
            pred = (s[i, 0] > 0.5).astype(np.float32)
            bf.append(ev.boundary_f1_score(l[i, 0], pred))
            out["bnd_pred_%d" % nb] = np.packbits(ev.seg2bnd(pred))
            out["bnd_gt_%d" % nb] = np.packbits(ev.seg2bnd(l[i, 0]))
            nb += 1
    out["boundary_f1"] = np.array(bf, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "metrics_eval.npz"), **out)


def unet_fixture():
    _, _, Unet = ref_import.load()
    torch.manual_seed(0)
    m = Unet(3, 1).double()
    xs, ys, _ = synth.batch(2, 64, 96, 1)
    x, y = torch.from_numpy(xs).double(), torch.from_numpy(ys).double()
    out = m(x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    fx = {"logits": out.detach().numpy(), "loss": np.float64(loss.item()),
          "grad_names": np.array([n for n, _ in m.named_parameters()]),
          "grad_norm": np.array([p.grad.norm().item() for p in m.parameters()]),
          "grad_sum": np.array([p.grad.sum().item() for p in m.parameters()])}
    for n, p in m.named_parameters():
        if p.numel() <= 2048:
            fx["grad:" + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "unet_2x64x96.npz"), **fx)


def eelunet_256_fixture():
    EELUnet, EdgeLoss, _ = ref_import.load()
    torch.manual_seed(0)
    m32 = EELUnet(3, 1)
    xs, ys, _ = synth.batch(8, 256, 256, 0)
    x, y = torch.from_numpy(xs), torch.from_numpy(ys)
    m64 = copy.deepcopy(m32).double().train()
    seg64, e64 = m64(x.double())
    l64 = EdgeLoss(1, 1)(e64, seg64, y.double())
    l64.backward()
    m32t = copy.deepcopy(m32).train()
    seg32, e32 = m32t(x)
    l32 = EdgeLoss(1, 1)(e32, seg32, y)
    l32.backward()
    out = {"seg": seg64.detach().numpy().astype(np.float16), "loss": np.float64(l64.item()),
           "seg_sum": np.float64(seg64.sum().item()), "seg_sqsum": np.float64(seg64.pow(2).sum().item()),
           "seg_f32_relerr": np.float64(((seg32.double() - seg64).norm() / seg64.norm()).item()),
           "loss_f32_abserr": np.float64(abs(l32.item() - l64.item()))}
    for k, (a, b) in enumerate(zip(e64, e32)):
        if a.numel() <= 8 * 64 * 64:
            out["edge%d" % (5 - k)] = a.detach().numpy().astype(np.float32)
        out["edge%d_sum" % (5 - k)] = np.float64(a.sum().item())
        out["edge%d_f32_relerr" % (5 - k)] = np.float64(((b.double() - a).norm() / a.norm()).item())
    names, gnorm, gsum, gerr32 = [], [], [], []
    for (n, p), (_, q) in zip(m64.named_parameters(), m32t.named_parameters()):
        names.append(n)
        gnorm.append(p.grad.norm().item())
        gsum.append(p.grad.sum().item())
        gerr32.append(((q.grad.double() - p.grad).norm() / (p.grad.norm() + 1e-300)).item())
        if p.numel() <= 1024:
            out["grad:" + n] = p.grad.numpy().astype(np.float64)
    out.update(grad_names=np.array(names), grad_norm=np.array(gnorm), grad_sum=np.array(gsum), grad_f32_relerr=np.array(gerr32))
    for k, v in m64.state_dict().items():
        if "running_" in k:
            out["stat:" + k] = v.numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, "eelunet_train_8x256.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["metrics", "unet", "eelunet256"]
    if "metrics" in which:
        metrics_fixture()
    if "unet" in which:
        unet_fixture()
    if "eelunet256" in which:
        eelunet_256_fixture()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
