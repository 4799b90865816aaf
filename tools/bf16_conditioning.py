#!/usr/bin/env python
"""How far is a bf16 TRAIN-mode forward from fp32, as a function of how conditioned the weights are?

    python tools/bf16_conditioning.py [--batch 8] [--size 256] [--lr 1e-4] [--steps 0,5,20,50,100,200]

At default initialisation the network amplifies rounding: a bf16 forward is ~24 % away from fp64, and so is
the reference's own torch.autocast(bfloat16).  The north_star bar for bf16 is 2e-2.  This tool trains the model with the fp32
path (parity-tested against the fp64 oracle) for a few Adam steps on synthetic batches and, at each checkpoint, reports the
relative L2 distance of a bf16 train-mode forward to the fp32 one on a held-out batch -- per output, next to stock autocast on
the same GPU.  tests/test_parity_configs_gpu.py pins the result as a test (fp64 oracle as ground truth).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--steps", default="0,5,20,50,100,200")
    ap.add_argument("--autocast", action="store_true", help="also run the oracle under torch.autocast(cuda, bf16) as a yardstick")
    args = ap.parse_args()
    from eel_unet_b200 import EELUnet, edge_BceDiceLoss, synth
    from eel_unet_b200.parallel import DataParallel, FusedAdam

    dev = torch.device("cuda", 0)
    marks = sorted(int(s) for s in args.steps.split(","))
    torch.manual_seed(0)
    model = EELUnet(3, 1, precision="fp32").to(dev).train()
    dp = DataParallel(model)
    opt = FusedAdam(dp.buckets, lr=args.lr, weight_decay=1e-5)
    crit = edge_BceDiceLoss(1, 1)
    pool = [synth.batch(args.batch, args.size, args.size, seed=100 + k)[:2] for k in range(4)]
    pool = [(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)) for x, y in pool]
    xs, ys, _ = synth.batch(args.batch, args.size, args.size, seed=7)
    xh = torch.from_numpy(xs).to(dev)

    def probe(step):
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        outs = {}
        for prec in ("fp32", "bf16"):
            m = EELUnet(3, 1, precision=prec).to(dev).train()
            m.load_state_dict(sd)
            with torch.no_grad():
                seg, edges = m(xh)
            outs[prec] = [seg] + list(edges)
        names = ["seg", "edge5", "edge4", "edge3", "edge2", "edge1"]
        row = {"step": step, "bf16_vs_fp32": {n: rel(a, b) for n, a, b in zip(names, outs["bf16"], outs["fp32"])}}
        if args.autocast:
            from oracle import eelunet_torch as O
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                seg, edges = O.forward(sd, xh, True, {})
            row["autocast_vs_fp32"] = {n: rel(a.float(), b) for n, a, b in zip(names, [seg] + list(edges), outs["fp32"])}
        print(json.dumps(row), flush=True)

    step = 0
    for mark in marks:
        while step < mark:
            x, y = pool[step % len(pool)]
            dp.zero_grad()
            seg, edges = dp(x)
            loss = crit(edges, seg, y)
            loss.backward()
            dp.finish_backward()
            opt.step()
            step += 1
        probe(step)


if __name__ == "__main__":
    main()
