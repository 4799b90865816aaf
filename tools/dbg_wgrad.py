import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from eel_unet_b200 import EELUnet, edge_BceDiceLoss, ops
from eel_unet_b200 import synth

def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

xs, ys, _ = synth.batch(2, 128, 128, 0)
x, y = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
crit = edge_BceDiceLoss(1, 1)
for precision in ("fp32", "bf16"):
    torch.manual_seed(0)
    a = EELUnet(3, 1, precision=precision).cuda().train()
    b = EELUnet(3, 1, precision=precision).cuda().train()
    b.load_state_dict(a.state_dict())
    ops.set_wgrad_stream(False)
    seg, e = a(x); crit(e, seg, y).backward()
    torch.cuda.synchronize()
    for trial in range(3):
        ops.set_wgrad_stream(True)
        for p in b.parameters():
            p.grad = None
        seg, e = b(x); crit(e, seg, y).backward()
        torch.cuda.synchronize()
        bad = []
        gmax = max(p.grad.norm().item() for p in a.parameters())
        for (n, p), q in zip(a.named_parameters(), b.parameters()):
            if p.grad.norm().item() > 1e-4 * gmax and rel(q.grad, p.grad) > (1e-3 if precision == "fp32" else 0.5):
                bad.append((n, round(rel(q.grad, p.grad), 4), tuple(p.shape)))
        print(precision, "trial", trial, "bad:", bad[:12], len(bad))
