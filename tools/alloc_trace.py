"""cudaMalloc calls of the caching allocator per training step (the timed region of bench.py must contain none)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from eel_unet_b200 import EELUnet, edge_BceDiceLoss, synth
from eel_unet_b200.parallel import DataParallel, FusedAdam

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = EELUnet(3, 1, precision="bf16").to(dev).train()
dp = DataParallel(model)
opt = FusedAdam(dp.buckets, lr=1e-4, weight_decay=1e-5)
crit = edge_BceDiceLoss(1, 1)
xs, ys, _ = synth.batch(16, 256, 256, seed=0)
x = torch.from_numpy(xs).repeat(4, 1, 1, 1).to(dev)
y = torch.from_numpy(ys).repeat(4, 1, 1, 1).to(dev)
prev = 0
for it in range(12):
    dp.zero_grad()
    seg, edges = dp(x)
    loss = crit(edges, seg, y)
    loss.backward()
    dp.finish_backward()
    opt.step()
    if it == 2 or it == 7:
        torch.cuda.synchronize()
    st = torch.cuda.memory_stats(dev)
    n = st["num_device_alloc"]
    print("step", it, "device allocs", n - prev, "reserved GB", round(st["reserved_bytes.all.current"] / 2 ** 30, 2),
          "largest new?", round(st["reserved_bytes.all.peak"] / 2 ** 30, 2))
    prev = n
