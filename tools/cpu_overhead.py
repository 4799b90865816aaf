import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from eel_unet_b200 import EELUnet, edge_BceDiceLoss, synth
from eel_unet_b200.parallel import DataParallel, FusedAdam
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = EELUnet(3, 1, precision="bf16").to(dev).train()
dp = DataParallel(model); opt = FusedAdam(dp.buckets, lr=1e-4, weight_decay=1e-5); crit = edge_BceDiceLoss(1, 1)
xs, ys, _ = synth.batch(16, 256, 256, 0)
x = torch.from_numpy(xs).repeat(4, 1, 1, 1).to(dev); y = torch.from_numpy(ys).repeat(4, 1, 1, 1).to(dev)
def step():
    dp.zero_grad(); seg, e = dp(x); loss = crit(e, seg, y); loss.backward(); dp.finish_backward(); opt.step(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
# small batch: GPU is fast, so wall time ~ CPU enqueue time
xs2 = x[:2].contiguous(); ys2 = y[:2].contiguous()
def step2():
    dp.zero_grad(); seg, e = dp(xs2); loss = crit(e, seg, ys2); loss.backward(); dp.finish_backward(); opt.step(); return loss
for _ in range(3): step2()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step2()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("batch 2: enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
t0 = time.perf_counter()
for _ in range(10): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("batch 64: enqueue %.2f ms/step, total %.2f ms/step" % ((t1 - t0) * 100, (t2 - t0) * 100))
