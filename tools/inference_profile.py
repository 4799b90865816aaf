import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from eel_unet_b200 import EELUnet, _lib, profiling
dev = torch.device("cuda", 0)
m = EELUnet(3, 1, precision="bf16").to(dev).eval()
x = torch.randn(32, 3, 1024, 1024, device=dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    rec = []
    _lib.set_profiler(rec)
    m(x)
    _lib.set_profiler(None)
    torch.cuda.synchronize()
fam = profiling.summarize(rec)
for r in profiling.table(fam, 6534.8, 1407.2)[:14]:
    print("%-22s %4d %8.3f ms %5.1f%% %8.1f TF/s %8.1f GB/s" % (r[0], r[1], r[2], r[3], r[4], r[6]))
