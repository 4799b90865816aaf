"""Dev probe: run HFT fwd+bwd at the two model shapes (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eel_unet_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for (c, s) in [(64, 256), (128, 128)]:
    x = torch.randn(n, s, s, c, device="cuda").bfloat16().requires_grad_(True)
    for _ in range(2):
        y = ops.HFT.apply(x, 20)
        y.backward(torch.ones_like(y))
    torch.cuda.synchronize()
print("ok")
