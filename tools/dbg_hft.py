import sys, torch
sys.path.insert(0, '/root/repo')
from eel_unet_b200 import ops
sys.path.insert(0, '/root/repo/tests')
def ref_hft(x, mask_range=20):
    h, w = x.shape[-2:]
    crow, ccol = h // 2, w // 2
    r = min(mask_range, crow, ccol)
    mask = torch.ones(h, w, dtype=x.dtype, device=x.device)
    mask[crow - r:crow + r, ccol - r:ccol + r] = 0
    d = torch.fft.fftshift(torch.fft.fft2(x), dim=(-2, -1)) * mask
    return torch.abs(torch.fft.ifft2(torch.fft.ifftshift(d, dim=(-2, -1))))
for dtype in (torch.float32, torch.bfloat16):
    for shape in [(1,64,256,256),(2,64,256,256),(3,64,256,256),(4,64,256,256),(3,64,128,128),(3,128,128,128)]:
        torch.manual_seed(0)
        x = torch.randn(*shape, device='cuda')
        a = x.permute(0,2,3,1).contiguous().to(dtype)
        y = ops.HFT.apply(a, 20).float().permute(0,3,1,2)
        r = ref_hft(a.double().permute(0,3,1,2))
        per = [((y[i].double()-r[i]).norm()/r[i].norm()).item() for i in range(shape[0])]
        print(dtype, shape, ["%.3g"%e for e in per], flush=True)
