"""GPU input pipeline: the per-sample host work of the reference's dataset (reference data/ToothDataset.py:58-61 with the
transform of train.py:249-252) on whole uint8 batches that already sit in device memory.

    images = preprocess_images(batch_u8_nhwc, (256, 256))    # Resize -> ToTensor -> Normalize(ImageNet): fp32 NCHW
    masks  = preprocess_masks(mask_u8_nhw, (256, 256))       # Resize -> ToTensor: fp32 [N, 1, H, W] in [0, 1]

The resize is Pillow's antialiased BILINEAR, bit-exact (what torchvision's Resize does to the PIL images the reference
feeds it).  ``edges.add_canny_edge`` / ``edges.canny_enhance`` produce the optional augmented uint8 batches
(ToothDataset.py:51-55) that go through the same call (4-channel input: pass ``mean``/``std`` of length 4 or None).
"""
import torch

from . import _lib
from ._lib import call, on_device, ptr, stream, workspace

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_consts = {}


def _const(vals, device):
    key = (tuple(float(v) for v in vals), device)
    t = _consts.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.float32, device=device)
        _consts[key] = t
    return t


@on_device
def _run(x, size, mean, std, want_float, want_u8):
    if x.dtype != torch.uint8 or not x.is_cuda:
        raise _lib.EelError("the input pipeline takes uint8 CUDA tensors (NHWC)")
    x = x if x.is_contiguous() else x.contiguous()
    N, Hs, Ws, C = x.shape
    H, W = int(size[0]), int(size[1])
    if (mean is None) != (std is None) or (mean is not None and (len(mean) != C or len(std) != C)):
        raise _lib.EelError("mean / std must both be given with one value per channel (%d)" % C)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device) if want_float else None
    u8 = torch.empty((N, H, W, C), dtype=torch.uint8, device=x.device) if want_u8 else None
    n = _lib.lib.eel_preprocess_workspace_bytes(Hs, Ws, H, W)
    ws = workspace(n, x.device, slot=3)
    m = _const(mean, x.device) if mean is not None else None
    s = _const(std, x.device) if std is not None else None
    call("eel_preprocess_u8", ptr(x), N, Hs, Ws, C, H, W, ptr(m), ptr(s), ptr(out), ptr(u8), ptr(ws), n, stream())
    return out, u8


def preprocess_images(images_u8, size, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """[N, Hs, Ws, C] uint8 -> fp32 [N, C, H, W]: Resize(size) + ToTensor + Normalize (ToothDataset.py:58-60)."""
    return _run(images_u8, size, mean, std, True, False)[0]


def preprocess_masks(masks_u8, size):
    """[N, Hs, Ws] uint8 -> fp32 [N, 1, H, W] in [0, 1]: Resize(size) + ToTensor (ToothDataset.py:61)."""
    return _run(masks_u8.unsqueeze(-1), size, None, None, True, False)[0]


def resize_u8(images_u8, size):
    """[N, Hs, Ws, C] uint8 -> [N, H, W, C] uint8: PIL.Image.resize((W, H), BILINEAR) per image, bit-exact."""
    x = images_u8.unsqueeze(-1) if images_u8.dim() == 3 else images_u8
    r = _run(x, size, None, None, False, True)[1]
    return r.squeeze(-1) if images_u8.dim() == 3 else r
