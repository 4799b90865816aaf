"""Integer edge maps with OpenCV semantics on the GPU, batched (reference augmentation/AddCannyEdge.py,
augmentation/CannyEnhance.py, augmentation/Sobel.py, utils/tools.py:126-155).

Device API: uint8 CUDA tensors in, uint8 CUDA tensors out.  Host API (``*_host``): numpy arrays /
pinned buffers in and out, with the H2D and D2H copies done here -- this is the call the reference's
dataset code would make in place of cv2.
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, on_device, ptr, stream, workspace

U8 = torch.uint8


def _chk(t, last=None):
    if t.dtype != U8 or not t.is_cuda:
        raise _lib.EelError("edge maps take uint8 CUDA tensors")
    if last is not None and t.shape[-1] != last:
        raise _lib.EelError("expected a trailing dimension of %d" % last)
    return t if t.is_contiguous() else t.contiguous()


@on_device
def gray(rgb):
    """cv2.cvtColor(RGB2GRAY): [N,H,W,3] u8 -> [N,H,W] u8."""
    rgb = _chk(rgb, 3)
    N, H, W, _ = rgb.shape
    out = torch.empty((N, H, W), dtype=U8, device=rgb.device)
    call("eel_gray_u8", ptr(rgb), ptr(out), N, H, W, stream())
    return out


@on_device
def canny(img, low=100, high=200):
    """cv2.Canny(gray(img), low, high) for [N,H,W,3] RGB or cv2.Canny(img, low, high) for [N,H,W] gray."""
    rgb = img.dim() == 4
    img = _chk(img, 3 if rgb else None)
    N, H, W = img.shape[:3]
    out = torch.empty((N, H, W), dtype=U8, device=img.device)
    n = _lib.lib.eel_canny_workspace_bytes(N, H, W)
    ws = workspace(n, img.device, slot=2)
    call("eel_canny_rgb" if rgb else "eel_canny_gray", ptr(img), ptr(out), N, H, W, int(low), int(high), ptr(ws), n, stream())
    return out


@on_device
def sobel_map(gray_img):
    """augmentation/Sobel.py:9-14."""
    g = _chk(gray_img)
    N, H, W = g.shape
    out = torch.empty_like(g)
    call("eel_sobel_map", ptr(g), ptr(out), N, H, W, stream())
    return out


@on_device
def laplacian_map(gray_img):
    """augmentation/Sobel.py:17-18."""
    g = _chk(gray_img)
    N, H, W = g.shape
    out = torch.empty_like(g)
    call("eel_laplacian_map", ptr(g), ptr(out), N, H, W, stream())
    return out


@on_device
def canny_enhance(rgb, low=100, high=200, edge_color=(0, 0, 0), alpha=0.5):
    """CannyEnhance.__call__ (augmentation/CannyEnhance.py:21-44) on a batch."""
    rgb = _chk(rgb, 3)
    e = canny(rgb, low, high)
    N, H, W, _ = rgb.shape
    out = torch.empty_like(rgb)
    call("eel_canny_enhance", ptr(rgb), ptr(e), ptr(out), N, H, W, int(edge_color[0]), int(edge_color[1]),
         int(edge_color[2]), float(alpha), stream())
    return out


@on_device
def add_canny_edge(rgb, low=100, high=200):
    """AddCannyEdge.__call__ (augmentation/AddCannyEdge.py:15-41): RGB + edge map as a 4th channel (RGBA)."""
    rgb = _chk(rgb, 3)
    e = canny(rgb, low, high)
    return torch.cat((rgb, e.unsqueeze(-1)), dim=-1)


@on_device
def edge_label(gt):
    """generate_edge_label (utils/tools.py:126-155): gt float [N,1,H,W] -> {0,1} float edge labels."""
    g = (gt[:, 0] * 255).to(U8)
    return (canny(g) .to(torch.float32) / 255.0).unsqueeze(1)


def canny_host(rgb_np, low=100, high=200, device="cuda"):
    """numpy [N,H,W,3] (or [H,W,3]) uint8 -> numpy edge maps; copies included."""
    single = rgb_np.ndim == 3
    a = np.ascontiguousarray(rgb_np[None] if single else rgb_np)
    d = torch.from_numpy(a).to(device, non_blocking=True)
    out = canny(d, low, high).cpu().numpy()
    return out[0] if single else out
