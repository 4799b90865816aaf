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


class DevicePrefetcher:
    """Host -> device hand-over of training batches that overlaps the copy of batch i+1 with the compute of batch i (what
    ``DataLoader(pin_memory=True)`` + ``.to(device, non_blocking=True)`` of the reference's loop, train.py:62-66, cannot do on
    one stream: there the 67 MB of a 64 x 3 x 256^2 fp32 batch cross PCIe in front of every step).

        for x, y in DevicePrefetcher(loader, device):       # loader yields tuples of (pinned) host tensors
            loss = criterion(*model(x)[::-1], y); ...

    Two device buffers per tensor, filled alternately on a copy stream.  No allocator traffic and no ``record_stream``: a buffer
    is refilled only after the compute stream has passed the point where the batch that used it was handed back (``get`` of the
    batch after next), and the compute stream waits for the copy's event before it reads.  ``put`` / ``get`` are the two halves
    for callers that drive it by hand (bench.py)."""

    def __init__(self, loader=None, device=None):
        self.loader = loader
        self.device = torch.device(device if device is not None else "cuda")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]          # per slot: list of device tensors
        self.ready = [None, None]         # event: the slot's copy finished
        self.free = [None, None]          # event on the compute stream: the slot's previous batch was handed back two gets ago
        self.n_put = self.n_get = 0

    def put(self, *host_tensors):
        """start copying one batch (tuple of host tensors, pinned for a truly asynchronous copy) into the next slot"""
        if self.n_put - self.n_get >= 2:
            raise _lib.EelError("DevicePrefetcher: both slots are in flight (call get() first)")
        k = self.n_put % 2
        if self.bufs[k] is None or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(self.bufs[k], host_tensors)) \
                or len(self.bufs[k]) != len(host_tensors):
            self.bufs[k] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host_tensors]
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])
            else:
                self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))      # (fresh buffers: allocation order)
            for b, h in zip(self.bufs[k], host_tensors):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[k] = ev
        self.n_put += 1

    def get(self):
        """the oldest batch in flight as device tensors (valid until the get() after next)"""
        if self.n_get >= self.n_put:
            raise _lib.EelError("DevicePrefetcher: nothing in flight (call put() first)")
        k = self.n_get % 2
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[k])
        # the OTHER slot's batch was handed out by the previous get(): everything enqueued so far has consumed it or will
        # have by the time this event fires, so the copy stream may refill it afterwards
        ev = torch.cuda.Event()
        ev.record(cur)
        self.free[1 - k] = ev
        self.n_get += 1
        return tuple(self.bufs[k])

    def __iter__(self):
        it = iter(self.loader)
        try:
            self.put(*next(it))
        except StopIteration:
            return
        for nxt in it:
            batch = self.get()
            self.put(*nxt)
            yield batch
        yield self.get()
