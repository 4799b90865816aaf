"""Deterministic synthetic inputs of the shape the reference trains on.

Input generator shared by bench.py, the tools and the tests.  Pure numpy (no cv2, no kernel) so the same bytes come out everywhere.

* images : uint8 N x H x W x 3 "tooth-like": 8-12 filled ellipses on a dark background, 5x5 binomial
           blur, uniform noise 0..11 (what data/ToothDataset.py:44-49 would hand to the transforms).
* masks  : float32 N x 1 x H x W in {0,1}: union of 1-3 ellipses (data/ToothDataset.py:61 -> ToTensor).
* model input = image/255 then ImageNet normalisation (data/ToothDataset.py:60).
"""
import numpy as np

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def _ellipse_mask(h, w, cy, cx, ry, rx, theta):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    c, s = np.cos(theta), np.sin(theta)
    u = (xx - cx) * c + (yy - cy) * s
    v = -(xx - cx) * s + (yy - cy) * c
    return (u / rx) ** 2 + (v / ry) ** 2 <= 1.0


def _blur5(img):
    """Separable 1-4-6-4-1 binomial blur, integer arithmetic with rounding, edge replicate."""
    k = np.array([1, 4, 6, 4, 1], dtype=np.int32)
    x = img.astype(np.int32)
    for axis in (0, 1):
        pad = [(0, 0)] * x.ndim
        pad[axis] = (2, 2)
        xp = np.pad(x, pad, mode="edge")
        acc = np.zeros_like(x)
        for i in range(5):
            sl = [slice(None)] * x.ndim
            sl[axis] = slice(i, i + x.shape[axis])
            acc += k[i] * xp[tuple(sl)]
        x = (acc + 8) >> 4
    return x.astype(np.uint8)


def tooth_images(n, h, w, seed=0):
    """uint8 [n, h, w, 3] images and float32 [n, 1, h, w] masks."""
    rng = np.random.default_rng(seed)
    imgs = np.empty((n, h, w, 3), dtype=np.uint8)
    masks = np.zeros((n, 1, h, w), dtype=np.float32)
    scale = min(h, w) / 512.0
    for i in range(n):
        img = np.empty((h, w, 3), dtype=np.uint8)
        img[:] = rng.integers(10, 50, size=3, dtype=np.uint8)
        ne = int(rng.integers(8, 13))
        nm = int(rng.integers(1, 4))
        for j in range(ne):
            ry = rng.uniform(10, 90) * scale + 2
            rx = rng.uniform(10, 90) * scale + 2
            cy = rng.uniform(0.1, 0.9) * h
            cx = rng.uniform(0.1, 0.9) * w
            th = rng.uniform(0, np.pi)
            m = _ellipse_mask(h, w, cy, cx, ry, rx, th)
            img[m] = rng.integers(60, 256, size=3, dtype=np.uint8)
            if j >= ne - nm:
                masks[i, 0][m] = 1.0
        img = _blur5(img)
        noise = rng.integers(0, 12, size=img.shape, dtype=np.int32)
        imgs[i] = np.clip(img.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return imgs, masks


def normalize(imgs_u8):
    """uint8 NHWC -> float32 NCHW, ToTensor + Normalize(ImageNet) (data/ToothDataset.py:58-60)."""
    x = imgs_u8.astype(np.float32) / np.float32(255.0)
    x = (x - IMAGENET_MEAN) / IMAGENET_STD
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def soften(masks, seed=0):
    """Non-binary targets (bilinear-resized masks are not exactly {0,1}): 3x3 box filter."""
    m = np.pad(masks, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    acc = np.zeros_like(masks)
    for dy in range(3):
        for dx in range(3):
            acc += m[:, :, dy:dy + masks.shape[2], dx:dx + masks.shape[3]]
    return (acc / 9.0).astype(np.float32)


def batch(n, h, w, seed=0):
    """(x float32 NCHW normalised, target float32 N1HW, raw uint8 NHWC)."""
    imgs, masks = tooth_images(n, h, w, seed)
    return normalize(imgs), masks, imgs
