"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of the image pre-processing the reference's dataset performs
per sample on the host (reference data/ToothDataset.py:58-61 with train.py:249-252):

    transforms.Resize((S, S))  ->  PIL.Image.resize(..., BILINEAR)   (torchvision hands PIL images to Pillow)
    transforms.ToTensor()      ->  uint8 HWC / 255 as float32 CHW
    transforms.Normalize(ImageNet mean / std)                        (image only; the mask is Resize + ToTensor)

Pillow is a third-party dependency of the reference (through torchvision; no version pinned by the reference, 12.2.0 in
this image) and is not vendored under /root/reference, so its published algorithm is restated here: Pillow's
src/libImaging/Resample.c -- ``precompute_coeffs`` (filter support scaled by the down-scale factor = antialiasing, double
arithmetic), ``normalize_coeffs_8bpc`` (fixed point, PRECISION_BITS = 32 - 8 - 2) and the two 8-bit passes,
horizontal first, each rounding to uint8 (``clip8``).  Pinned bit-exactly against Pillow itself by
tests/golden/make_golden_resize.py -> tests/golden/resize_pil.npz (tests/test_oracle.py).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def bilinear_coeffs(in_size, out_size):
    """-> (ksize, bounds[out,2] int32 (xmin, count), kk[out,ksize] int32 fixed-point weights)  (Resample.c precompute_coeffs)"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, np.float64)
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
        ww = w[:xmax].sum() if xmax > 0 else 0.0
        # Pillow accumulates ww in loop order; a sequential python sum is the same order
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, out_size, axis):
    """one 8-bit resampling pass along `axis` of an [H, W, C] uint8 array"""
    in_size = img.shape[axis]
    _, bounds, kk = bilinear_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def resize_bilinear_u8(img, out_h, out_w):
    """PIL.Image.resize((out_w, out_h), BILINEAR) of an [H, W, C] (or [H, W]) uint8 image"""
    squeeze = img.ndim == 2
    a = img[:, :, None] if squeeze else img
    if a.shape[1] != out_w:
        a = _pass(a, out_w, 1)      # horizontal first (Resample.c ImagingResampleInner)
    if a.shape[0] != out_h:
        a = _pass(a, out_h, 0)
    return a[:, :, 0] if squeeze else a


def preprocess_image(img_u8, size, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """ToothDataset.__getitem__ image branch: [H, W, 3] uint8 -> float32 [3, S, S]"""
    r = resize_bilinear_u8(img_u8, size[0], size[1]).astype(np.float32) / np.float32(255.0)
    r = (r - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return np.ascontiguousarray(r.transpose(2, 0, 1))


def preprocess_mask(mask_u8, size):
    """mask branch: [H, W] uint8 -> float32 [1, S, S] in [0, 1]"""
    r = resize_bilinear_u8(mask_u8, size[0], size[1]).astype(np.float32) / np.float32(255.0)
    return r[None]
