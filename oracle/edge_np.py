"""numpy restatement of the integer edge-map pipeline the reference gets from OpenCV.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The algorithm lives in a third-party dependency that is not under /root/reference: OpenCV
(``cv2``; the reference pins no version, the build image has 4.13.0).  The call sites that define
the path are
  * augmentation/AddCannyEdge.py:25-27   cvtColor(RGB2GRAY) + Canny(gray, 100, 200)
  * augmentation/CannyEnhance.py:32-43   same + white/colour overlay with addWeighted
  * utils/tools.py:143-145               Canny on (mask*255).astype(uint8)
  * augmentation/Sobel.py:9-18           Sobel(CV_64F, k=3) -> magnitude -> convertScaleAbs; Laplacian
What follows restates OpenCV's published algorithm for those calls (imgproc color/canny/deriv):
fixed-point luma, 3x3 Sobel with replicated border, L1 magnitude, tangent-table non-maximum
suppression, and hysteresis by 8-connected flood fill from the strong pixels.  It is pinned
bit-exactly against cv2 itself in tests/test_oracle_edges.py and tests/golden/edges_*.npz.
"""
import numpy as np

TG22 = 13573  # round(tan(22.5 deg) * 2**15)


def gray_u8(rgb):
    """cv2.cvtColor(RGB2GRAY) for uint8: 15-bit fixed point, coefficients sum to 2**15."""
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    return ((9798 * r + 19235 * g + 3735 * b + 16384) >> 15).astype(np.uint8)


def _sobel_i32(gray, border):
    """3x3 Sobel dx, dy as int32.  border: 'edge' (BORDER_REPLICATE) or 'reflect' (BORDER_REFLECT_101)."""
    p = np.pad(gray.astype(np.int32), ((0, 0),) * (gray.ndim - 2) + ((1, 1), (1, 1)), mode=border)
    h, w = gray.shape[-2:]

    def s(dy, dx):
        return p[..., dy:dy + h, dx:dx + w]

    dx = (s(0, 2) + 2 * s(1, 2) + s(2, 2)) - (s(0, 0) + 2 * s(1, 0) + s(2, 0))
    dy = (s(2, 0) + 2 * s(2, 1) + s(2, 2)) - (s(0, 0) + 2 * s(0, 1) + s(0, 2))
    return dx, dy


def canny_nms(gray, low=100, high=200):
    """Stage 1 of cv2.Canny(gray, low, high) with aperture 3 and the L1 norm.

    Returns an int8 map: 2 = strong edge pixel, 0 = weak candidate, 1 = not an edge
    (OpenCV's own encoding of its intermediate map).
    """
    dx, dy = _sobel_i32(gray, "edge")
    mag = np.abs(dx) + np.abs(dy)
    h, w = gray.shape[-2:]
    mp = np.pad(mag, ((0, 0),) * (gray.ndim - 2) + ((1, 1), (1, 1)), mode="constant")

    def m(oy, ox):
        return mp[..., 1 + oy:1 + oy + h, 1 + ox:1 + ox + w]

    ax = np.abs(dx).astype(np.int64)
    ay = np.abs(dy).astype(np.int64) << 15
    t22 = ax * TG22
    t67 = t22 + (ax << 16)
    c = mag
    horiz = ay < t22
    vert = ay > t67
    neg = (dx ^ dy) < 0
    keep_h = (c > m(0, -1)) & (c >= m(0, 1))
    keep_v = (c > m(-1, 0)) & (c >= m(1, 0))
    keep_d_same = (c > m(-1, -1)) & (c > m(1, 1))   # dx, dy same sign: up-left / down-right
    keep_d_diff = (c > m(-1, 1)) & (c > m(1, -1))   # opposite signs: up-right / down-left
    keep = np.where(horiz, keep_h, np.where(vert, keep_v, np.where(neg, keep_d_diff, keep_d_same)))
    cand = (c > low) & keep
    out = np.ones(gray.shape, dtype=np.int8)
    out[cand] = 0
    out[cand & (c > high)] = 2
    return out


def hysteresis(nms_map):
    """Stage 2: every weak candidate 8-connected (through candidates) to a strong pixel becomes an edge."""
    if nms_map.ndim > 2:
        return np.stack([hysteresis(x) for x in nms_map])
    h, w = nms_map.shape
    m = np.ones((h + 2, w + 2), dtype=np.int8)
    m[1:-1, 1:-1] = nms_map
    stack = list(zip(*np.nonzero(m == 2)))
    while stack:
        y, x = stack.pop()
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if m[y + dy, x + dx] == 0:
                    m[y + dy, x + dx] = 2
                    stack.append((y + dy, x + dx))
    return np.where(m[1:-1, 1:-1] == 2, 255, 0).astype(np.uint8)


def canny(gray, low=100, high=200):
    """cv2.Canny(gray, low, high) (AddCannyEdge.py:27, CannyEnhance.py:35, tools.py:145)."""
    return hysteresis(canny_nms(gray, low, high))


def canny_rgb(rgb, low=100, high=200):
    """AddCannyEdge.py:25-27: gray then Canny."""
    return canny(gray_u8(rgb), low, high)


def sobel_map(gray):
    """Sobel.py:9-14: Sobel(CV_64F,k3) x/y (reflect-101 border) -> magnitude -> convertScaleAbs."""
    dx, dy = _sobel_i32(gray, "reflect")
    mag = np.sqrt((dx.astype(np.float64) ** 2 + dy.astype(np.float64) ** 2))
    return np.minimum(np.rint(mag), 255).astype(np.uint8)


def laplacian_map(gray):
    """Sobel.py:17-18: Laplacian(CV_64F) (aperture 1: 4-neighbour kernel, reflect-101) -> convertScaleAbs."""
    p = np.pad(gray.astype(np.int32), ((0, 0),) * (gray.ndim - 2) + ((1, 1), (1, 1)), mode="reflect")
    h, w = gray.shape[-2:]
    lap = p[..., 0:h, 1:w + 1] + p[..., 2:h + 2, 1:w + 1] + p[..., 1:h + 1, 0:w] + p[..., 1:h + 1, 2:w + 2] \
        - 4 * p[..., 1:h + 1, 1:w + 1]
    return np.minimum(np.abs(lap), 255).astype(np.uint8)


def canny_enhance(rgb, edges, color=(255, 255, 255), alpha=0.2):
    """CannyEnhance.py:38-43: overlay[edges != 0] = color; addWeighted(img, 1, overlay, alpha, 0).

    OpenCV evaluates addWeighted for uint8 in float32 and rounds half-to-even, then saturates.
    """
    ov = np.zeros(rgb.shape, dtype=np.float32)
    ov[edges != 0] = np.asarray(color, dtype=np.float32)
    t = rgb.astype(np.float32) * np.float32(1.0) + ov * np.float32(alpha) + np.float32(0.0)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def edge_label(mask01):
    """utils/tools.py:143-148: Canny(100,200) of (gt*255).astype(uint8), scaled to {0,1} float32."""
    g = (mask01 * 255).astype(np.uint8)
    return canny(g).astype(np.float32) / np.float32(255.0)
