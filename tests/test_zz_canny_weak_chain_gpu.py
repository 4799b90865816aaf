"""Hysteresis on a LONG chain of weak candidates (GPU): a one-pixel spiral whose gradient magnitude (4 x 30 = 120 on the sides, up to
6 x 30 = 180 at the corners) lies between the two thresholds everywhere except next to one bright run -- cv2.Canny reaches ~16 000
pixels of it (at 256 x 256) from that single strong seed, and nothing without it, and so
must each of the hysteresis implementations in csrc/edge.cu: the lock-free union-find (W = 257), the column-strip bitmap sweeps
(H <= 512) and the band-per-warp bitmap flood (H > 512).  (tests/test_edges_gpu.py::test_canny_long_snake_component draws its
spiral at 60, i.e. magnitude 240: strong everywhere, so it never needed the hysteresis to walk.)  The thread-level CPU transcription
of the same case is tests/test_canny_design_cpu.py::test_flood_follows_a_long_weak_chain_across_strips."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _weak_spiral(h, w, level=30):
    g = np.zeros((h, w), np.uint8)
    y0, x0, y1, x1 = 2, 2, h - 3, w - 3
    while y1 - y0 > 8 and x1 - x0 > 8:
        g[y0, x0:x1] = level; g[y0:y1, x1] = level; g[y1, x0 + 4:x1 + 1] = level; g[y0 + 4:y1 + 1, x0 + 4] = level
        y0 += 4; x0 += 4; y1 -= 4; x1 -= 4
    g[2, 2:6] = 255
    return g


@pytest.mark.parametrize("h,w", [(257, 257), (260, 260), (256, 256), (300, 512), (600, 512)])   # generation 1, 2, 3 (strips), 3 (strips), 3 (bands)
def test_hysteresis_walks_a_long_weak_chain(h, w):
    import cv2

    from eel_unet_b200 import edges
    from oracle import edge_np

    g = _weak_spiral(h, w)
    nms = edge_np.canny_nms(g, 100, 200)                         # 0 weak candidate, 1 none, 2 strong
    assert 0 < (nms == 2).sum() < 50 and (nms == 0).sum() > 10000  # one strong seed, a long weak chain
    ref = cv2.Canny(g, 100, 200)
    assert (ref != 0).sum() > 0.4 * (nms != 1).sum()              # cv2 follows the chain (one of the spiral's two edge lines) to its end
    out = edges.canny(torch.from_numpy(g[None]).cuda()).cpu().numpy()[0]
    assert np.array_equal(out, ref), "%d pixels differ" % int((out != ref).sum())
    # the same image without its seed: nothing is strong, nothing may survive
    g2 = g.copy(); g2[2, 2:6] = 30
    out2 = edges.canny(torch.from_numpy(g2[None]).cuda()).cpu().numpy()[0]
    assert not cv2.Canny(g2, 100, 200).any() and not out2.any()
