"""Bit-exact parity (GPU) of the integer edge-map kernels with OpenCV (the library the reference calls)
and with the numpy oracle, through the public eel_unet_b200.edges API (-> C ABI)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cases():
    from oracle import synth

    rng = np.random.default_rng(11)
    out = []
    for (n, h, w) in [(3, 256, 256), (2, 37, 53), (1, 1, 1), (1, 2, 7), (2, 128, 300), (1, 16, 64), (1, 17, 65), (1, 600, 40),
                      # W % 32 == 0 and W <= 512: register-band stage 1 + bitmap hysteresis (third generation)
                      (2, 100, 96), (1, 33, 32), (1, 5, 512), (1, 1, 32), (1, 600, 480), (1, 1700, 512), (5, 64, 512)]:
        imgs, _ = synth.tooth_images(n, h, w, seed=h * 7 + w)
        out.append(imgs)
        out.append(rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8))   # dense-edge stress
    return out


def test_gray_canny_sobel_laplacian_enhance_match_cv2():
    import cv2

    from eel_unet_b200 import edges
    from oracle import edge_np

    for imgs in _cases():
        d = torch.from_numpy(imgs).cuda()
        g = edges.gray(d)
        e = edges.canny(d)
        eg = edges.canny(g)
        sm = edges.sobel_map(g)
        lp = edges.laplacian_map(g)
        en = edges.canny_enhance(d, edge_color=(255, 255, 255), alpha=0.2)
        en2 = edges.canny_enhance(d, edge_color=(13, 200, 77), alpha=0.5)
        rgba = edges.add_canny_edge(d)
        for i, img in enumerate(imgs):
            cg = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY).reshape(img.shape[:2])
            ce = cv2.Canny(cg, 100, 200).reshape(img.shape[:2])
            assert np.array_equal(g[i].cpu().numpy(), cg)
            assert np.array_equal(e[i].cpu().numpy(), ce), "canny mismatch %s: %d px" % (img.shape, (e[i].cpu().numpy() != ce).sum())
            assert np.array_equal(eg[i].cpu().numpy(), ce)
            assert np.array_equal(e[i].cpu().numpy(), edge_np.canny(cg))
            assert np.array_equal(sm[i].cpu().numpy(), edge_np.sobel_map(cg))
            assert np.array_equal(lp[i].cpu().numpy(), edge_np.laplacian_map(cg))
            ov = np.zeros_like(img)
            ov[ce != 0] = (255, 255, 255)
            assert np.array_equal(en[i].cpu().numpy(), cv2.addWeighted(img, 1.0, ov, 0.2, 0).reshape(img.shape))
            assert np.array_equal(en2[i].cpu().numpy(), edge_np.canny_enhance(img, ce, (13, 200, 77), 0.5))
            assert np.array_equal(rgba[i, ..., 3].cpu().numpy(), ce) and np.array_equal(rgba[i, ..., :3].cpu().numpy(), img)


def test_canny_other_thresholds_and_edge_label():
    import cv2

    from eel_unet_b200 import edges
    from oracle import synth

    imgs, masks = synth.tooth_images(2, 160, 224, seed=5)
    d = torch.from_numpy(imgs).cuda()
    for lo, hi in [(5, 200), (50, 60), (0, 0), (300, 900)]:
        e = edges.canny(d, lo, hi).cpu().numpy()
        for i, img in enumerate(imgs):
            assert np.array_equal(e[i], cv2.Canny(cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), lo, hi))
    lab = edges.edge_label(torch.from_numpy(masks).cuda()).cpu().numpy()
    for i in range(2):
        ref = cv2.Canny((masks[i, 0] * 255).astype(np.uint8), 100, 200).astype(np.float32) / 255.0
        assert np.array_equal(lab[i, 0], ref)


def test_canny_config2_full_size_golden_and_idempotent_output():
    """BASELINE config 2: 64 x 512 x 512 x 3 uint8, bit-exact vs cv2; plus size-independent properties."""
    import cv2

    from eel_unet_b200 import edges
    from oracle import synth

    imgs, _ = synth.tooth_images(64, 512, 512, seed=0)
    d = torch.from_numpy(imgs).cuda()
    e = edges.canny(d)
    e2 = edges.canny(d)
    assert torch.equal(e, e2)                       # deterministic despite the lock-free union-find
    en = e.cpu().numpy()
    assert set(np.unique(en)) <= {0, 255}
    for i in range(64):
        assert np.array_equal(en[i], cv2.Canny(cv2.cvtColor(imgs[i], cv2.COLOR_RGB2GRAY), 100, 200))
    # host-buffer entry point (the call the reference's dataset code would make)
    assert np.array_equal(edges.canny_host(imgs[:4]), en[:4])


def test_canny_long_snake_component():
    """A one-pixel-wide weak spiral reachable from a single strong pixel: hysteresis must follow it all."""
    import cv2

    from eel_unet_b200 import edges

    for h, w in ((257, 257), (256, 256), (300, 512)):     # 257: tile kernels + union-find; the others: bitmap flood fill
        _snake(h, w)


def _snake(h, w):
    import cv2

    from eel_unet_b200 import edges

    g = np.zeros((h, w), np.uint8)
    y0, x0, y1, x1 = 2, 2, h - 3, w - 3
    while y1 - y0 > 8 and x1 - x0 > 8:
        g[y0, x0:x1] = 60; g[y0:y1, x1] = 60; g[y1, x0 + 4:x1 + 1] = 60; g[y0 + 4:y1 + 1, x0 + 4] = 60
        y0 += 4; x0 += 4; y1 -= 4; x1 -= 4
    g[2, 2:6] = 255
    ref = cv2.Canny(g, 100, 200)
    out = edges.canny(torch.from_numpy(g[None]).cuda()).cpu().numpy()[0]
    assert np.array_equal(out, ref)
