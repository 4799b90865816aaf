"""CPU: the plain-C oracle (oracle/edge_c.c) against cv2 and the numpy oracle, bit-exact."""
import numpy as np
import pytest


def test_c_oracle_matches_cv2_and_numpy():
    cv2 = pytest.importorskip("cv2")
    from oracle import edge_c, edge_np, synth

    rng = np.random.default_rng(9)
    for (n, h, w) in [(2, 128, 160), (1, 37, 53), (1, 1, 1), (1, 2, 9)]:
        imgs, _ = synth.tooth_images(n, h, w, seed=h)
        noise = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        for batch in (imgs, noise):
            c = edge_c.canny_rgb(batch)
            for i in range(n):
                g = cv2.cvtColor(batch[i], cv2.COLOR_RGB2GRAY).reshape(h, w)
                assert np.array_equal(c[i], cv2.Canny(g, 100, 200).reshape(h, w))
                assert np.array_equal(c[i], edge_np.canny(g))
