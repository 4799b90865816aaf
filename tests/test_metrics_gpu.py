"""GPU: evaluate() metrics against the numpy/cv2 restatement of evaluate.py -- exact (integer counts)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_metrics_match_reference_arithmetic():
    from eel_unet_b200.metrics import SegmentationMetrics
    from oracle import metrics_np, synth

    rng = np.random.default_rng(0)
    m = SegmentationMetrics("cuda")
    batches = []
    for k, (n, h, w) in enumerate([(3, 256, 256), (2, 96, 160), (1, 48, 48)]):
        _, lab, _ = synth.batch(n, h, w, seed=10 + k)
        _, other, _ = synth.batch(n, h, w, seed=20 + k)
        seg = np.clip(0.7 * other + 0.25 * lab + rng.uniform(-0.2, 0.2, size=lab.shape), 0, 1).astype(np.float32)
        if k == 1:
            seg[0] = 0.0          # empty prediction: boundary precision 0/0 path
        batches.append((seg, lab))
        m.update(torch.from_numpy(seg).cuda(), torch.from_numpy(lab).cuda())
    ours, ref = m.compute(), metrics_np.evaluate_batches(batches)
    for a, b in zip(ours, ref):
        assert abs(a - b) <= 1e-12, (ours, ref)


def test_evaluate_function_signature():
    from eel_unet_b200 import EELUnet
    from eel_unet_b200.metrics import evaluate
    from oracle import synth

    torch.manual_seed(0)
    model = EELUnet(3, 1).cuda()
    xs, ys, _ = synth.batch(2, 64, 64, 0)
    out = evaluate(model, [(torch.from_numpy(xs), torch.from_numpy(ys))], "cuda")
    assert len(out) == 9 and all(0.0 <= float(v) <= 1.0 for v in out)
