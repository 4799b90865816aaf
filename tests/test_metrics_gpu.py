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


def test_metrics_match_reference_evaluate_golden():
    """the GPU metrics against the 9-tuple the reference's OWN evaluate() returned on the same batches
    (tests/golden/metrics_eval.npz, made by tests/golden/make_golden_r2.py from evaluate.py:62-124), and the per-sample
    boundary counts against its seg2bnd()."""
    import os

    from eel_unet_b200.metrics import SegmentationMetrics
    from oracle import metrics_np

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_eval.npz"))
    m = SegmentationMetrics("cuda")
    batches = metrics_np.seeded_batches()
    for seg, lab in batches:
        m.update(torch.from_numpy(seg).cuda(), torch.from_numpy(lab).cuda())
    got = np.array(m.compute())
    assert np.abs(got - g["metrics"]).max() <= 1e-12, (got, g["metrics"])
    ps = torch.cat(m._per_sample).cpu().numpy()
    k = 0
    for seg, lab in batches:
        for i in range(seg.shape[0]):
            h, w = seg.shape[-2:]
            pb = np.unpackbits(g["bnd_pred_%d" % k])[:h * w].astype(bool)
            gb = np.unpackbits(g["bnd_gt_%d" % k])[:h * w].astype(bool)
            assert tuple(ps[k]) == (int((pb & gb).sum()), int(pb.sum()), int(gb.sum())), k
            k += 1
