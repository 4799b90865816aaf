"""numpy/cv2 restatement of the reference's evaluate() arithmetic (evaluate.py:25-124).  TEST INFRASTRUCTURE ONLY."""
import numpy as np


def seg2bnd(mask, dilation_ratio=0.02):
    """evaluate.py:25-41."""
    import cv2

    h, w = mask.shape
    d = max(int(round(np.mean([h, w]) * dilation_ratio)), 1)
    m = (mask * 255).astype(np.uint8)
    er = cv2.erode(m, np.ones((3, 3), np.uint8), iterations=d)
    return (m - er) > 0


def boundary_f1(gt, pred):
    """evaluate.py:43-60."""
    gb, pb = seg2bnd(gt), seg2bnd(pred)
    tp = np.logical_and(pb, gb).sum()
    precision = tp / (pb.sum() + 1e-7)
    recall = tp / (gb.sum() + 1e-7)
    return 0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)


def evaluate_batches(batches):
    """batches: iterable of (seg_prob [B,1,H,W] float32 numpy, labels [B,1,H,W]); evaluate.py:70-124."""
    TP = TN = FP = FN = 0
    bf, cnt = 0.0, 0
    for seg, lab in batches:
        p = (seg > 0.5).astype(np.float32).reshape(-1)
        l = lab.reshape(-1)
        TP += int(((p == 1) & (l == 1)).sum()); TN += int(((p == 0) & (l == 0)).sum())
        FP += int(((p == 1) & (l == 0)).sum()); FN += int(((p == 0) & (l == 1)).sum())
        for i in range(seg.shape[0]):
            bf += boundary_f1(lab[i, 0], (seg[i, 0] > 0.5).astype(np.float32))
            cnt += 1
    e = 1e-7
    acc = (TP + TN) / (TP + TN + FP + FN + e)
    prec = TP / (TP + FP + e); rec = TP / (TP + FN + e)
    f1 = 2 * prec * rec / (prec + rec + e)
    iou = TP / (TP + FP + FN + e)
    dfg = 2 * TP / (2 * TP + FP + FN + e); dbg = 2 * TN / (2 * TN + FP + FN + e)
    ioub = TN / (TN + FP + FN + e)
    return acc, prec, rec, f1, iou, dfg, (iou + ioub) / 2, bf / (cnt + e), (dfg + dbg) / 2


def seeded_batches():
    """the seeded (probabilities, labels) batches shared by tests/golden/make_golden_r2.py (which runs the reference's own
    evaluate() on them) and the metric tests; includes an empty prediction (boundary precision 0/0 path)"""
    from . import synth

    rng = np.random.default_rng(0)
    batches = []
    for k, (n, h, w) in enumerate([(3, 256, 256), (2, 96, 160), (1, 48, 48)]):
        _, lab, _ = synth.batch(n, h, w, seed=10 + k)
        _, other, _ = synth.batch(n, h, w, seed=20 + k)
        seg = np.clip(0.7 * other + 0.25 * lab + rng.uniform(-0.2, 0.2, size=lab.shape), 0, 1).astype(np.float32)
        if k == 1:
            seg[0] = 0.0
        batches.append((seg, lab))
    return batches
