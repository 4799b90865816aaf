"""GPU ``evaluate()`` (reference evaluate.py:62-124): same 9-tuple, computed from integer counts accumulated on the
device -- one host synchronisation at the end instead of four ``.item()`` calls and a cv2 round trip per sample."""
import torch

from . import _lib
from ._lib import call, on_device, ptr, stream, workspace


class SegmentationMetrics:
    def __init__(self, device):
        self.device = torch.device(device)
        self.conf = torch.zeros(4, dtype=torch.int64, device=self.device)   # TP, TN, FP, FN
        self._per_sample = []

    @on_device
    def update(self, seg_prob, labels):
        """seg_prob: [B,1,H,W] probabilities (eelunet) -- evaluate.py:91 thresholds at 0.5; labels: [B,1,H,W]."""
        seg = seg_prob.detach().to(torch.float32).contiguous()
        lab = labels.detach().to(torch.float32).contiguous()
        B, _, H, W = seg.shape
        call("eel_confusion_counts", ptr(seg), ptr(lab), seg.numel(), ptr(self.conf), stream())
        d = max(int(round((H + W) / 2 * 0.02)), 1)                     # evaluate.py:34-35
        ps = torch.zeros((B, 3), dtype=torch.int64, device=self.device)
        n = _lib.lib.eel_boundary_workspace_bytes(B, H, W)
        ws = workspace(n, self.device, slot=3)
        call("eel_boundary_counts", ptr(seg), ptr(lab), B, H, W, d, ptr(ps), ptr(ws), n, stream())
        self._per_sample.append(ps)

    def compute(self):
        """(pixel_accuracy, precision, recall, f1, iou, dice_fg, miou, avg_boundary_f1, mdice) -- evaluate.py:112-124."""
        TP, TN, FP, FN = (int(v) for v in self.conf.cpu())
        bf1_total, count = 0.0, 0
        if self._per_sample:
            for tp, pb, gb in torch.cat(self._per_sample).cpu().tolist():
                precision = tp / (pb + 1e-7)
                recall = tp / (gb + 1e-7)
                bf1_total += 0 if precision + recall == 0 else 2 * precision * recall / (precision + recall)
                count += 1
        eps = 1e-7
        pixel_accuracy = (TP + TN) / (TP + TN + FP + FN + eps)
        precision = TP / (TP + FP + eps)
        recall = TP / (TP + FN + eps)
        f1 = 2 * precision * recall / (precision + recall + eps)
        iou = TP / (TP + FP + FN + eps)
        dice_fg = 2 * TP / (2 * TP + FP + FN + eps)
        dice_bg = 2 * TN / (2 * TN + FP + FN + eps)
        mdice = (dice_fg + dice_bg) / 2
        iou_bg = TN / (TN + FP + FN + eps)
        miou = (iou + iou_bg) / 2
        return pixel_accuracy, precision, recall, f1, iou, dice_fg, miou, bf1_total / (count + eps), mdice


def evaluate(model, dataloader, device):
    """Drop-in for the reference's evaluate(model, dataloader, device) (evaluate.py:62)."""
    model.eval()
    m = SegmentationMetrics(device)
    with torch.no_grad():
        for inputs, labels in dataloader:
            inputs, labels = inputs.to(device), labels.to(device)
            outputs = model(inputs)
            name = getattr(model, "name", "")
            seg = outputs[0] if name == "eelunet" else (outputs[1] if name == "egeunet" else outputs)
            m.update(seg, labels)
    return m.compute()
