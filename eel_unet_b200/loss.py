"""Drop-in ``edge_BceDiceLoss`` (reference utils/Loss.py:92-113): one fused reduction kernel forward,
one gradient kernel backward, instead of ~100 tiny ATen launches."""
import torch.nn as nn

from . import ops
from ._lib import on_device


class edge_BceDiceLoss(nn.Module):  # noqa: N801  (name kept: train.py:305 constructs it by this name)
    def __init__(self, wb=1, wd=1):
        super().__init__()
        self.wb, self.wd = wb, wd

    @on_device
    def forward(self, gt_pre, out, target):
        gt_pre5, gt_pre4, gt_pre3, gt_pre2, gt_pre1 = gt_pre
        return ops.EdgeBceDice.apply(out, gt_pre5, gt_pre4, gt_pre3, gt_pre2, gt_pre1, target, self.wb, self.wd)
