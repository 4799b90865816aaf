"""Drop-in ``Unet`` for the reference's ``models/Unet.py`` (BASELINE config 5 comparator): same constructor, same
``.name == "unet"`` (train.py:67), same state_dict keys, logits out.  It reuses the conv3x3 / ConvTranspose /
max-pool kernels of the EELUnet path; conv + ReLU is one fused kernel (no BatchNorm in this model)."""
import torch
import torch.nn as nn

from . import ops
from ._lib import EelError
from .model import _PRECISIONS


def _block(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True),
                         nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True))


def _up(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2))


class Unet(nn.Module):
    def __init__(self, in_channels, out_channels, precision="fp32"):
        super().__init__()
        self.name = "unet"
        self.enc1, self.enc2, self.enc3, self.enc4 = _block(in_channels, 64), _block(64, 128), _block(128, 256), _block(256, 512)
        self.bottleneck = _block(512, 1024)
        self.upconv4, self.conv4 = _up(1024, 512), _block(1024, 512)
        self.upconv3, self.conv3 = _up(512, 256), _block(512, 256)
        self.upconv2, self.conv2 = _up(256, 128), _block(256, 128)
        self.upconv1, self.conv1 = _up(128, 64), _block(128, 64)
        self.final_conv = nn.Conv2d(64, out_channels, kernel_size=1)
        self.set_precision(precision)

    def set_precision(self, precision):
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        self.compute_dtype = _PRECISIONS[precision]
        self._packer = None
        return self

    @staticmethod
    def _b(blk, x):
        x = ops.Conv3x3.apply(x, blk[0].weight, blk[0].bias, True)
        return ops.Conv3x3.apply(x, blk[2].weight, blk[2].bias, True)

    def forward(self, x):
        if not x.is_cuda:
            raise EelError("eel_unet_b200.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[2] % 16 or x.shape[3] % 16:
            raise EelError("input must be N x C x H x W with H and W multiples of 16 (got %s)" % (tuple(x.shape),))
        with torch.cuda.device(x.device):
            return self._forward(x)

    def _forward(self, x):
        if self.compute_dtype == torch.bfloat16:
            if self._packer is None or self._packer.stale():
                self._packer = ops.build_packer(self)
            self._packer.refresh(x.device)
        a = ops.nchw_to_nhwc(x, self.compute_dtype)
        e1 = self._b(self.enc1, a)
        e2 = self._b(self.enc2, ops.MaxPool2.apply(e1))
        e3 = self._b(self.enc3, ops.MaxPool2.apply(e2))
        e4 = self._b(self.enc4, ops.MaxPool2.apply(e3))
        d = self._b(self.bottleneck, ops.MaxPool2.apply(e4))
        for up, blk, skip in ((self.upconv4, self.conv4, e4), (self.upconv3, self.conv3, e3),
                              (self.upconv2, self.conv2, e2), (self.upconv1, self.conv1, e1)):
            d = ops.ConvT2x2.apply(d, up[0].weight, up[0].bias)
            d = self._b(blk, ops.Concat.apply(d, skip))      # models/Unet.py:78: concat((dec, enc), dim=1)
        out = ops.Linear.apply(d, self.final_conv.weight, self.final_conv.bias, False)
        return ops.ToNCHW.apply(out)
