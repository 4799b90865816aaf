"""Reference-format ``state_dict`` of EELUnet / Unet from a flat layer table, with PyTorch's default initialisation
drawn in the reference's construction order (models/EELUnet.py:229-333, models/Unet.py:5-31) -- so that
``torch.manual_seed(s)`` gives the very weights the reference (and the drop-in) would have.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Exists so that the CPU reference arm of bench.py and the checkers
can build weights WITHOUT importing the product package (which maps libeel.so into the process).
Pinned by tests/test_oracle.py::test_seed0_weights_match_reference_checksums (per-tensor checksums the real reference
produced) and against the live reference when it is present.
"""
from collections import OrderedDict

import torch
import torch.nn as nn


def _put(sd, prefix, module):
    for k, v in module.state_dict().items():
        sd[prefix + "." + k] = v.detach().clone()


def _conv3(sd, p, ci, co):
    _put(sd, p, nn.Conv2d(ci, co, 3, padding=1))


def _bn(sd, p, c):
    _put(sd, p, nn.BatchNorm2d(c))


def _convt(sd, p, ci, co):
    _put(sd, p, nn.ConvTranspose2d(ci, co, 2, stride=2))


def _capmlp(sd, p, ci, co, token=64):
    """ChannelAwarePatchedMLP.__init__ (models/EELUnet.py:102-112): to_patch, ChannelAttention(fc1, fc2), mlp, to_space"""
    _put(sd, p + ".to_patch", nn.Conv2d(ci, token, 1))
    _put(sd, p + ".channel_attention.fc1", nn.Conv2d(token, token // 16, 1))
    _put(sd, p + ".channel_attention.fc2", nn.Conv2d(token // 16, token, 1))
    _put(sd, p + ".mlp.0", nn.Linear(token, token * 4))
    _put(sd, p + ".mlp.2", nn.Linear(token * 4, co))
    _put(sd, p + ".to_space", nn.Conv2d(co, co, 1))


def _conv_block(sd, p, ci, co):            # models/EELUnet.py:335-345
    _conv3(sd, p + ".0", ci, co); _bn(sd, p + ".1", co); _conv3(sd, p + ".3", co, co); _bn(sd, p + ".4", co)


def _mlp_conv_block(sd, p, ci, co):        # :347-359
    _conv3(sd, p + ".0", ci, co); _bn(sd, p + ".1", co); _capmlp(sd, p + ".3", co, co); _bn(sd, p + ".4", co)


def _upconv_block(sd, p, ci, co):          # :361-366
    _convt(sd, p + ".0", ci, co); _bn(sd, p + ".1", co)


def _mlp_upconv_block(sd, p, ci, co):      # :368-374
    _convt(sd, p + ".0", ci, co); _capmlp(sd, p + ".1", co, co); _bn(sd, p + ".2", co)


def eelunet_state_dict(in_channels=3, out_channels=1):
    """the 365 entries of EELUnet(in_channels, out_channels).state_dict(); consumes the global torch RNG like the reference"""
    sd = OrderedDict()
    _conv_block(sd, "enc1.0", in_channels, 64)
    _conv_block(sd, "enc2.0", 64, 128)
    _mlp_conv_block(sd, "enc3.0", 128, 256)
    _mlp_conv_block(sd, "enc4.0", 256, 512)
    _bn(sd, "bottleneck.0", 512); _conv3(sd, "bottleneck.1", 512, 1024); _capmlp(sd, "bottleneck.3", 1024, 1024)
    _mlp_upconv_block(sd, "upconv4", 1024, 512); _mlp_conv_block(sd, "dec4", 1024, 512)
    _mlp_upconv_block(sd, "upconv3", 512, 256); _mlp_conv_block(sd, "dec3", 512, 256)
    _upconv_block(sd, "upconv2", 256, 128); _conv_block(sd, "dec2", 256, 128)
    _upconv_block(sd, "upconv1", 128, 64); _conv_block(sd, "dec1", 128, 64)
    for k, c in ((5, 1024), (4, 512), (3, 256), (2, 128), (1, 64)):
        _put(sd, "pred%d.conv" % k, nn.Conv2d(c, 1, 1))
    _mlp_upconv_block(sd, "edge_upconv_4.0", 1024, 512); _mlp_conv_block(sd, "edge_upconv_4.1", 512, 512)
    _mlp_upconv_block(sd, "edge_upconv_3.0", 512, 256); _mlp_conv_block(sd, "edge_upconv_3.1", 256, 256)
    _upconv_block(sd, "edge_upconv_2.0", 256, 128); _conv_block(sd, "edge_upconv_2.2", 128, 128)
    _upconv_block(sd, "edge_upconv_1.0", 128, 64); _conv_block(sd, "edge_upconv_1.2", 64, 64)
    sd["final.0.weight"], sd["final.0.bias"] = torch.ones(64), torch.zeros(64)           # LayerNorm, :209-210
    _put(sd, "final.1", nn.Conv2d(64, out_channels, 1))
    return sd


def unet_state_dict(in_channels=3, out_channels=1):
    """models/Unet.py:5-31 (no BatchNorm)"""
    sd = OrderedDict()

    def blk(p, ci, co):
        _conv3(sd, p + ".0", ci, co); _conv3(sd, p + ".2", co, co)

    blk("enc1", in_channels, 64); blk("enc2", 64, 128); blk("enc3", 128, 256); blk("enc4", 256, 512)
    blk("bottleneck", 512, 1024)
    for k, c in ((4, 512), (3, 256), (2, 128), (1, 64)):
        _convt(sd, "upconv%d.0" % k, 2 * c, c)
        blk("conv%d" % k, 2 * c, c)
    _put(sd, "final_conv", nn.Conv2d(64, out_channels, 1))
    return sd
