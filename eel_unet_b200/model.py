"""Drop-in ``EELUnet`` for the reference's ``models/EELUnet.py``.

Same constructor ``EELUnet(in_channels, out_channels)``, same ``.name == "eelunet"`` dispatch key
(reference train.py:63, evaluate.py:84, test.py:109), same 365 ``state_dict`` keys / shapes / default
initialisation order (so a reference checkpoint loads with ``strict=True`` and the same seed gives the
same weights), same return value ``(seg_prob, [edge_5, ..., edge_1])`` (reference
models/EELUnet.py:471).  The module tree below only *holds parameters*; ``forward`` never calls the
sub-modules -- it drives the sm_100a kernels of libeel.so through ``ops`` on NHWC activations.

There is no CPU path: a CPU tensor raises.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import EelError


# ----------------------------------------------------------------------------------------------
# parameter containers (mirror the reference's module tree; reference models/EELUnet.py:8-225)
class ChannelAttention(nn.Module):
    def __init__(self, in_channels, reduction=16):
        super().__init__()
        self.in_channels, self.reduction = in_channels, reduction
        self.global_avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(in_channels, in_channels // reduction, kernel_size=1, bias=True)
        self.fc2 = nn.Conv2d(in_channels // reduction, in_channels, kernel_size=1, bias=True)
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()


class ShiftedChannel(nn.Module):
    def __init__(self, shift_ratio=0.25):
        super().__init__()
        self.shift_ratio = shift_ratio


class ChannelAwarePatchedMLP(nn.Module):
    def __init__(self, in_channels, out_channels, token_dim=64):
        super().__init__()
        self.shift = ShiftedChannel()
        self.to_patch = nn.Conv2d(in_channels, token_dim, kernel_size=1)
        self.channel_attention = ChannelAttention(token_dim)
        self.mlp = nn.Sequential(nn.Linear(token_dim, token_dim * 4), nn.GELU(), nn.Linear(token_dim * 4, out_channels))
        self.to_space = nn.Conv2d(out_channels, out_channels, kernel_size=1)


class FeatureInterleaveBridge(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.channels = channels


class HighFourierTransform(nn.Module):
    def __init__(self, mask_range=20):
        super().__init__()
        self.mask_range = mask_range


class PredictionGuidedRefinement(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.conv = nn.Conv2d(in_channels, 1, kernel_size=1, stride=1)


class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError
        self.normalized_shape = (normalized_shape,)


def _conv_block(cin, cout):
    return nn.Sequential(
        nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.Conv2d(cout, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _mlp_conv_block(cin, cout):
    return nn.Sequential(
        nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        ChannelAwarePatchedMLP(cout, cout), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _upconv_block(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2), nn.BatchNorm2d(cout))


def _mlp_upconv_block(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2), ChannelAwarePatchedMLP(cout, cout),
                         nn.BatchNorm2d(cout))


_PRECISIONS = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}


class _Bridge:
    """a skip bridge whose tensors have not been combined yet: (upconv BatchNorm, its input z, edge feature b, encoder skip e)"""
    __slots__ = ("bn", "z", "b", "e")

    def __init__(self, bn, z, b, e):
        self.bn, self.z, self.b, self.e = bn, z, b, e


class EELUnet(nn.Module):
    """B200-native EEL-UNet (reference models/EELUnet.py:228-471)."""

    # token MLP of ChannelAwarePatchedMLP as one kernel (False / EEL_FUSED_MLP=0: the three separate launches; kept for A/B runs)
    fused_mlp = __import__("os").environ.get("EEL_FUSED_MLP", "1") != "0"
    fused_shift = __import__("os").environ.get("EEL_FUSED_SHIFT", "1") != "0"
    fused_bridge = __import__("os").environ.get("EEL_FUSED_BRIDGE", "1") != "0"

    def __init__(self, in_channels, out_channels, precision="fp32"):
        super().__init__()
        self.name = "eelunet"
        # construction order = the reference's, so that default init consumes the RNG identically
        self.enc1 = nn.Sequential(_conv_block(in_channels, 64))
        self.enc2 = nn.Sequential(_conv_block(64, 128))
        self.enc3 = nn.Sequential(_mlp_conv_block(128, 256))
        self.enc4 = nn.Sequential(_mlp_conv_block(256, 512))
        self.bottleneck = nn.Sequential(
            nn.BatchNorm2d(512), nn.Conv2d(512, 1024, kernel_size=3, padding=1), nn.ReLU(inplace=True),
            ChannelAwarePatchedMLP(1024, 1024), nn.ReLU(inplace=True))
        self.upconv4 = _mlp_upconv_block(1024, 512)
        self.dec4 = _mlp_conv_block(1024, 512)
        self.upconv3 = _mlp_upconv_block(512, 256)
        self.dec3 = _mlp_conv_block(512, 256)
        self.upconv2 = _upconv_block(256, 128)
        self.dec2 = _conv_block(256, 128)
        self.upconv1 = _upconv_block(128, 64)
        self.dec1 = _conv_block(128, 64)
        self.pred5 = PredictionGuidedRefinement(1024)
        self.pred4 = PredictionGuidedRefinement(512)
        self.pred3 = PredictionGuidedRefinement(256)
        self.pred2 = PredictionGuidedRefinement(128)
        self.pred1 = PredictionGuidedRefinement(64)
        self.channel_interleave_bridge4 = FeatureInterleaveBridge(1024)
        self.channel_interleave_bridge3 = FeatureInterleaveBridge(512)
        self.channel_interleave_bridge2 = FeatureInterleaveBridge(256)
        self.channel_interleave_bridge1 = FeatureInterleaveBridge(128)
        self.edge_upconv_4 = nn.Sequential(_mlp_upconv_block(1024, 512), _mlp_conv_block(512, 512))
        self.edge_upconv_3 = nn.Sequential(_mlp_upconv_block(512, 256), _mlp_conv_block(256, 256))
        self.edge_upconv_2 = nn.Sequential(_upconv_block(256, 128), HighFourierTransform(), _conv_block(128, 128))
        self.edge_upconv_1 = nn.Sequential(_upconv_block(128, 64), HighFourierTransform(), _conv_block(64, 64))
        self.final = nn.Sequential(LayerNorm(normalized_shape=64, data_format="channels_first"), nn.Conv2d(64, out_channels, 1))
        self.set_precision(precision)

    # ------------------------------------------------------------------------------------------
    def set_precision(self, precision):
        """'fp32': fp32 storage + exact FFMA GEMMs; 'bf16': bf16 activations, fp32 accumulation/statistics."""
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_PRECISIONS))
        self.compute_dtype = _PRECISIONS[precision]
        self._packer = None
        self._fpacker = None
        self._cpacker = None
        return self

    def _weight_packer(self):
        if self._packer is None or self._packer.stale():
            self._packer = ops.build_packer(self)
            if EELUnet.fused_bridge:
                for blk in (self.dec4, self.dec3, self.dec2, self.dec1):      # the convs that read a skip bridge (ops.BridgeConv3x3)
                    self._packer.want_split(blk[0].weight)
        return self._packer

    def _composed_packer(self):
        """(mlp[2], to_space) pairs of the ChannelAwarePatchedMLP blocks, composed into one matrix per step"""
        if self._cpacker is None or self._cpacker.dtype != self.compute_dtype or self._cpacker.stale():
            self._cpacker = ops.build_composed(self, self.compute_dtype)
        return self._cpacker

    def _folded_packer(self):
        """inference: (producer, BatchNorm) pairs whose BatchNorm is folded into the producer's packed weight"""
        if self._fpacker is None or self._fpacker.stale():
            fp = ops.FoldedPacker()

            def conv(c, bn):
                co, ci = c.weight.shape[0], c.weight.shape[1]
                if ci % 64 == 0 and co % 64 == 0:
                    fp.add(c.weight, c.bias, bn, c.weight.shape, (2, 3, 0, 1), 0, c)          # [ky][kx][co][ci], scale over co

            def convt(c, bn):
                ci, co = c.weight.shape[0], c.weight.shape[1]
                if ci % 64 == 0 and co % 64 == 0:
                    fp.add(c.weight, c.bias, bn, c.weight.shape, (2, 3, 1, 0), 1, c)          # [ky][kx][co][ci]

            def lin(m, bn):
                # to_space(mlp[2](.)) composed into one matrix, then the BatchNorm folded into that (ops.ComposedPacker)
                no, k = m.to_space.weight.shape[0], m.mlp[2].weight.shape[1]
                if no % 64 == 0 and k % 64 == 0:
                    fp.composed.add(m.mlp[2], m.to_space, bn)

            if EELUnet.fused_bridge:
                for blk in (self.dec4, self.dec3, self.dec2, self.dec1):      # the convs that read a skip bridge
                    fp.want_split(blk[0].weight)
            for blk in (self.enc1[0], self.enc2[0], self.dec2, self.dec1, self.edge_upconv_2[2], self.edge_upconv_1[2]):
                conv(blk[0], blk[1]); conv(blk[3], blk[4])
            for blk in (self.enc3[0], self.enc4[0], self.dec4, self.dec3, self.edge_upconv_4[1], self.edge_upconv_3[1]):
                conv(blk[0], blk[1]); lin(blk[3], blk[4])
            for blk in (self.upconv2, self.upconv1, self.edge_upconv_2[0], self.edge_upconv_1[0]):
                convt(blk[0], blk[1])
            for blk in (self.upconv4, self.upconv3, self.edge_upconv_4[0], self.edge_upconv_3[0]):
                lin(blk[1], blk[2])
            self._fpacker = fp
        return self._fpacker

    # ---- fused stages ------------------------------------------------------------------------
    @staticmethod
    def _bn_mode(bn):
        """training flag of a BatchNorm + the bookkeeping nn.BatchNorm2d.forward does (num_batches_tracked)"""
        training = bn.training or bn.running_mean is None
        if training and bn.track_running_stats:
            if bn.momentum is None:
                raise EelError("BatchNorm momentum=None (cumulative average) is not used by the reference and not supported")
            bn.num_batches_tracked += 1
        return training

    @staticmethod
    def _bn_add_interleave(bn, z, b, e):
        """BatchNorm(z) + b, interleaved with e (decoder skip bridge) without materialising BatchNorm(z)"""
        training = EELUnet._bn_mode(bn)
        return ops.BNAddInterleave.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                         bn.momentum if bn.momentum is not None else 0.1, bn.eps, b, e, True)

    @staticmethod
    def _bn(bn, z, relu, producer_bias=True, single_conv_consumer=False, shift_out=False):
        training = EELUnet._bn_mode(bn)
        # producer_bias: z comes straight from a biased conv / linear, whose bias gradient (= column sums of dz) the
        # BatchNorm backward then delivers for free.  single_conv_consumer: the result feeds exactly one conv3x3, whose
        # data-gradient launch then also delivers this BatchNorm's backward sums (ops.Conv3x3.backward)
        return ops.BNAct.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, relu,
                               bn.momentum if bn.momentum is not None else 0.1, bn.eps, producer_bias, single_conv_consumer, shift_out)

    @staticmethod
    def _capmlp(m, x, bn=None, relu=False, defer=False, single_conv_consumer=False, pre_shifted=False):
        """ChannelAwarePatchedMLP (reference models/EELUnet.py:114-123); the channel shift is folded into to_patch.
        bn: the BatchNorm (and `relu`) that consume the result -- applied here: in training its statistics come out of
        to_space's epilogue, in inference it is folded into to_space's weights."""
        t = ops.Linear.apply(x, m.to_patch.weight, m.to_patch.bias, "pre" if pre_shifted else True)
        ca = m.channel_attention
        t = ops.SE.apply(t, ca.fc1.weight, ca.fc1.bias, ca.fc2.weight, ca.fc2.bias, True)
        pair = (m.mlp[2].weight, m.mlp[2].bias, m.to_space.weight, m.to_space.bias)
        fused = EELUnet.fused_mlp and ops.mlp_chain_supported(t, m.mlp[0].weight, m.to_space.weight.shape[0])
        f = ops.folded(m.to_space.weight) if bn is not None else None
        if f is not None and fused:      # inference: the whole token MLP + folded BatchNorm (+ ReLU) in one kernel
            return ops.mlp_chain_folded(t, m.mlp[0].weight, m.mlp[0].bias, f[0], f[1], relu)
        if not fused:
            t = ops.Linear.apply(t, m.mlp[0].weight, m.mlp[0].bias, False)
            t = ops.Gelu.apply(t, True)
            if f is not None:
                return ops.linear_folded(t, f[0], f[1], relu)

        def tail(u):
            if fused:                    # mlp[0] -> GELU -> composed (mlp[2], to_space) as ONE kernel (ops.MlpChain, csrc/capmlp_tc.cu)
                return ops.MlpChain.apply(u, m.mlp[0].weight, m.mlp[0].bias, *pair)
            # mlp[2] and to_space are two linear maps with nothing in between: ONE GEMM with the composed matrix (ops.ComposedLinear)
            return ops.ComposedLinear.apply(u, *pair)

        if bn is None:
            return tail(t)
        ops.expect_bn(bn.training or bn.running_mean is None)
        try:
            z = tail(t)
        finally:
            ops.expect_bn(False)
        if defer:                      # the caller fuses this BatchNorm into its next op (decoder skip bridge)
            return z, bn
        return EELUnet._bn(bn, z, relu, single_conv_consumer=single_conv_consumer)

    @staticmethod
    def _conv_bn(conv, bn, x, relu=True, defer=False, single_conv_consumer=False, shift_out=False):
        """conv3x3 -> BatchNorm[-> ReLU]; in training the conv's epilogue also delivers the BatchNorm sums"""
        f = ops.folded(conv.weight)
        if isinstance(x, _Bridge):
            fs = ops.folded_split(conv.weight) if f is not None else None
            if x.bn is None and fs is not None:
                # inference: (upconv + edge feature) and the encoder skip go into the conv as two tensors
                return ops.conv3x3_folded_2src(ops.add2(x.z, x.b), x.e, fs, f[1], relu)
            if x.bn is None:
                x = ops.AddInterleave.apply(x.z, x.b, x.e)
            elif f is not None or not ops.BridgeConv3x3.supported(x.z, conv.weight):
                x = EELUnet._bn_add_interleave(x.bn, x.z, x.b, x.e)   # (not fusable after all: the interleaved tensor, then a plain conv)
        if f is not None:
            return ops.conv3x3_folded(x, f[0], f[1], relu)
        ops.expect_bn(bn.training or bn.running_mean is None)
        try:
            if isinstance(x, _Bridge):
                ub = x.bn
                z = ops.BridgeConv3x3.apply(x.z, ub.weight, ub.bias, ub.running_mean, ub.running_var, EELUnet._bn_mode(ub),
                                            ub.momentum if ub.momentum is not None else 0.1, ub.eps, x.b, x.e, True,
                                            conv.weight, conv.bias)
            else:
                z = ops.conv3x3(x, conv.weight, conv.bias, False)
        finally:
            ops.expect_bn(False)
        if defer:
            return z, bn
        return EELUnet._bn(bn, z, relu, single_conv_consumer=single_conv_consumer, shift_out=shift_out)

    def _conv_block(self, blk, x, defer=False, single_consumer=False):
        """defer: return (pre-BatchNorm tensor, BatchNorm) for the block's LAST BatchNorm + ReLU (fused into the PGR that follows);
        single_consumer: the block's output feeds exactly one skip bridge (whose backward then delivers the last BatchNorm's sums)"""
        x = self._conv_bn(blk[0], blk[1], x, single_conv_consumer=True)
        return self._conv_bn(blk[3], blk[4], x, defer=defer, single_conv_consumer=single_consumer)

    def _mlp_conv_block(self, blk, x, defer=False):
        # bf16 training path: the BatchNorm + ReLU in front of the token MLP stores its result through ShiftedChannel, so
        # to_patch reads it as is (no eel_shift_channels copy); inference folds that BatchNorm into the conv instead
        C = blk[0].weight.shape[0]
        pre = (EELUnet.fused_shift and self.compute_dtype == torch.bfloat16 and ops.folded(blk[0].weight) is None and C % 128 == 0
               and blk[0].weight.shape[1] % 64 == 0)
        x = self._conv_bn(blk[0], blk[1], x, shift_out=pre)
        return self._capmlp(blk[3], x, bn=blk[4], relu=True, defer=defer, pre_shifted=pre)

    def _pool(self, d):
        """end of an encoder stage: (skip tensor, 2x2 max-pooled tensor).  `d` is the finished stage output or
        (pre-BatchNorm tensor, BatchNorm): then BatchNorm + ReLU + pool are one pass (ops.BNReluPool) and, in the backward,
        the skip and pool gradients meet inside the BatchNorm backward"""
        if isinstance(d, tuple):
            z, bn = d
            if z.shape[1] % 2 == 0 and z.shape[2] % 2 == 0:
                training = self._bn_mode(bn)
                return ops.BNReluPool.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                            bn.momentum if bn.momentum is not None else 0.1, bn.eps, True)
            d = self._bn(bn, z, True)
        return d, ops.MaxPool2.apply(d)

    def _pgr_block(self, m, d):
        """PredictionGuidedRefinement on a decoder block's output; `d` is either the finished block output or
        (pre-BatchNorm tensor, BatchNorm) whose BatchNorm + ReLU are then fused into the PGR pass"""
        if isinstance(d, tuple):
            z, bn = d
            if ops.bn_pgr_supported(z):
                training = self._bn_mode(bn)
                return ops.BNReluPGR.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, training,
                                           bn.momentum if bn.momentum is not None else 0.1, bn.eps, m.conv.weight, m.conv.bias, True)
            d = self._bn(bn, z, True)
        return self._pgr(m, d)

    def _upconv(self, blk, x, defer=False):
        """ConvT -> BatchNorm.  defer: return (z, bn) so that the caller fuses the BatchNorm into the skip bridge"""
        f = ops.folded(blk[0].weight)
        if f is not None:
            return ops.convt2x2_folded(x, f[0], f[1])
        ops.expect_bn(blk[1].training or blk[1].running_mean is None)
        try:
            z = ops.ConvT2x2.apply(x, blk[0].weight, blk[0].bias)
        finally:
            ops.expect_bn(False)
        if defer:
            return z, blk[1]
        return self._bn(blk[1], z, False)

    def _mlp_upconv(self, blk, x, defer=False, single_conv_consumer=False):
        return self._capmlp(blk[1], ops.ConvT2x2.apply(x, blk[0].weight, blk[0].bias), bn=blk[2], relu=False, defer=defer,
                            single_conv_consumer=single_conv_consumer)

    def _bridge(self, up, b, e):
        """(upconv output + edge feature) interleaved with the encoder skip (reference models/EELUnet.py:422-426); `up` is
        either the finished upconv output or (pre-BatchNorm tensor, BatchNorm) when its BatchNorm is fused in here"""
        if isinstance(up, tuple):
            if EELUnet.fused_bridge and up[0].dtype == torch.bfloat16:
                return _Bridge(up[1], up[0], b, e)          # consumed by the decoder block's first conv (_conv_bn)
            return self._bn_add_interleave(up[1], up[0], b, e)
        if EELUnet.fused_bridge and up.dtype == torch.bfloat16 and getattr(ops._TLS, "folded", None) is not None:
            return _Bridge(None, up, b, e)                  # inference: the upconv's BatchNorm is already folded in (bn = None)
        return ops.AddInterleave.apply(up, b, e)

    @staticmethod
    def _pgr(m, x):
        return ops.PGR.apply(x, m.conv.weight, m.conv.bias)

    # ------------------------------------------------------------------------------------------
    def forward(self, x):
        if not x.is_cuda:
            raise EelError("eel_unet_b200.EELUnet runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.dim() != 4 or x.shape[2] % 16 or x.shape[3] % 16:
            raise EelError("input must be N x C x H x W with H and W multiples of 16 (got %s): the reference's center_crop path "
                           "for other sizes (models/EELUnet.py:376-382) is not implemented" % (tuple(x.shape),))
        if x.requires_grad and torch.is_grad_enabled():
            raise EelError("the input image is treated as a leaf without gradient (the first conv computes no data gradient); "
                           "pass x.detach() -- input gradients (saliency maps, adversarial examples) are not supported")
        # kernels launch on the CURRENT device's current stream: make the input's device current for the whole forward
        # (autograd's backward threads do the same for the gradients' device)
        with torch.cuda.device(x.device):
            return self._forward(x)

    def _forward(self, x):
        # the input conversion is launched before the host-side bookkeeping of the weight tables (version checks over ~70
        # parameters): after a synchronisation (validation loops, loss.item() per step) the GPU starts a few hundred
        # microseconds earlier
        a = ops.nchw_to_nhwc(x, self.compute_dtype)
        fold = None
        if self.compute_dtype == torch.bfloat16:
            self._weight_packer().refresh(x.device)      # publishes the packed operands on the weights themselves
            if not self.training and not torch.is_grad_enabled():
                # inference: eval-mode BatchNorms are folded into their producers' weights (ops.FoldedPacker)
                fold = self._folded_packer()
                if all(not b.training and b.running_mean is not None for b in fold.bns()):
                    fold.refresh(x.device)
                else:
                    fold = None
        ops.set_folded(fold)
        self._composed_packer().refresh(x.device)

        enc1, p = self._pool(self._conv_block(self.enc1[0], a, defer=True))
        enc2, p = self._pool(self._conv_block(self.enc2[0], p, defer=True))
        enc3, p = self._pool(self._mlp_conv_block(self.enc3[0], p, defer=True))
        enc4, p = self._pool(self._mlp_conv_block(self.enc4[0], p, defer=True))

        bt = self.bottleneck
        b = self._bn(bt[0], p, False, producer_bias=False, single_conv_consumer=True)
        b = ops.Conv3x3.apply(b, bt[1].weight, bt[1].bias, True)
        b = ops.Relu.apply(self._capmlp(bt[3], b))
        b, edge_5 = self._pgr(self.pred5, b)

        # edge branch (reference models/EELUnet.py:300-328, 415-418)
        # (the mlp_upconv outputs feed exactly one conv3x3: its data-gradient launch delivers their BatchNorm's backward sums)
        e4 = self._mlp_conv_block(self.edge_upconv_4[1], self._mlp_upconv(self.edge_upconv_4[0], b, single_conv_consumer=True))
        e3 = self._mlp_conv_block(self.edge_upconv_3[1], self._mlp_upconv(self.edge_upconv_3[0], e4, single_conv_consumer=True))
        e2 = ops.HFT.apply(self._upconv(self.edge_upconv_2[0], e3), self.edge_upconv_2[1].mask_range)
        e2 = self._conv_block(self.edge_upconv_2[2], e2)
        e1 = ops.HFT.apply(self._upconv(self.edge_upconv_1[0], e2), self.edge_upconv_1[1].mask_range)
        e1 = self._conv_block(self.edge_upconv_1[2], e1, single_consumer=True)      # e1 feeds only dec1's bridge

        # decoder (reference models/EELUnet.py:421-465)
        d = self._mlp_conv_block(self.dec4, self._bridge(self._mlp_upconv(self.upconv4, b, defer=True), e4, enc4), defer=True)
        d, edge_4 = self._pgr_block(self.pred4, d)
        d = self._mlp_conv_block(self.dec3, self._bridge(self._mlp_upconv(self.upconv3, d, defer=True), e3, enc3), defer=True)
        d, edge_3 = self._pgr_block(self.pred3, d)
        d = self._conv_block(self.dec2, self._bridge(self._upconv(self.upconv2, d, defer=True), e2, enc2), defer=True)
        d, edge_2 = self._pgr_block(self.pred2, d)
        d = self._conv_block(self.dec1, self._bridge(self._upconv(self.upconv1, d, defer=True), e1, enc1), defer=True)
        d, edge_1 = self._pgr_block(self.pred1, d)

        seg = ops.Head.apply(d, self.final[0].weight, self.final[0].bias, self.final[1].weight, self.final[1].bias)
        return seg, [edge_5, edge_4, edge_3, edge_2, edge_1]
