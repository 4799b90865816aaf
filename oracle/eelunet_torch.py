"""fp32 PyTorch-CPU restatement of the EELUnet forward and of edge_BceDiceLoss.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Never imported by the product package.

It is a *functional* restatement over a reference-format ``state_dict`` (the 365 keys of
SURVEY.md section 8b): no nn.Module tree, no hooks, no file output.  Gradients come from torch autograd
over these same functions.  Each function cites the reference lines it follows
(paths relative to /root/reference).

Parity pinning: checked against the imported reference in the build container
(tests/test_oracle_vs_reference.py) and against tests/golden/*.npz, which were produced by the
reference itself (tests/golden/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F

EDGE_WEIGHTS = (0.1, 0.2, 0.3, 0.4, 0.5)  # utils/Loss.py:108-112, for gt_pre5..gt_pre1


# ----------------------------------------------------------------------------- building blocks
def _bn(sd, p, x, train, new_stats):
    """nn.BatchNorm2d, eps 1e-5, momentum 0.1 (models/EELUnet.py:339,343,352,356,365,373,256)."""
    w, b = sd[p + ".weight"], sd[p + ".bias"]
    rm, rv = sd[p + ".running_mean"], sd[p + ".running_var"]
    if not train:
        return F.batch_norm(x, rm, rv, w, b, False, 0.1, 1e-5)
    rm2, rv2 = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm2, rv2, w, b, True, 0.1, 1e-5)
    if new_stats is not None:
        new_stats[p + ".running_mean"] = rm2.detach()
        new_stats[p + ".running_var"] = rv2.detach()
        new_stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    return y


def _conv3(sd, p, x):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)


def _shift(x):
    """ShiftedChannel, models/EELUnet.py:88-97: circular rolls of three channel quarters."""
    c = x.shape[1]
    s = int(c * 0.25)
    out = x.clone()
    out[:, :s] = torch.roll(x[:, :s], 1, 2)
    out[:, s:2 * s] = torch.roll(x[:, s:2 * s], -1, 2)
    out[:, 2 * s:3 * s] = torch.roll(x[:, 2 * s:3 * s], 1, 3)
    return out


def _capmlp(sd, p, x):
    """ChannelAwarePatchedMLP, models/EELUnet.py:114-123 (+ ChannelAttention :57-80)."""
    n, c, h, w = x.shape
    t = F.conv2d(_shift(x), sd[p + ".to_patch.weight"], sd[p + ".to_patch.bias"])
    g = t.mean(dim=(2, 3), keepdim=True)
    g = F.relu(F.conv2d(g, sd[p + ".channel_attention.fc1.weight"], sd[p + ".channel_attention.fc1.bias"]))
    g = torch.sigmoid(F.conv2d(g, sd[p + ".channel_attention.fc2.weight"], sd[p + ".channel_attention.fc2.bias"]))
    t = (t * g).permute(0, 2, 3, 1).reshape(n, h * w, -1)
    t = F.gelu(F.linear(t, sd[p + ".mlp.0.weight"], sd[p + ".mlp.0.bias"]))
    t = F.linear(t, sd[p + ".mlp.2.weight"], sd[p + ".mlp.2.bias"])
    t = t.reshape(n, h, w, -1).permute(0, 3, 1, 2)
    return F.conv2d(t, sd[p + ".to_space.weight"], sd[p + ".to_space.bias"])


def _conv_block(sd, p, x, train, ns):
    """EELUnet.conv_block, models/EELUnet.py:335-345."""
    x = F.relu(_bn(sd, p + ".1", _conv3(sd, p + ".0", x), train, ns))
    return F.relu(_bn(sd, p + ".4", _conv3(sd, p + ".3", x), train, ns))


def _mlp_conv_block(sd, p, x, train, ns):
    """EELUnet.mlp_conv_block, models/EELUnet.py:347-359."""
    x = F.relu(_bn(sd, p + ".1", _conv3(sd, p + ".0", x), train, ns))
    return F.relu(_bn(sd, p + ".4", _capmlp(sd, p + ".3", x), train, ns))


def _convT(sd, p, x):
    return F.conv_transpose2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=2)


def _upconv(sd, p, x, train, ns):
    """EELUnet.upconv_block, models/EELUnet.py:361-366 (no ReLU)."""
    return _bn(sd, p + ".1", _convT(sd, p + ".0", x), train, ns)


def _mlp_upconv(sd, p, x, train, ns):
    """EELUnet.mlp_upconv_block, models/EELUnet.py:368-374 (no ReLU)."""
    return _bn(sd, p + ".2", _capmlp(sd, p + ".1", _convT(sd, p + ".0", x)), train, ns)


def hft(x, mask_range=20):
    """HighFourierTransform.forward, models/EELUnet.py:153-191.

    The reference shifts the spectrum, zeroes the centred square [c-r, c+r) and shifts back; here the
    same mask is built directly in unshifted frequency order.
    """
    h, w = x.shape[-2:]
    if x.dtype in (torch.bfloat16, torch.float16):
        x = x.float()          # torch.fft has no half-precision kernels for these sizes (autocast keeps FFTs in fp32)
    r = min(mask_range, h // 2, w // 2)
    mh = torch.ones(h)
    mw = torch.ones(w)
    # centred index j in [c-r, c+r) <-> unshifted frequency (j - c) mod n, i.e. -r .. r-1
    fh = torch.arange(-r, r) % h
    fw = torch.arange(-r, r) % w
    mh[fh] = 0
    mw[fw] = 0
    keep = 1.0 - torch.outer(1.0 - mh, 1.0 - mw)
    return torch.abs(torch.fft.ifft2(torch.fft.fft2(x) * keep.to(x.device)))


def _pgr(sd, p, x):
    """PredictionGuidedRefinement, models/EELUnet.py:200-203."""
    s = torch.sigmoid(F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"]))
    return x + x * s, s


def _interleave(a, b):
    """FeatureInterleaveBridge, models/EELUnet.py:132-141: out[:,2c]=a[:,c], out[:,2c+1]=b[:,c]."""
    n, c, h, w = a.shape
    return torch.stack((a, b), dim=2).reshape(n, 2 * c, h, w)


def _head(sd, x):
    """final = LayerNorm(channels_first, eps 1e-6) + conv1x1, then sigmoid (models/EELUnet.py:217-225,330-333,467-469)."""
    u = x.mean(1, keepdim=True)
    v = (x - u).pow(2).mean(1, keepdim=True)
    y = (x - u) / torch.sqrt(v + 1e-6)
    y = sd["final.0.weight"][None, :, None, None] * y + sd["final.0.bias"][None, :, None, None]
    return torch.sigmoid(F.conv2d(y, sd["final.1.weight"], sd["final.1.bias"]))


# ----------------------------------------------------------------------------- whole forward
def forward(sd, x, train=True, new_stats=None):
    """EELUnet.forward, models/EELUnet.py:384-471.  Returns (seg, [edge_5..edge_1])."""
    ns = new_stats
    enc1 = _conv_block(sd, "enc1.0", x, train, ns)
    enc2 = _conv_block(sd, "enc2.0", F.max_pool2d(enc1, 2), train, ns)
    enc3 = _mlp_conv_block(sd, "enc3.0", F.max_pool2d(enc2, 2), train, ns)
    enc4 = _mlp_conv_block(sd, "enc4.0", F.max_pool2d(enc3, 2), train, ns)
    b = _bn(sd, "bottleneck.0", F.max_pool2d(enc4, 2), train, ns)
    b = F.relu(_conv3(sd, "bottleneck.1", b))
    b = F.relu(_capmlp(sd, "bottleneck.3", b))
    b, e5 = _pgr(sd, "pred5", b)

    # edge branch, models/EELUnet.py:300-328, 415-418
    ed4 = _mlp_conv_block(sd, "edge_upconv_4.1", _mlp_upconv(sd, "edge_upconv_4.0", b, train, ns), train, ns)
    ed3 = _mlp_conv_block(sd, "edge_upconv_3.1", _mlp_upconv(sd, "edge_upconv_3.0", ed4, train, ns), train, ns)
    ed2 = _conv_block(sd, "edge_upconv_2.2", hft(_upconv(sd, "edge_upconv_2.0", ed3, train, ns)), train, ns)
    ed1 = _conv_block(sd, "edge_upconv_1.2", hft(_upconv(sd, "edge_upconv_1.0", ed2, train, ns)), train, ns)

    # decoder, models/EELUnet.py:421-465 (center_crop is the identity for H, W multiples of 16)
    d = _mlp_upconv(sd, "upconv4", b, train, ns) + ed4
    d = _mlp_conv_block(sd, "dec4", _interleave(d, enc4), train, ns)
    d, e4 = _pgr(sd, "pred4", d)
    d = _mlp_upconv(sd, "upconv3", d, train, ns) + ed3
    d = _mlp_conv_block(sd, "dec3", _interleave(d, enc3), train, ns)
    d, e3 = _pgr(sd, "pred3", d)
    d = _upconv(sd, "upconv2", d, train, ns) + ed2
    d = _conv_block(sd, "dec2", _interleave(d, enc2), train, ns)
    d, e2 = _pgr(sd, "pred2", d)
    d = _upconv(sd, "upconv1", d, train, ns) + ed1
    d = _conv_block(sd, "dec1", _interleave(d, enc1), train, ns)
    d, e1 = _pgr(sd, "pred1", d)
    return _head(sd, d), [e5, e4, e3, e2, e1]


# ----------------------------------------------------------------------------- loss
def bce_dice(pred, target, wb=1.0, wd=1.0):
    """BceDiceLoss, utils/Loss.py:28-73.  nn.BCELoss clamps each log term at -100."""
    n = pred.shape[0]
    p = pred.reshape(n, -1)
    t = target.reshape(n, -1)
    bce = -(t * torch.clamp(torch.log(p), min=-100.0) + (1 - t) * torch.clamp(torch.log(1 - p), min=-100.0)).mean()
    dice = 1 - ((2 * (p * t).sum(1) + 1) / (p.sum(1) + t.sum(1) + 1)).sum() / n
    return wd * dice + wb * bce


def edge_bce_dice_loss(edges, seg, target, wb=1.0, wd=1.0):
    """edge_BceDiceLoss.forward, utils/Loss.py:97-113."""
    loss = bce_dice(seg, target, wb, wd)
    for k, (e, wk) in enumerate(zip(edges, EDGE_WEIGHTS)):
        s = 16 >> k
        t = F.max_pool2d(target, s, s) if s > 1 else target
        loss = loss + wk * bce_dice(e, t, wb, wd)
    return loss


def dice_metric(seg, target):
    """evaluate.py:94-117: foreground Dice = 2TP/(2TP+FP+FN+1e-7) at threshold 0.5."""
    p = (seg > 0.5).reshape(-1)
    t = (target == 1).reshape(-1)
    tp = (p & t).sum().item()
    fp = (p & ~t).sum().item()
    fn = (~p & t).sum().item()
    return 2 * tp / (2 * tp + fp + fn + 1e-7)


def train_step(sd, x, target, train=True):
    """One fwd + loss + bwd on CPU; returns (loss, seg, edges, grads dict, new running stats)."""
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running_" not in k)
              for k, v in sd.items()}
    ns = {}
    seg, edges = forward(params, x, train, ns)
    loss = edge_bce_dice_loss(edges, seg, target)
    loss.backward()
    grads = {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}
    return loss.detach(), seg.detach(), [e.detach() for e in edges], grads, ns


def unet_forward(sd, x):
    """Unet.forward, models/Unet.py:58-98 (no BatchNorm; logits out; skip = concat((dec, enc), dim=1))."""
    def blk(p, t):
        t = F.relu(F.conv2d(t, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1))
        return F.relu(F.conv2d(t, sd[p + ".2.weight"], sd[p + ".2.bias"], padding=1))

    e1 = blk("enc1", x)
    e2 = blk("enc2", F.max_pool2d(e1, 2))
    e3 = blk("enc3", F.max_pool2d(e2, 2))
    e4 = blk("enc4", F.max_pool2d(e3, 2))
    d = blk("bottleneck", F.max_pool2d(e4, 2))
    for k, skip in ((4, e4), (3, e3), (2, e2), (1, e1)):
        d = F.conv_transpose2d(d, sd["upconv%d.0.weight" % k], sd["upconv%d.0.bias" % k], stride=2)
        d = blk("conv%d" % k, torch.cat((d, skip), dim=1))
    return F.conv2d(d, sd["final_conv.weight"], sd["final_conv.bias"])


def gelu_exact(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))
