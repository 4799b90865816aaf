"""Per-kernel parity (GPU): every autograd op of eel_unet_b200.ops, forward and backward, against the
plain PyTorch fp32 formulation of the reference op it replaces (computed in fp64 on the same inputs).

Tolerances (relative L2 per tensor): fp32 storage 2e-5 forward / 1e-4 gradients; bf16 storage 2e-2
(the north_star's stated bf16 tolerance), with the checker fed the bf16-rounded inputs.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def tol(dtype, fwd=True):
    if dtype == torch.bfloat16:
        return 2e-2
    return 2e-5 if fwd else 1e-4


def run_case(fn_mine, fn_ref, inputs, params, dtype, out_nhwc=True, atol_scale=1.0, all_nhwc=False):
    """inputs: list of NCHW fp32 tensors (activations); params: list of fp32 parameter tensors.

    fn_mine(nhwc activations (dtype, requires_grad), params) -> NHWC output (or tuple)
    fn_ref(nchw fp64 activations, fp64 params) -> NCHW output (or tuple)
    """
    from eel_unet_b200 import ops  # noqa: F401

    torch.manual_seed(0)
    acts = [nhwc(x).to(dtype).requires_grad_(True) for x in inputs]
    ps = [p.clone().requires_grad_(True) for p in params]
    outs = fn_mine(acts, ps)
    outs = outs if isinstance(outs, tuple) else (outs,)
    racts = [nchw(a.detach().double()).requires_grad_(True) for a in acts]
    rps = [p.detach().double().requires_grad_(True) for p in params]
    routs = fn_ref(racts, rps)
    routs = routs if isinstance(routs, tuple) else (routs,)
    gs = []
    for o, r in zip(outs, routs):
        # (all_nhwc: every 4-D output is NHWC even when C == H == W makes the shapes coincide)
        o_cmp = nchw(o.float()) if (o.dim() == 4 and out_nhwc and o.dtype == dtype and (all_nhwc or o.shape != r.shape)) else o.float()
        e = rel(o_cmp, r)
        assert e < tol(dtype) * atol_scale, "forward mismatch %g" % e
        g = torch.randn_like(r)
        gs.append(g)
    # backward
    mine_g = []
    for o, g, r in zip(outs, gs, routs):
        if o.dim() == 4 and (all_nhwc or o.shape != r.shape):
            mine_g.append(nhwc(g).to(o.dtype))
        else:
            mine_g.append(g.to(o.dtype))
    torch.autograd.backward(outs, mine_g)
    # the checker sees the same (possibly bf16-rounded) upstream gradients
    rg = []
    for o, g, r in zip(outs, mine_g, routs):
        gg = g.double()
        rg.append(nchw(gg) if (all_nhwc or gg.shape != r.shape) else gg)
    torch.autograd.backward(routs, rg)
    for a, ra in zip(acts, racts):
        e = rel(nchw(a.grad.float()), ra.grad)
        assert e < tol(dtype, False) * atol_scale, "input-grad mismatch %g" % e
    for p, rp in zip(ps, rps):
        e = rel(p.grad, rp.grad)
        assert e < tol(dtype, False) * atol_scale, "param-grad mismatch %g (shape %s)" % (e, tuple(p.shape))


DTYPES = [torch.float32, torch.bfloat16]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 8, 12, 20, 16), (1, 64, 16, 16, 72), (3, 16, 9, 7, 8),
                                   (2, 64, 24, 20, 64), (1, 128, 16, 16, 256), (2, 256, 8, 8, 128), (1, 64, 32, 40, 512),
                                   (3, 192, 5, 9, 64), (2, 128, 40, 24, 64), (1, 64, 16, 16, 64), (1, 128, 70, 36, 128)])
@pytest.mark.parametrize("relu", [False, True])
def test_conv3x3(dtype, shape, relu):
    from eel_unet_b200 import ops

    n, cin, h, w, cout = shape
    x = torch.randn(n, cin, h, w, device=DEV)
    wt = torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(9 * cin)
    b = torch.randn(cout, device=DEV)

    def mine(a, p):
        return ops.Conv3x3.apply(a[0], p[0], p[1], relu)

    def ref(a, p):
        y = F.conv2d(a[0], p[0], p[1], padding=1)
        return F.relu(y) if relu else y

    # bf16 + fused ReLU: outputs within bf16 rounding of 0 flip the mask, a 100 % error on those elements
    run_case(mine, ref, [x], [wt, b], dtype, atol_scale=4.0 if (relu and dtype == torch.bfloat16) else 1.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv3x3_first_layer_three_channels(dtype):
    from eel_unet_b200 import ops

    x = torch.randn(2, 3, 32, 32, device=DEV)
    wt = torch.randn(64, 3, 3, 3, device=DEV) / math.sqrt(27)
    b = torch.randn(64, device=DEV)
    a = ops.nchw_to_nhwc(x, dtype)
    y = ops.Conv3x3.apply(a, wt.requires_grad_(True), b.requires_grad_(True), False)
    r = F.conv2d(nchw(a.double()), wt.double(), b.double(), padding=1)
    assert rel(nchw(y.float()), r) < tol(dtype)
    y.backward(torch.ones_like(y))   # leaf input: only wgrad/bias paths run
    assert wt.grad is not None and torch.isfinite(wt.grad).all()


@pytest.mark.parametrize("shape", [(2, 32, 32), (1, 16, 48), (3, 18, 10)])
def test_first_conv_tensor_core_path(shape):
    """bf16 mode runs the 3 -> 64 first conv (reference models/EELUnet.py:338) through a compact im2col + the tcgen05 GEMM /
    weight-gradient kernels (ops.StemConv): forward, weight and bias gradients, and the fused BatchNorm sums"""
    from eel_unet_b200 import _lib, ops

    n, h, w = shape
    x = torch.randn(n, 3, h, w, device=DEV)
    wt = (torch.randn(64, 3, 3, 3, device=DEV) / math.sqrt(27)).requires_grad_(True)
    b = torch.randn(64, device=DEV).requires_grad_(True)
    a = ops.nchw_to_nhwc(x, torch.bfloat16)
    rec = []
    _lib.set_profiler(rec)
    try:
        y = ops.conv3x3(a, wt, b, False)
    finally:
        _lib.set_profiler(None)
    assert "eel_stem_im2col" in [r[0] for r in rec]
    wr, br = wt.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    r = F.conv2d(nchw(a.double()), wr, br, padding=1)
    assert rel(nchw(y.float()), r) < 2e-2
    g = torch.randn_like(r)
    gm = nhwc(g).to(torch.bfloat16)
    y.backward(gm)
    r.backward(nchw(gm.double()))
    assert rel(wt.grad, wr.grad) < 2e-2 and rel(b.grad, br.grad) < 2e-2
    # in front of a training-mode BatchNorm: z is stored WITHOUT the bias (it cancels), its sums come out of the epilogue
    ops.expect_bn(True)
    try:
        z = ops.conv3x3(a, wt.detach(), b.detach(), False)
    finally:
        ops.expect_bn(False)
    sums, skipped = ops._take(z, "_eel_bn_sums")
    assert torch.equal(skipped, b.detach())
    r0 = F.conv2d(nchw(a.double()), wt.detach().double(), None, padding=1)
    assert rel(nchw(z.float()), r0) < 2e-2
    zs = z.double().reshape(-1, 64)
    assert rel(sums[0], zs.sum(0)) < 1e-3 and rel(sums[1], (zs * zs).sum(0)) < 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 6, 10, 8), (1, 64, 8, 8, 32), (2, 128, 8, 16, 64), (1, 64, 3, 5, 128),
                                   (1, 256, 16, 16, 128)])
def test_convt2x2(dtype, shape):
    from eel_unet_b200 import ops

    n, cin, h, w, cout = shape
    x = torch.randn(n, cin, h, w, device=DEV)
    wt = torch.randn(cin, cout, 2, 2, device=DEV) / math.sqrt(cin)
    b = torch.randn(cout, device=DEV)
    run_case(lambda a, p: ops.ConvT2x2.apply(a[0], p[0], p[1]),
             lambda a, p: F.conv_transpose2d(a[0], p[0], p[1], stride=2), [x], [wt, b], dtype)


def ref_shift(x):
    s = int(x.shape[1] * 0.25)
    return torch.cat([x[:, :s].roll(1, 2), x[:, s:2 * s].roll(-1, 2), x[:, 2 * s:3 * s].roll(1, 3), x[:, 3 * s:]], 1)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shift", [False, True])
@pytest.mark.parametrize("shape", [(2, 32, 8, 12, 24), (1, 256, 16, 16, 64), (2, 64, 12, 20, 256), (1, 256, 5, 7, 512),
                                   (3, 512, 4, 4, 64)])
def test_linear(dtype, shift, shape):
    from eel_unet_b200 import ops

    n, k, h, w, nout = shape
    x = torch.randn(n, k, h, w, device=DEV)
    wt = torch.randn(nout, k, 1, 1, device=DEV) / math.sqrt(k)
    b = torch.randn(nout, device=DEV)
    run_case(lambda a, p: ops.Linear.apply(a[0], p[0], p[1], shift),
             lambda a, p: F.conv2d(ref_shift(a[0]) if shift else a[0], p[0], p[1]), [x], [wt, b], dtype)


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", [(2, 256, 8, 12, 64), (1, 512, 5, 7, 64), (3, 128, 4, 4, 64)])
def test_bn_stored_through_the_shift_feeds_to_patch(relu, shape):
    """mlp_conv_block (reference models/EELUnet.py:350-357 with :88-97,118): BatchNorm + ReLU -> ShiftedChannel -> to_patch.  In bf16
    mode the BatchNorm stores its result through the shift (eel_bn_act_shift_fwd) and the Linear takes it as is (shift="pre");
    the Linear's data gradient comes back through the adjoint shift, the BatchNorm backward is the plain one"""
    from eel_unet_b200 import _lib, ops

    n, c, h, w, nout = shape
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, device=DEV) * 1.5 + 0.3
    g = torch.rand(c, device=DEV) + 0.5
    b = torch.randn(c, device=DEV) * 0.5
    wt = torch.randn(nout, c, 1, 1, device=DEV) / math.sqrt(c)
    bias = torch.randn(nout, device=DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)

    def mine(a, p):
        y = ops.BNAct.apply(a[0], p[0], p[1], rm, rv, True, relu, 0.1, 1e-5, False, False, True)
        return ops.Linear.apply(y, p[2], p[3], "pre")

    def ref(a, p):
        y = F.batch_norm(a[0], None, None, p[0], p[1], True, 0.1, 1e-5)
        y = F.relu(y) if relu else y
        y = y + (y.to(dtype).double() - y).detach()          # the Linear reads the activation as stored
        return F.conv2d(ref_shift(y), p[2], p[3])

    rec = []
    _lib.set_profiler(rec)
    try:
        run_case(mine, ref, [x], [g, b, wt, bias], dtype, atol_scale=2.0)
    finally:
        _lib.set_profiler(None)
    names = [r[0] for r in rec]
    assert "eel_bn_act_shift_fwd" in names and "eel_shift_channels" not in names


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 256, 8, 16, 256), (1, 256, 16, 16, 512), (2, 24, 5, 7, 40), (1, 256, 4, 8, 1024),
                                   (1, 64, 8, 8, 72)])
def test_composed_linear(dtype, shape):
    """to_space(mlp[2](x)) as one GEMM with the composed matrix (reference models/EELUnet.py:109-111,121-122): forward,
    input gradient and the FOUR parameter gradients of the two reference layers against the two-layer formulation"""
    from eel_unet_b200 import ops

    n, k, h, w, cout = shape
    x = torch.randn(n, k, h, w, device=DEV)
    w1 = torch.randn(cout, k, device=DEV) / math.sqrt(k)            # mlp[2]: Linear(k -> cout)
    b1 = torch.randn(cout, device=DEV)
    w2 = torch.randn(cout, cout, 1, 1, device=DEV) / math.sqrt(cout)  # to_space: 1x1 conv cout -> cout
    b2 = torch.randn(cout, device=DEV)

    def ref(a, p):
        t = F.linear(a[0].permute(0, 2, 3, 1), p[0], p[1]).permute(0, 3, 1, 2)
        return F.conv2d(t, p[2], p[3])

    run_case(lambda a, p: ops.ComposedLinear.apply(a[0], p[0], p[1], p[2], p[3]), ref, [x], [w1, b1, w2, b2], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_bn_act(dtype, relu, training):
    from eel_unet_b200 import ops

    n, c, h, w = 3, 32, 10, 6
    x = torch.randn(n, c, h, w, device=DEV) * 2 + 3.0   # mean >> 0 exercises the shifted-moment statistics
    g = torch.rand(c, device=DEV) + 0.5
    b = torch.randn(c, device=DEV)
    rm0 = torch.randn(c, device=DEV) * 0.1
    rv0 = torch.rand(c, device=DEV) + 0.5
    rm, rv = rm0.clone(), rv0.clone()
    rrm, rrv = rm0.double(), rv0.double()

    def mine(a, p):
        return ops.BNAct.apply(a[0], p[0], p[1], rm, rv, training, relu, 0.1, 1e-5, False)

    def ref(a, p):
        y = F.batch_norm(a[0], rrm, rrv, p[0], p[1], training, 0.1, 1e-5)
        return F.relu(y) if relu else y

    run_case(mine, ref, [x], [g, b], dtype)
    if training:
        t = 1e-5 if dtype == torch.float32 else 1e-2
        assert rel(rm, rrm) < t and rel(rv, rrv) < t
    else:
        assert torch.equal(rm, rm0) and torch.equal(rv, rv0)


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("shape", [(2, 64, 40, 24, 64), (1, 128, 36, 20, 128), (2, 512, 8, 8, 256), (1, 64, 16, 16, 128)])
def test_bn_backward_sums_from_conv_dgrad_epilogue(relu, shape):
    """conv_block (reference models/EELUnet.py:338-344): BatchNorm(+ReLU) -> conv3x3.  In bf16 mode the conv's data-gradient
    launch also accumulates the BatchNorm's backward sums (eel_tc_conv3x3_dgrad_bnsums), the BatchNorm backward is apply-only"""
    from eel_unet_b200 import _lib, ops

    n, c, h, w, cout = shape
    dtype = torch.bfloat16
    x = torch.randn(n, c, h, w, device=DEV) * 1.5 + 0.5
    g = torch.rand(c, device=DEV) + 0.5
    b = torch.randn(c, device=DEV) * 0.5
    wt = torch.randn(cout, c, 3, 3, device=DEV) / math.sqrt(9 * c)
    bias = torch.randn(cout, device=DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)

    def mine(a, p):
        y = ops.BNAct.apply(a[0], p[0], p[1], rm, rv, True, relu, 0.1, 1e-5, False, True)
        return ops.Conv3x3.apply(y, p[2], p[3], False)

    def ref(a, p):
        y = F.batch_norm(a[0], None, None, p[0], p[1], True, 0.1, 1e-5)
        y = F.relu(y) if relu else y
        # the conv reads the activation as stored (bf16)
        y = y + (y.to(dtype).double() - y).detach()
        return F.conv2d(y, p[2], p[3], padding=1)

    rec = []
    _lib.set_profiler(rec)
    try:
        run_case(mine, ref, [x], [g, b, wt, bias], dtype, atol_scale=2.0)
    finally:
        _lib.set_profiler(None)
    names = [r[0] for r in rec]
    assert "eel_tc_conv3x3_dgrad_bnsums" in names and "eel_bn_act_bwd_apply" in names and "eel_bn_act_bwd" not in names


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("shape", [(3, 32, 10, 6), (2, 64, 16, 24), (1, 8, 4, 2)])
def test_bn_relu_pool(dtype, training, shape):
    """end of an encoder stage (reference models/EELUnet.py:387-406): BatchNorm -> ReLU -> (skip, MaxPool2d(2)) in one pass;
    backward with BOTH upstream gradients (skip + pool) meeting inside the BatchNorm backward"""
    from eel_unet_b200 import ops

    n, c, h, w = shape
    x = torch.randn(n, c, h, w, device=DEV) * 2 + 1.0
    g = torch.rand(c, device=DEV) + 0.5
    b = torch.randn(c, device=DEV) * 0.5
    rm0 = torch.randn(c, device=DEV) * 0.1 + 1.0
    rv0 = torch.rand(c, device=DEV) + 3.5
    rm, rv = rm0.clone(), rv0.clone()
    rrm, rrv = rm0.double(), rv0.double()

    def mine(a, p):
        return ops.BNReluPool.apply(a[0], p[0], p[1], rm, rv, training, 0.1, 1e-5, False)

    def ref(a, p):
        y = F.relu(F.batch_norm(a[0], rrm, rrv, p[0], p[1], training, 0.1, 1e-5))
        # the pool sees the activation as stored (rounded to the storage dtype): ties after rounding go to the first maximum
        ys = y + (y.to(dtype).double() - y).detach()
        return y, F.max_pool2d(ys, 2)

    run_case(mine, ref, [x], [g, b], dtype)
    if training:
        t = 1e-5 if dtype == torch.float32 else 1e-2
        assert rel(rm, rrm) < t and rel(rv, rrv) < t
    else:
        assert torch.equal(rm, rm0) and torch.equal(rv, rv0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_relu_gelu(dtype):
    from eel_unet_b200 import ops

    x = torch.randn(2, 16, 8, 12, device=DEV)
    run_case(lambda a, p: ops.MaxPool2.apply(a[0]), lambda a, p: F.max_pool2d(a[0], 2), [x], [], dtype)
    run_case(lambda a, p: ops.Relu.apply(a[0]), lambda a, p: F.relu(a[0]), [x], [], dtype)
    run_case(lambda a, p: ops.Gelu.apply(a[0]), lambda a, p: F.gelu(a[0]), [x], [], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 64, 8, 8, 256), (1, 16, 5, 6, 32), (1, 64, 4, 4, 48)])
def test_linear_gelu_bias_gradient_from_gelu_backward(dtype, shape):
    """mlp[0] -> GELU (reference models/EELUnet.py:107-108): the fused GELU backward hands the Linear its bias gradient"""
    from eel_unet_b200 import ops

    n, k, h, w, nout = shape
    x = torch.randn(n, k, h, w, device=DEV)
    wt = torch.randn(nout, k, device=DEV) / math.sqrt(k)
    b = torch.randn(nout, device=DEV)
    run_case(lambda a, p: ops.Gelu.apply(ops.Linear.apply(a[0], p[0], p[1], False), True),
             lambda a, p: F.gelu(F.conv2d(a[0], p[0][:, :, None, None], p[1])), [x], [wt, b], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_ties_go_to_first_max(dtype):
    from eel_unet_b200 import ops

    x = torch.zeros(1, 8, 4, 4, device=DEV)   # all ties (post-ReLU zeros)
    a = nhwc(x).to(dtype).requires_grad_(True)
    y = ops.MaxPool2.apply(a)
    y.backward(torch.ones_like(y))
    r = x.clone().requires_grad_(True)
    F.max_pool2d(r, 2).backward(torch.ones(1, 8, 2, 2, device=DEV))
    assert torch.equal(nchw(a.grad.float()), r.grad)


@pytest.mark.parametrize("dtype", DTYPES)
def test_add_interleave(dtype):
    from eel_unet_b200 import ops

    xs = [torch.randn(2, 16, 6, 6, device=DEV) for _ in range(3)]

    def ref(a, p):
        s = a[0] + a[1]
        n, c, h, w = s.shape
        return torch.stack([s, a[2]], dim=2).reshape(n, 2 * c, h, w)

    run_case(lambda a, p: ops.AddInterleave.apply(a[0], a[1], a[2]), ref, xs, [], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("edge_bn", [False, True])
@pytest.mark.parametrize("shape", [(2, 64, 12, 20), (1, 128, 16, 8), (3, 16, 6, 10), (1, 512, 4, 6)])
def test_bn_add_interleave(dtype, edge_bn, shape):
    """decoder skip bridge (reference models/EELUnet.py:422-426) with the upconv block's BatchNorm (:365,373) applied inside it.
    Backward: the de-interleaving pass also accumulates that BatchNorm's backward sums (eel_add_interleave_bwd_bnsums) and --
    edge_bn -- those of the BatchNorm + ReLU that produced the edge feature when the bridge is its only consumer (bf16 mode);
    both BatchNorm backwards are then single apply passes"""
    from eel_unet_b200 import _lib, ops

    n, c, h, w = shape
    z = torch.randn(n, c, h, w, device=DEV) * 1.5 + 0.3
    zb = torch.randn(n, c, h, w, device=DEV) * 0.7 - 0.2
    e = torch.randn(n, c, h, w, device=DEV)
    params = [torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.5]
    if edge_bn:
        params += [torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.5]
    rm0, rv0 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    rm1, rv1 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)

    def mine(a, p):
        feat = ops.BNAct.apply(a[1], p[2], p[3], rm1, rv1, True, True, 0.1, 1e-5, False, True) if edge_bn else a[1]
        return ops.BNAddInterleave.apply(a[0], p[0], p[1], rm0, rv0, True, 0.1, 1e-5, feat, a[2], False)

    def ref(a, p):
        up = F.batch_norm(a[0], None, None, p[0], p[1], True, 0.1, 1e-5)
        feat = a[1]
        if edge_bn:
            feat = F.relu(F.batch_norm(a[1], None, None, p[2], p[3], True, 0.1, 1e-5))
            feat = feat + (feat.to(dtype).double() - feat).detach()      # the bridge reads the activation as stored
        s = up + feat
        return torch.stack([s, a[2]], dim=2).reshape(n, 2 * c, h, w)

    rec = []
    _lib.set_profiler(rec)
    try:
        run_case(mine, ref, [z, zb, e], params, dtype, atol_scale=2.0)
    finally:
        _lib.set_profiler(None)
    names = [r[0] for r in rec]
    assert "eel_add_interleave_bwd_bnsums" in names and names.count("eel_bn_act_bwd_apply") >= 1
    if dtype == torch.bfloat16:
        assert "eel_bn_act_bwd" not in names and names.count("eel_bn_act_bwd_apply") == (2 if edge_bn else 1)


@pytest.mark.parametrize("shape", [(2, 64, 40, 24, 64), (1, 128, 36, 20, 128), (1, 256, 16, 16, 256), (2, 512, 8, 8, 512),
                                   (1, 64, 16, 16, 128)])
def test_bridge_conv_backward_in_two_halves(shape):
    """skip bridge + the decoder block's first conv3x3 as one node (ops.BridgeConv3x3; reference models/EELUnet.py:365/373,
    :422-426, :132-141, :338).  Forward: the conv reads BatchNorm(z) + b and e through two tensor maps on a de-interleaved
    operand -- no interleaved tensor.  Backward at >= 128 channels: the data gradient stores d(BatchNorm(z) + b) and d(e) as two
    tensors with the BatchNorm's backward sums from its epilogue (eel_tc_conv3x3_dgrad_split); at 64 channels the interleaved
    gradient + the de-interleaving pass with the sums.  Weight gradient: two half launches interleaved into the reference layout"""
    from eel_unet_b200 import _lib, ops

    n, c, h, w, cout = shape
    dtype = torch.bfloat16
    z = torch.randn(n, c, h, w, device=DEV) * 1.5 + 0.3
    b = torch.randn(n, c, h, w, device=DEV) * 0.7
    e = torch.randn(n, c, h, w, device=DEV)
    g = torch.rand(c, device=DEV) + 0.5
    bt = torch.randn(c, device=DEV) * 0.5
    wt = torch.randn(cout, 2 * c, 3, 3, device=DEV) / math.sqrt(18 * c)
    bias = torch.randn(cout, device=DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    assert ops.BridgeConv3x3.supported(nhwc(z).to(dtype), wt)

    def mine(a, p):
        return ops.BridgeConv3x3.apply(a[0], p[0], p[1], rm, rv, True, 0.1, 1e-5, a[1], a[2], False, p[2], p[3])

    def ref(a, p):
        s = F.batch_norm(a[0], None, None, p[0], p[1], True, 0.1, 1e-5) + a[1]
        x = torch.stack([s, a[2]], dim=2).reshape(n, 2 * c, h, w)
        x = x + (x.to(dtype).double() - x).detach()               # the conv reads the interleaved tensor as stored
        return F.conv2d(x, p[2], p[3], padding=1)

    rec = []
    _lib.set_profiler(rec)
    try:
        run_case(mine, ref, [z, b, e], [g, bt, wt, bias], dtype, atol_scale=2.0)
    finally:
        _lib.set_profiler(None)
    names = [r[0] for r in rec]
    assert "eel_tc_conv3x3_2src" in names and "eel_add_interleave_fwd" not in names          # the interleaved tensor is never built
    assert ("eel_tc_conv3x3_dgrad_split" in names) == (c >= 128) and ("eel_add_interleave_bwd_bnsums" in names) == (c < 128)
    assert "eel_bn_act_bwd" not in names and names.count("eel_bn_act_bwd_apply") == 1 and names.count("eel_tc_conv3x3_wgrad") == 2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("c", [64, 128, 1024])
def test_pgr(dtype, c):
    from eel_unet_b200 import ops

    x = torch.randn(2, c, 6, 10, device=DEV)
    wt = torch.randn(1, c, 1, 1, device=DEV) / math.sqrt(c)
    b = torch.randn(1, device=DEV)

    def ref(a, p):
        s = torch.sigmoid(F.conv2d(a[0], p[0], p[1]))
        return a[0] + a[0] * s, s

    run_case(lambda a, p: ops.PGR.apply(a[0], p[0], p[1]), ref, [x], [wt, b], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("o", [1, 3])
def test_head(dtype, o):
    from eel_unet_b200 import ops

    x = torch.randn(2, 64, 10, 6, device=DEV) * 1.5 + 0.3
    lw = torch.rand(64, device=DEV) + 0.5
    lb = torch.randn(64, device=DEV) * 0.1
    wt = torch.randn(o, 64, 1, 1, device=DEV) / 8
    b = torch.randn(o, device=DEV)

    def ref(a, p):
        u = a[0].mean(1, keepdim=True)
        s = (a[0] - u).pow(2).mean(1, keepdim=True)
        y = (a[0] - u) / torch.sqrt(s + 1e-6)
        y = p[0][None, :, None, None] * y + p[1][None, :, None, None]
        return torch.sigmoid(F.conv2d(y, p[2], p[3]))

    run_case(lambda a, p: ops.Head.apply(a[0], p[0], p[1], p[2], p[3]), ref, [x], [lw, lb, wt, b], dtype, out_nhwc=False)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(3, 64, 6, 10), (2, 64, 32, 32), (5, 64, 16, 16), (300, 64, 2, 2), (2, 128, 8, 24)])
def test_se(dtype, shape):
    """ChannelAttention (reference models/EELUnet.py:57-80).  One cooperative launch per direction (k blocks per image, grid
    barrier, the tiny MLP repeated per block): shapes with 1, 16 and 4 blocks per image, one with more images than the
    cooperative grid allows (multi-launch path), and a wider token"""
    from eel_unet_b200 import ops

    n, c, h, w = shape
    x = torch.randn(n, c, h, w, device=DEV)
    w1 = torch.randn(4, c, 1, 1, device=DEV) / 8
    b1 = torch.randn(4, device=DEV) * 0.5
    w2 = torch.randn(c, 4, 1, 1, device=DEV) / 2
    b2 = torch.randn(c, device=DEV) * 0.5

    def ref(a, p):
        g = a[0].mean(dim=(2, 3), keepdim=True)
        g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(g, p[0], p[1])), p[2], p[3]))
        return a[0] * g

    run_case(lambda a, p: ops.SE.apply(a[0], p[0], p[1], p[2], p[3]), ref, [x], [w1, b1, w2, b2], dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_to_patch_bias_gradient_from_se_backward(dtype):
    """to_patch -> ChannelAttention (reference models/EELUnet.py:118-119): the squeeze-excite backward hands the 1x1 conv in
    front of it its bias gradient (column sums of dt, from the per-image sums it reduces anyway)"""
    from eel_unet_b200 import ops

    x = torch.randn(3, 128, 8, 16, device=DEV)
    wt = torch.randn(64, 128, 1, 1, device=DEV) / math.sqrt(128)
    bt = torch.randn(64, device=DEV)
    w1 = torch.randn(4, 64, 1, 1, device=DEV) / 8
    b1 = torch.randn(4, device=DEV) * 0.5
    w2 = torch.randn(64, 4, 1, 1, device=DEV) / 2
    b2 = torch.randn(64, device=DEV) * 0.5

    def ref(a, p):
        t = F.conv2d(a[0], p[0], p[1])
        g = t.mean(dim=(2, 3), keepdim=True)
        g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(g, p[2], p[3])), p[4], p[5]))
        return t * g

    run_case(lambda a, p: ops.SE.apply(ops.Linear.apply(a[0], p[0], p[1], False), p[2], p[3], p[4], p[5], True), ref,
             [x], [wt, bt, w1, b1, w2, b2], dtype)


def ref_hft(x, mask_range=20):
    h, w = x.shape[-2:]
    crow, ccol = h // 2, w // 2
    r = min(mask_range, crow, ccol)
    mask = torch.ones(h, w, dtype=x.dtype, device=x.device)
    mask[crow - r:crow + r, ccol - r:ccol + r] = 0
    d = torch.fft.fftshift(torch.fft.fft2(x), dim=(-2, -1)) * mask
    return torch.abs(torch.fft.ifft2(torch.fft.ifftshift(d, dim=(-2, -1))))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 8, 64, 64), (1, 16, 48, 80), (1, 8, 16, 16), (2, 64, 128, 128), (1, 128, 64, 256),
                                   (1, 64, 48, 192), (1, 64, 512, 512), (1, 128, 256, 128), (1, 64, 1024, 1024),
                                   (3, 64, 256, 256), (2, 128, 128, 128), (1, 128, 256, 256), (2, 64, 128, 256)])
def test_hft(dtype, shape):
    from eel_unet_b200 import ops

    x = torch.randn(*shape, device=DEV)
    if shape[-1] == 16:
        # fully masked spectrum (r = H/2): output is exactly the rounding residue of x - Px; compare absolutely
        a = nhwc(x).to(dtype)
        y = ops.HFT.apply(a, 20)
        assert y.float().abs().max().item() < (1e-4 if dtype == torch.float32 else 5e-2)
        return
    run_case(lambda a, p: ops.HFT.apply(a[0], 20), lambda a, p: ref_hft(a[0]), [x], [], dtype,
             atol_scale=5.0 if dtype == torch.float32 else 1.5, all_nhwc=True)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 64, 128, 128), (1, 64, 512, 512), (1, 16, 48, 80), (1, 128, 256, 128)])
def test_hft_without_a_backward_keeps_no_phase(dtype, shape):
    """inference (the input needs no gradient): eel_hft_fwd is called with phase == NULL on every path (16-bit code kernels,
    1024-wide tensor-core step 4, SIMT) and must return the same y"""
    from eel_unet_b200 import _lib, ops

    a = nhwc(torch.randn(*shape, device=DEV)).to(dtype)
    with_grad = ops.HFT.apply(a.clone().requires_grad_(True), 20)
    rec = []
    _lib.set_profiler(rec)
    try:
        with torch.no_grad():
            plain = ops.HFT.apply(a, 20)
    finally:
        _lib.set_profiler(None)
    assert [r[1][2] for r in rec if r[0] == "eel_hft_fwd"] == [None]         # the phase argument
    assert torch.equal(plain, with_grad.detach())


@pytest.mark.parametrize("shape", [(2, 64, 256, 256), (1, 128, 128, 128)])
def test_hft_phase_code(shape):
    """bf16 training shapes keep the unit phase z/|z| as one 16-bit code per element (csrc/hft_tc.cu): bit 15 = the smaller
    component is re, bit 14 = the larger one is negative, bits 13..0 = 8192 + the smaller one in fixed point (step
    sqrt(1/2)/8191), field 0 = zero vector.  Decoded on the host it must match the phase of the fp64 FFT formulation; |z| likewise."""
    from eel_unet_b200 import _lib
    from eel_unet_b200._lib import call, ptr, stream, workspace

    torch.manual_seed(2)
    n, c, h, w = shape
    x = torch.randn(n, c, h, w, device=DEV)
    x[0, :, :, : w // 2] = 0                                   # large flat regions: small |z|
    a = nhwc(x).bfloat16()
    assert _lib.lib.eel_hft_phase_elems(n, h, w, c, 20, _lib.EEL_BF16) == n * h * w * c
    y = torch.empty_like(a)
    code = torch.empty((n, h, w, c), dtype=torch.int16, device=DEV)
    nb = _lib.lib.eel_hft_workspace_bytes(n, h, w, c, 20)
    ws = workspace(nb, a.device, slot=1)
    call("eel_hft_fwd", ptr(a), ptr(y), ptr(code), n, h, w, c, 20, ptr(ws), nb, _lib.EEL_BF16, stream())
    xd = nchw(a.double())
    crow, ccol, r = h // 2, w // 2, 20
    mask = torch.ones(h, w, dtype=torch.float64, device=DEV)
    mask[crow - r:crow + r, ccol - r:ccol + r] = 0
    z = torch.fft.ifft2(torch.fft.ifftshift(torch.fft.fftshift(torch.fft.fft2(xd), dim=(-2, -1)) * mask, dim=(-2, -1)))
    z = nhwc(z)
    assert rel(y.float(), z.abs()) < 1e-2
    u = code.to(torch.int32) & 0xFFFF
    q = (u & 0x3FFF) - 8192                                     # biased 14-bit field
    small = q.double() * (0.5 ** 0.5 / 8191)
    large = torch.sqrt((1 - small * small).clamp_min(0)) * torch.where((u & 0x4000) != 0, -1.0, 1.0)
    sel = (u & 0x8000) != 0
    re, im = torch.where(sel, small, large), torch.where(sel, large, small)
    zero = (u & 0x3FFF) == 0
    re, im = torch.where(zero, 0.0, re), torch.where(zero, 0.0, im)
    ok = z.abs() > 0.05 * z.abs().mean()                        # where the phase is well defined against bf16 input rounding
    ph = z / z.abs().clamp_min(1e-300)
    err = torch.sqrt((re - ph.real) ** 2 + (im - ph.imag) ** 2)[ok]
    assert err.max().item() < 0.3 and err.mean().item() < 1e-2, (err.max().item(), err.mean().item())
    assert ((re * re + im * im)[~zero] - 1).abs().max().item() < 1e-3


@pytest.mark.parametrize("soft", [False, True])
def test_edge_loss(soft):
    from eel_unet_b200 import edge_BceDiceLoss

    torch.manual_seed(3)
    n, h, w = 3, 64, 96
    t = (torch.rand(n, 1, h, w, device=DEV) > 0.7).float()
    if soft:
        t = F.avg_pool2d(t, 3, 1, 1)
    preds = [torch.rand(n, 1, h // s, w // s, device=DEV).clamp(1e-4, 1 - 1e-4).requires_grad_(True) for s in (1, 16, 8, 4, 2, 1)]
    loss = edge_BceDiceLoss(1, 1)(preds[1:], preds[0], t)

    def bd(p, tt):
        p, tt = p.double(), tt.double()
        nn_ = p.shape[0]
        p_, t_ = p.reshape(nn_, -1), tt.reshape(nn_, -1)
        bce = F.binary_cross_entropy(p_, t_)
        dice = 1 - ((2 * (p_ * t_).sum(1) + 1) / (p_.sum(1) + t_.sum(1) + 1)).sum() / nn_
        return bce + dice

    rp = [p.detach().double().requires_grad_(True) for p in preds]
    rl = bd(rp[0], t)
    for k, (s, wk) in enumerate(zip((16, 8, 4, 2, 1), (0.1, 0.2, 0.3, 0.4, 0.5))):
        tt = F.max_pool2d(t, s, s) if s > 1 else t
        rl = rl + wk * bd(rp[k + 1], tt)
    assert abs(loss.item() - rl.item()) < 1e-5 * abs(rl.item())
    (loss * 1.7).backward()
    (rl * 1.7).backward()
    for p, r in zip(preds, rp):
        assert rel(p.grad, r.grad) < 1e-5


def test_edge_loss_saturated_probabilities():
    """p exactly 0/1: log clamps at -100 (nn.BCELoss) and the gradient divides by max(p(1-p), 1e-12)."""
    from eel_unet_b200 import edge_BceDiceLoss

    n, h, w = 1, 16, 16
    t = torch.zeros(n, 1, h, w, device=DEV)
    t[..., :8] = 1
    preds = [torch.full((n, 1, h // s, w // s), 0.5, device=DEV) for s in (1, 16, 8, 4, 2, 1)]
    preds[0] = torch.where(torch.rand(n, 1, h, w, device=DEV) > 0.5, torch.ones(()).to(DEV), torch.zeros(()).to(DEV))
    preds = [p.requires_grad_(True) for p in preds]
    loss = edge_BceDiceLoss(1, 1)(preds[1:], preds[0], t)
    rp = [p.detach().clone().requires_grad_(True) for p in preds]
    crit_b = torch.nn.BCELoss()

    def bd(p, tt):
        nn_ = p.shape[0]
        p_, t_ = p.view(nn_, -1), tt.view(nn_, -1)
        return crit_b(p_, t_) + 1 - ((2 * (p_ * t_).sum(1) + 1) / (p_.sum(1) + t_.sum(1) + 1)).sum() / nn_

    rl = bd(rp[0], t)
    for k, (s, wk) in enumerate(zip((16, 8, 4, 2, 1), (0.1, 0.2, 0.3, 0.4, 0.5))):
        rl = rl + wk * bd(rp[k + 1], F.max_pool2d(t, s, s) if s > 1 else t)
    assert abs(loss.item() - rl.item()) < 1e-4 * abs(rl.item())
    loss.backward()
    rl.backward()
    assert rel(preds[0].grad, rp[0].grad) < 1e-5


def test_adam_matches_torch():
    from eel_unet_b200 import _lib

    torch.manual_seed(0)
    n = 10007
    p = torch.randn(n, device=DEV)
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=1e-3, weight_decay=1e-5)
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    for step in range(1, 4):
        g = torch.randn(n, device=DEV)
        ref_p.grad = g.clone()
        opt.step()
        _lib.call("eel_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 1e-5,
                  step, _lib.stream())
    assert rel(p, ref_p.detach()) < 1e-6


def test_fused_adam_is_a_torch_optimizer_with_steplr_and_checkpoints():
    """SURVEY 8f-1 / reference train.py:312-315,118: ``optim.Adam(params, lr, weight_decay=1e-5)`` + ``StepLR(optimizer, 30, 0.5)``.
    FusedAdam must be schedulable by the stock StepLR, follow torch.optim.Adam step for step through the decays, skip
    parameters without a gradient like torch does, and exchange optimizer checkpoints with torch.optim.Adam in both directions."""
    from torch.optim.lr_scheduler import StepLR

    from eel_unet_b200.parallel import FusedAdam, GradBuckets

    def net():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(9, 17), torch.nn.Tanh(), torch.nn.Linear(17, 6), torch.nn.Tanh(),
                                   torch.nn.Linear(6, 2)).to(DEV)

    ref, mine = net(), net()
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-5)
    rsch = StepLR(ropt, 3, 0.5)
    gb = GradBuckets(list(mine.parameters()))
    try:
        opt = FusedAdam(gb, lr=1e-2, weight_decay=1e-5)
        assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 1
        sch = StepLR(opt, 3, 0.5)
        torch.manual_seed(11)
        saved = None
        for it in range(10):
            x = torch.randn(5, 9, device=DEV)
            skip_last = it in (4, 5)                 # the last layer gets no gradient in these steps
            for model, o, is_ref in ((ref, ropt, True), (mine, opt, False)):
                o.zero_grad()
                h = model[:4](x) if skip_last else model(x)
                h.pow(2).mean().backward()
                if not is_ref:
                    gb.finish()
                o.step()
            rsch.step()
            sch.step()
            assert abs(opt.param_groups[0]["lr"] - ropt.param_groups[0]["lr"]) < 1e-15
            for a, b in zip(mine.parameters(), ref.parameters()):
                assert rel(a.detach(), b.detach()) < 2e-6, it
            if it == 6:
                import copy
                # (torch's state_dict() hands out the LIVE moment tensors, which later steps update in place)
                saved = (opt.state_dict(), copy.deepcopy(ropt.state_dict()), [p.detach().clone() for p in mine.parameters()])
        assert abs(opt.param_groups[0]["lr"] - 1e-2 * 0.5 ** 3) < 1e-12

        # checkpoint round trips: ours -> torch.optim.Adam, torch.optim.Adam -> ours, ours -> ours; one more step must agree
        mine_sd, ref_sd, params = saved
        assert set(mine_sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
        for k in range(6):                           # the skipped layer's step count is 2 behind in both
            assert rel(mine_sd["state"][k]["exp_avg"], ref_sd["state"][k]["exp_avg"]) < 1e-5
            assert float(mine_sd["state"][k]["step"]) == float(ref_sd["state"][k]["step"]) == (7.0 if k < 4 else 5.0)
        x = torch.randn(5, 9, device=DEV)
        outs = []
        for kind in ("torch<-ours", "ours<-torch", "ours<-ours"):
            m2 = net()
            with torch.no_grad():
                for a, b in zip(m2.parameters(), params):
                    a.copy_(b)
            if kind == "torch<-ours":
                o2 = torch.optim.Adam(m2.parameters(), lr=1.0)
                o2.load_state_dict(copy.deepcopy(mine_sd))      # (torch adopts the given moment tensors and steps them in place)
                fin = lambda: None
            else:
                gb2 = GradBuckets(list(m2.parameters()))
                o2 = FusedAdam(gb2, lr=1.0)
                o2.load_state_dict(ref_sd if kind == "ours<-torch" else mine_sd)
                fin = gb2.finish
            assert abs(o2.param_groups[0]["lr"] - mine_sd["param_groups"][0]["lr"]) < 1e-15
            o2.zero_grad()
            m2(x).pow(2).mean().backward()
            fin()
            o2.step()
            outs.append([p.detach().clone() for p in m2.parameters()])
            if kind != "torch<-ours":
                gb2.remove()
        for other in outs[1:]:
            for k, (a, b) in enumerate(zip(other, outs[0])):
                assert rel(a, b) < 2e-6, k
    finally:
        gb.remove()


@pytest.mark.parametrize("case", ["plain", "hook", "second_consumer"])
def test_side_channels_between_ops_are_safe_against_hooks_and_gradient_accumulation(case):
    """Linear -> training-mode BatchNorm(+ReLU) hand each other two things on the connecting tensor z: the BatchNorm sums from
    the GEMM epilogue (forward) and the bias gradient from the BatchNorm backward (backward).  Both ride on the tensor OBJECT
    with its version counter (ops._attach / ops._take), so a user hook that replaces the gradient, or a second consumer whose
    gradient autograd accumulates IN PLACE into the BatchNorm's dz buffer, must make the consumer fall back to computing the
    quantity itself -- never use a stale one."""
    from eel_unet_b200 import ops

    torch.manual_seed(0)
    N, H, W, K, C = 4, 16, 16, 64, 128
    x = torch.randn(N, H, W, K, device=DEV).bfloat16()
    w = (torch.randn(C, K, device=DEV) * 0.2).requires_grad_(True)
    b = torch.randn(C, device=DEV).requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    bet = (0.1 * torch.randn(C, device=DEV)).requires_grad_(True)
    r = torch.randn(N, H, W, C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)

    ops.expect_bn(True)
    try:
        z = ops.Linear.apply(x, w, b, False)
    finally:
        ops.expect_bn(False)
    assert hasattr(z, "_eel_bn_sums")
    if case == "hook":
        z.register_hook(lambda g: g * 2.0)
    y = ops.BNAct.apply(z, gam, bet, rm, rv, True, True, 0.1, 1e-5, True)
    assert not hasattr(z, "_eel_bn_sums")                       # consumed exactly once
    loss = (y.float() * r).sum()
    if case == "second_consumer":
        loss = loss + 0.5 * z.float().sum()
    loss.backward()

    xr = x.double()
    wr, br = w.detach().double().bfloat16().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    gr, ber = gam.detach().double().requires_grad_(True), bet.detach().double().requires_grad_(True)
    zr = F.linear(xr, wr, br)
    if case == "hook":
        zr.register_hook(lambda g: g * 2.0)
    yr = F.relu(F.batch_norm(zr.permute(0, 3, 1, 2), None, None, gr, ber, True, 0.1, 1e-5)).permute(0, 2, 3, 1)
    lr = (yr * r.double()).sum()
    if case == "second_consumer":
        lr = lr + 0.5 * zr.sum()
    lr.backward()
    assert rel(y, yr) < 2e-2
    assert rel(w.grad, wr.grad) < 3e-2 and rel(gam.grad, gr.grad) < 2e-2 and rel(bet.grad, ber.grad) < 2e-2
    if case == "second_consumer":
        # the bias gradient is 0.5 * pixels here; a stale hand-over would have delivered the BatchNorm's ~0
        assert rel(b.grad, br.grad) < 1e-2 and abs(br.grad.mean().item() - 0.5 * N * H * W) < 1e-6
    else:
        assert b.grad.abs().max().item() < 1e-2 * w.grad.abs().max().item() * K      # analytically zero


@pytest.mark.parametrize("dims,perm", [
    ((128, 64, 3, 3), (2, 3, 0, 1)), ((128, 64, 3, 3), (2, 3, 1, 0)), ((64, 96, 2, 2), (0, 2, 3, 1)), ((64, 96, 2, 2), (2, 3, 1, 0)),
    ((3, 3, 64, 160), (3, 2, 0, 1)), ((96, 2, 2, 64), (0, 3, 1, 2)), ((256, 64, 1, 1), (2, 3, 1, 0)), ((256, 64, 1, 1), (2, 3, 0, 1)),
    ((64, 3, 3, 3), (2, 3, 0, 1)), ((5, 7, 3, 2), (3, 1, 0, 2))])
def test_permute4_tiled_and_generic_paths(dims, perm):
    """eel_permute4: the weight / weight-gradient layout changes go through shared-memory tiles (csrc/elementwise.cu
    tile3_*), everything else through the generic kernel; both against torch.permute, fp32 -> fp32 and fp32 -> bf16."""
    from eel_unet_b200 import ops

    torch.manual_seed(1)
    w = torch.randn(*dims, device=DEV)
    ref = w.permute(*perm).contiguous()
    assert torch.equal(ops._pack(w, perm, torch.float32), ref)
    assert torch.equal(ops._pack(w, perm, torch.bfloat16), ref.bfloat16())


@pytest.mark.parametrize("shape", [(2, 32, 32, 256), (1, 16, 32, 512), (1, 16, 16, 1024), (3, 64, 64, 256)])
@pytest.mark.parametrize("with_bn", [False, True])
def test_fused_token_mlp_is_bit_identical_to_the_three_launches(shape, with_bn):
    """ops.MlpChain (csrc/capmlp_tc.cu: mlp[0] -> GELU -> composed mlp[2] + to_space in one tcgen05 kernel, reference
    models/EELUnet.py:107-111,121-122) against Linear -> Gelu -> ComposedLinear: outputs, BatchNorm statistics and every
    gradient must agree exactly (same MMA order, same rounding points, same backward kernels); and against the fp64 formulation."""
    from eel_unet_b200 import ops

    torch.manual_seed(4)
    N, H, W, C = shape
    u = torch.randn(N, H, W, 64, device=DEV).bfloat16()
    w0 = (torch.randn(256, 64, device=DEV) * 0.15)
    b0 = torch.randn(256, device=DEV) * 0.1
    w1 = (torch.randn(C, 256, device=DEV) * 0.08)            # mlp[2]
    b1 = torch.randn(C, device=DEV) * 0.1
    w2 = (torch.randn(C, C, 1, 1, device=DEV) * (1.0 / C ** 0.5))   # to_space
    b2 = torch.randn(C, device=DEV) * 0.1
    r = torch.randn(N, H, W, C, device=DEV).bfloat16()
    assert ops.mlp_chain_supported(u, w0, C)

    def run(fused):
        ps = [t.clone().requires_grad_(True) for t in (w0, b0, w1, b1, w2, b2)]
        uu = u.clone().requires_grad_(True)
        a = None if fused else ops.Gelu.apply(ops.Linear.apply(uu, ps[0], ps[1], False), True)
        ops.expect_bn(with_bn)               # (only the LAST layer of the chain feeds the BatchNorm)
        try:
            z = ops.MlpChain.apply(uu, *ps) if fused else ops.ComposedLinear.apply(a, *ps[2:])
        finally:
            ops.expect_bn(False)
        sums = ops._take(z, "_eel_bn_sums")
        z.backward(r)
        return z.detach(), (None if sums is None else sums[0]), uu.grad, [p.grad for p in ps]

    z1, s1, du1, g1 = run(True)
    z0, s0, du0, g0 = run(False)
    assert torch.equal(z1, z0) and torch.equal(du1, du0)
    assert (s1 is None) == (not with_bn) and (s1 is None or rel(s1, s0) < 1e-5)       # (fp32 atomics: order differs)
    for a, b in zip(g1, g0):
        assert rel(a, b) < 1e-5
    if not with_bn:
        ud = u.double()
        hd = F.linear(ud, w0.double().bfloat16().double(), b0.double())
        zd = F.linear(F.linear(F.gelu(hd), w1.double(), b1.double()), w2.double().view(C, C), b2.double())
        assert rel(z1, zd) < 2e-2
