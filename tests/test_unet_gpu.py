"""GPU: drop-in Unet (reference models/Unet.py, BASELINE config 5 comparator) against the CPU oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_forward_backward(precision):
    from eel_unet_b200 import Unet
    from oracle import eelunet_torch as O
    from oracle import synth

    torch.manual_seed(0)
    model = Unet(3, 1)
    assert model.name == "unet" and len(model.state_dict()) == 46
    xs, ys, _ = synth.batch(2, 64, 96, 1)
    x, y = torch.from_numpy(xs), torch.from_numpy(ys)

    def run(dtype):
        sd = {k: v.detach().to(dtype).requires_grad_(True) for k, v in model.state_dict().items()}
        out = O.unet_forward(sd, x.to(dtype))
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y.to(dtype))
        loss.backward()
        return out.detach(), {k: v.grad for k, v in sd.items()}

    o64, g64 = run(torch.float64)
    o32, g32 = run(torch.float32)
    m = model.cuda().set_precision(precision)
    out = m(x.cuda())
    assert out.dtype == torch.float32 and tuple(out.shape) == (2, 1, 64, 96)
    torch.nn.functional.binary_cross_entropy_with_logits(out, y.cuda()).backward()
    tol_f, tol_g = (1e-4, 1e-3) if precision == "fp32" else (2e-2, 5e-2)
    assert rel(out, o64) <= max(tol_f, 3 * rel(o32, o64))
    for name, p in m.named_parameters():
        assert rel(p.grad, g64[name]) <= max(tol_g, 5 * rel(g32[name], g64[name])), name


def test_unet_matches_reference_golden():
    """drop-in Unet against the fp64 outputs of the reference's models/Unet.py itself (tests/golden/unet_2x64x96.npz,
    made by tests/golden/make_golden_r2.py): logits 1e-4 / gradients 1e-3 in fp32 mode, 2e-2 logits in bf16 mode."""
    import os

    import numpy as np

    from eel_unet_b200 import Unet
    from oracle import synth

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "unet_2x64x96.npz"))
    xs, ys, _ = synth.batch(2, 64, 96, 1)
    x, y = torch.from_numpy(xs).cuda(), torch.from_numpy(ys).cuda()
    ref = torch.from_numpy(g["logits"])
    for precision in ("fp32", "bf16"):
        torch.manual_seed(0)
        m = Unet(3, 1, precision=precision).cuda()
        out = m(x)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        assert rel(out, ref) <= (1e-4 if precision == "fp32" else 2e-2), (precision, rel(out, ref))
        assert abs(loss.item() - float(g["loss"])) <= (1e-5 if precision == "fp32" else 2e-3)
        if precision == "fp32":
            grads = dict(m.named_parameters())
            for n, gn in zip([str(s) for s in g["grad_names"]], g["grad_norm"]):
                assert abs(grads[n].grad.norm().item() - gn) <= 1e-3 * gn, n
            for key in g.files:
                if key.startswith("grad:"):
                    assert rel(grads[key[5:]].grad, torch.from_numpy(g[key])) <= 1e-3, key
