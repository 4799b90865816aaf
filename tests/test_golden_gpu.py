"""GPU: the CUDA path against the committed fixtures the REAL reference produced (tests/golden/)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def test_train_step_vs_reference_golden_fp32():
    from eel_unet_b200 import EELUnet, edge_BceDiceLoss
    from oracle import synth

    g = np.load(os.path.join(GOLD, "eelunet_train_2x128.npz"))
    torch.manual_seed(0)
    model = EELUnet(3, 1).cuda().train()
    xs, ys, _ = synth.batch(2, 128, 128, 0)
    seg, edges = model(torch.from_numpy(xs).cuda())
    loss = edge_BceDiceLoss(1, 1)(edges, seg, torch.from_numpy(ys).cuda())
    loss.backward()
    # bar: the north_star tolerance, or 3x the reference's own fp32-vs-fp64 deviation stored in the fixture
    assert rel(seg, g["seg"]) <= max(1e-4, 3 * float(g["seg_f32_relerr"]))
    for k, e in enumerate(edges):
        assert rel(e, g["edge%d" % (5 - k)]) <= max(1e-4, 3 * float(g["edge%d_f32_relerr" % (5 - k)]))
    assert abs(loss.item() - float(g["loss"])) <= max(1e-4 * float(g["loss"]), 3 * float(g["loss_f32_abserr"]))
    names = [str(n) for n in g["grad_names"]]
    params = dict(model.named_parameters())
    gmax = float(g["grad_norm"].max())
    ratios = []
    for n, gn, ferr in zip(names, g["grad_norm"], g["grad_f32_relerr"]):
        if gn < 1e-6 * gmax:
            continue
        mine = abs(params[n].grad.norm().item() - gn) / gn          # norm agreement (full tensors are not stored)
        ratios.append(mine / max(float(ferr), 1e-7))
        if ferr < 0.05:
            assert mine <= max(1e-3, 5 * float(ferr)), n
    assert float(np.median(ratios)) <= 2.0
    for key in g.files:
        if key.startswith("grad:") and np.linalg.norm(g[key]) > 1e-6 * gmax:
            n = key[5:]
            ferr = float(g["grad_f32_relerr"][names.index(n)])
            if ferr < 0.05:
                assert rel(params[n].grad, g[key]) <= max(1e-3, 5 * ferr), n
        if key.startswith("stat:"):
            assert rel(model.state_dict()[key[5:]], g[key]) <= 1e-4, key


def test_eval_forward_vs_reference_golden():
    from eel_unet_b200 import EELUnet
    from oracle import synth

    g = np.load(os.path.join(GOLD, "eelunet_eval_2x128.npz"))
    torch.manual_seed(0)
    model = EELUnet(3, 1).cuda().eval()
    xs, _, _ = synth.batch(2, 128, 128, 0)
    with torch.no_grad():
        seg, edges = model(torch.from_numpy(xs).cuda())
        assert rel(seg, g["seg"]) <= 1e-4                                   # fp32 mode: north_star 1e-4
        for k, e in enumerate(edges):
            assert rel(e, g["edge%d" % (5 - k)]) <= 1e-4
        segb, _ = model.set_precision("bf16")(torch.from_numpy(xs).cuda())
        assert rel(segb, g["seg"]) <= 2e-2                                  # bf16 mode: north_star 2e-2


def test_loss_vs_reference_golden():
    from eel_unet_b200 import edge_BceDiceLoss

    g = np.load(os.path.join(GOLD, "loss_cases.npz"))
    preds = [torch.from_numpy(g["pred%d" % i]).float().cuda().requires_grad_(True) for i in range(6)]
    loss = edge_BceDiceLoss(1, 1)(preds[1:], preds[0], torch.from_numpy(g["target"]).float().cuda())
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 2e-6 * float(g["loss"])
    for i, p in enumerate(preds):
        # fixture probabilities are fp64 and reach 1 - 1e-4: casting them to fp32 perturbs 1 - p by 6e-4 relative,
        # which the BCE gradient 1 / (p (1 - p)) passes straight through
        assert rel(p.grad, g["grad%d" % i]) <= 2e-4


@pytest.mark.parametrize("name", ["edges_2x96x128.npz", "edges_2x37x53.npz", "edges_1x256x256.npz"])
def test_edge_maps_vs_cv2_golden(name):
    from eel_unet_b200 import edges
    from oracle import synth

    g = np.load(os.path.join(GOLD, name))
    n, h, w = g["gray"].shape
    imgs, masks = synth.tooth_images(n, h, w, seed=int(g["seed"]))
    d = torch.from_numpy(imgs).cuda()
    gray = edges.gray(d)
    assert np.array_equal(gray.cpu().numpy(), g["gray"])
    assert np.array_equal(edges.canny(d).cpu().numpy(), g["canny"])
    assert np.array_equal(edges.sobel_map(gray).cpu().numpy(), g["sobel"])
    assert np.array_equal(edges.laplacian_map(gray).cpu().numpy(), g["laplacian"])
    assert np.array_equal(edges.canny_enhance(d, edge_color=(255, 255, 255), alpha=0.2).cpu().numpy(), g["enhance"])
    lab = edges.edge_label(torch.from_numpy(masks).cuda())
    assert np.array_equal((lab[:, 0].cpu().numpy() * 255).astype(np.uint8), g["label"])
