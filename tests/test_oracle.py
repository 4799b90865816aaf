"""CPU: pin the oracle (oracle/) to the fixtures the real reference produced (tests/golden/, made by
tests/golden/make_golden.py) and -- when /root/reference is present (build container) -- to the live reference."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def _weights():
    """Seed-0 default-init weights from the oracle's own layer table (oracle/params.py: the reference's construction order)."""
    from oracle import params

    torch.manual_seed(0)
    return params.eelunet_state_dict(3, 1)


def test_drop_in_module_tree_initialises_like_the_oracle_table():
    """eel_unet_b200.EELUnet / Unet (the nn.Module trees a user constructs) and oracle/params.py draw the same weights
    from the same seed -- key order, shapes, values."""
    from eel_unet_b200.model import EELUnet
    from eel_unet_b200.unet import Unet
    from oracle import params

    for cin, cout, seed in ((3, 1, 0), (4, 3, 5)):
        torch.manual_seed(seed)
        a = EELUnet(cin, cout).state_dict()
        torch.manual_seed(seed)
        b = params.eelunet_state_dict(cin, cout)
        assert list(a.keys()) == list(b.keys()) and len(a) == 365
        assert all(torch.equal(a[k], b[k]) for k in a)
    torch.manual_seed(2)
    a = Unet(3, 2).state_dict()
    torch.manual_seed(2)
    b = params.unet_state_dict(3, 2)
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)


def test_seed0_weights_match_reference_checksums():
    g = np.load(os.path.join(GOLD, "eelunet_train_2x128.npz"))
    sd = _weights()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    chk = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in sd.values()])
    assert np.allclose(chk, g["state_checksums"], rtol=1e-12, atol=1e-12)


def test_oracle_train_step_matches_reference_golden():
    from oracle import eelunet_torch as O
    from oracle import synth

    g = np.load(os.path.join(GOLD, "eelunet_train_2x128.npz"))
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in _weights().items()}
    xs, ys, _ = synth.batch(2, 128, 128, 0)
    loss, seg, edges, grads, ns = O.train_step(sd, torch.from_numpy(xs).double(), torch.from_numpy(ys).double())
    assert abs(loss.item() - float(g["loss"])) < 1e-12
    assert rel(seg, g["seg"]) < 1e-6                      # fixture stored as fp32
    for k, e in enumerate(edges):
        assert rel(e, g["edge%d" % (5 - k)]) < 1e-6
    names = [str(n) for n in g["grad_names"]]
    for n, gn, gs in zip(names, g["grad_norm"], g["grad_sum"]):
        assert abs(grads[n].norm().item() - gn) <= 1e-9 * max(gn, 1e-12) + 1e-14, n
        assert abs(grads[n].sum().item() - gs) <= 1e-8 * max(gn, 1e-12) * grads[n].numel() ** 0.5 + 1e-14, n
    for key in g.files:
        if key.startswith("grad:"):
            assert rel(grads[key[5:]], g[key]) < 1e-9 or np.linalg.norm(g[key]) < 1e-12, key
        if key.startswith("stat:"):
            assert rel(ns[key[5:]], g[key]) < 1e-12, key


def test_oracle_eval_forward_matches_reference_golden():
    from oracle import eelunet_torch as O
    from oracle import synth

    g = np.load(os.path.join(GOLD, "eelunet_eval_2x128.npz"))
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in _weights().items()}
    xs, _, _ = synth.batch(2, 128, 128, 0)
    with torch.no_grad():
        seg, edges = O.forward(sd, torch.from_numpy(xs).double(), False)
    assert rel(seg, g["seg"]) < 1e-6
    for k, e in enumerate(edges):
        assert rel(e, g["edge%d" % (5 - k)]) < 1e-6


def test_oracle_loss_matches_reference_golden():
    from oracle import eelunet_torch as O

    g = np.load(os.path.join(GOLD, "loss_cases.npz"))
    preds = [torch.from_numpy(g["pred%d" % i]).requires_grad_(True) for i in range(6)]
    loss = O.edge_bce_dice_loss(preds[1:], preds[0], torch.from_numpy(g["target"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-12
    for i, p in enumerate(preds):
        assert rel(p.grad, g["grad%d" % i]) < 1e-12


def test_hft_restatement_equals_fft_definition():
    """oracle.hft (mask built in unshifted order) == the reference's shift / mask / unshift formulation."""
    from oracle import eelunet_torch as O

    torch.manual_seed(0)
    for h, w in [(64, 64), (48, 80), (16, 16), (33, 47)]:
        x = torch.randn(2, 3, h, w, dtype=torch.float64)
        crow, ccol = h // 2, w // 2
        r = min(20, crow, ccol)
        mask = torch.ones(h, w, dtype=torch.float64)
        mask[crow - r:crow + r, ccol - r:ccol + r] = 0
        ref = torch.abs(torch.fft.ifft2(torch.fft.ifftshift(torch.fft.fftshift(torch.fft.fft2(x)) * mask)))
        assert (O.hft(x) - ref).abs().max().item() < 1e-12


@pytest.mark.parametrize("name", ["edges_2x96x128.npz", "edges_2x37x53.npz", "edges_1x256x256.npz"])
def test_edge_oracle_matches_cv2_golden(name):
    from oracle import edge_np, synth

    g = np.load(os.path.join(GOLD, name))
    n, h, w = g["gray"].shape
    imgs, masks = synth.tooth_images(n, h, w, seed=int(g["seed"]))
    gray = edge_np.gray_u8(imgs)
    assert np.array_equal(gray, g["gray"])
    assert np.array_equal(edge_np.canny(gray), g["canny"])
    assert np.array_equal(edge_np.sobel_map(gray), g["sobel"])
    assert np.array_equal(edge_np.laplacian_map(gray), g["laplacian"])
    assert np.array_equal(edge_np.canny_enhance(imgs, g["canny"], (255, 255, 255), 0.2), g["enhance"])
    assert np.array_equal((edge_np.edge_label(masks[:, 0]) * 255).astype(np.uint8), g["label"])


def test_edge_oracle_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import edge_np, synth

    rng = np.random.default_rng(3)
    for (h, w) in [(64, 80), (1, 1), (2, 7), (31, 33)]:
        imgs, _ = synth.tooth_images(2, h, w, seed=h + w)
        for img in list(imgs) + [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)]:
            g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY).reshape(h, w)
            assert np.array_equal(edge_np.gray_u8(img), g)
            assert np.array_equal(edge_np.canny(g), cv2.Canny(g, 100, 200).reshape(h, w))


def test_oracle_matches_live_reference():
    from oracle import ref_import

    if not ref_import.available():
        pytest.skip("reference only exists in the build container")
    from oracle import eelunet_torch as O
    from oracle import synth

    EELUnet, EdgeLoss, Unet = ref_import.load()
    torch.manual_seed(1)
    u = Unet(3, 1).double()
    xu = torch.randn(1, 3, 32, 48, dtype=torch.float64)
    with torch.no_grad():
        assert rel(O.unet_forward(u.state_dict(), xu), u(xu)) < 1e-12        # models/Unet.py:58-98
    torch.manual_seed(0)
    m = EELUnet(3, 1).double().train()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    xs, ys, _ = synth.batch(1, 128, 128, 3)
    x, y = torch.from_numpy(xs).double(), torch.from_numpy(synth.soften(ys)).double()
    seg, edges = m(x)
    loss = EdgeLoss(1, 1)(edges, seg, y)
    loss.backward()
    l2, seg2, e2, grads, ns = O.train_step(sd, x, y)
    assert abs(loss.item() - l2.item()) < 1e-12 and rel(seg2, seg.detach()) < 1e-12
    for n, p in m.named_parameters():
        if p.grad.norm().item() > 1e-9:
            assert rel(grads[n], p.grad) < 1e-9, n
    after = m.state_dict()
    for k, v in ns.items():
        assert rel(v, after[k]) < 1e-12, k


def test_resize_oracle_matches_pillow_goldens():
    """oracle/resize_np.py (restated Pillow Resample.c + ToTensor + Normalize) against outputs of Pillow / torchvision
    themselves (tests/golden/make_golden_resize.py)."""
    import os

    import numpy as np

    from oracle import resize_np as R

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_pil.npz"))
    for name in ["down2", "down_frac", "up", "square", "same_w", "tooth"]:
        oh, ow = (int(v) for v in g[name + "_size"])
        assert np.array_equal(R.resize_bilinear_u8(g[name + "_img"], oh, ow), g[name + "_img_resized"])
        assert np.array_equal(R.resize_bilinear_u8(g[name + "_mask"], oh, ow), g[name + "_mask_resized"])
        assert np.array_equal(R.preprocess_image(g[name + "_img"], (oh, ow)), g[name + "_img_tensor"])
        assert np.array_equal(R.preprocess_mask(g[name + "_mask"], (oh, ow)), g[name + "_mask_tensor"])


def test_metrics_oracle_matches_reference_evaluate_golden():
    """oracle/metrics_np.py against what the reference's OWN evaluate() / seg2bnd() / boundary_f1_score() returned
    (evaluate.py:25-124, run by tests/golden/make_golden_r2.py)."""
    from oracle import metrics_np

    g = np.load(os.path.join(GOLD, "metrics_eval.npz"))
    batches = metrics_np.seeded_batches()
    got = np.array(metrics_np.evaluate_batches(batches))
    assert np.abs(got - g["metrics"]).max() <= 1e-15, (got, g["metrics"])
    k = 0
    for seg, lab in batches:
        for i in range(seg.shape[0]):
            pred = (seg[i, 0] > 0.5).astype(np.float32)
            assert np.array_equal(np.packbits(metrics_np.seg2bnd(pred)), g["bnd_pred_%d" % k])
            assert np.array_equal(np.packbits(metrics_np.seg2bnd(lab[i, 0])), g["bnd_gt_%d" % k])
            assert abs(metrics_np.boundary_f1(lab[i, 0], pred) - g["boundary_f1"][k]) <= 1e-15
            k += 1


def test_metrics_oracle_matches_live_reference_evaluate():
    from oracle import ref_import

    if not ref_import.available():
        pytest.skip("reference only exists in the build container")
    from oracle import metrics_np

    ev = ref_import.load_evaluate()

    class Passthrough:
        name = "eelunet"

        def eval(self):
            return self

        def __call__(self, x):
            return x, []

    rng = np.random.default_rng(5)
    batches = []
    for (n, h, w) in [(2, 64, 80), (1, 33, 47)]:
        lab = (rng.uniform(size=(n, 1, h, w)) > 0.6).astype(np.float32)
        lab = (torch.nn.functional.avg_pool2d(torch.from_numpy(lab), 5, 1, 2) > 0.5).float().numpy()
        seg = np.clip(lab * 0.6 + rng.uniform(0, 0.6, size=lab.shape), 0, 1).astype(np.float32)
        batches.append((seg, lab))
    ref = ev.evaluate(Passthrough(), [(torch.from_numpy(s), torch.from_numpy(l)) for s, l in batches], torch.device("cpu"))
    got = metrics_np.evaluate_batches(batches)
    assert max(abs(a - b) for a, b in zip(got, ref)) <= 1e-15


def test_unet_oracle_matches_reference_golden_and_live_reference():
    """oracle.unet_forward against models/Unet.py: committed fp64 fixture, and the live module when present."""
    from oracle import eelunet_torch as O
    from oracle import params, ref_import, synth

    g = np.load(os.path.join(GOLD, "unet_2x64x96.npz"))
    torch.manual_seed(0)
    sd = {k: v.double().requires_grad_(True) for k, v in params.unet_state_dict(3, 1).items()}
    xs, ys, _ = synth.batch(2, 64, 96, 1)
    x, y = torch.from_numpy(xs).double(), torch.from_numpy(ys).double()
    out = O.unet_forward(sd, x)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()
    assert rel(out.detach(), g["logits"]) < 1e-12 and abs(loss.item() - float(g["loss"])) < 1e-13
    for n, gn, gs in zip([str(s) for s in g["grad_names"]], g["grad_norm"], g["grad_sum"]):
        assert abs(sd[n].grad.norm().item() - gn) <= 1e-9 * gn + 1e-15, n
        assert abs(sd[n].grad.sum().item() - gs) <= 1e-8 * gn * sd[n].numel() ** 0.5 + 1e-15, n
    for key in g.files:
        if key.startswith("grad:"):
            assert rel(sd[key[5:]].grad, g[key]) < 1e-9, key
    if ref_import.available():
        _, _, Unet = ref_import.load()
        torch.manual_seed(4)
        m = Unet(3, 2).double()
        x2 = torch.randn(1, 3, 48, 80, dtype=torch.float64)
        ref = m(x2)
        ref.pow(2).mean().backward()
        sd2 = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        got = O.unet_forward(sd2, x2)
        got.pow(2).mean().backward()
        assert rel(got.detach(), ref.detach()) < 1e-12
        for n, p in m.named_parameters():
            assert rel(sd2[n].grad, p.grad) < 1e-9, n


def test_oracle_train_step_matches_reference_golden_at_config1_shape():
    """BASELINE config 1's shape (8 x 3 x 256 x 256): the oracle against the fp64 train step of the real reference
    (tests/golden/eelunet_train_8x256.npz, tests/golden/make_golden_r2.py)."""
    from oracle import eelunet_torch as O
    from oracle import synth

    g = np.load(os.path.join(GOLD, "eelunet_train_8x256.npz"))
    sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in _weights().items()}
    xs, ys, _ = synth.batch(8, 256, 256, 0)
    loss, seg, edges, grads, ns = O.train_step(sd, torch.from_numpy(xs).double(), torch.from_numpy(ys).double())
    assert abs(loss.item() - float(g["loss"])) < 1e-11
    assert abs(seg.sum().item() - float(g["seg_sum"])) < 1e-7 and abs(seg.pow(2).sum().item() - float(g["seg_sqsum"])) < 1e-7
    assert rel(seg, g["seg"].astype(np.float64)) < 1e-3                      # stored as fp16 (size); the sums above are fp64
    for k, e in enumerate(edges):
        assert abs(e.sum().item() - float(g["edge%d_sum" % (5 - k)])) < 1e-8
        if "edge%d" % (5 - k) in g.files:
            assert rel(e, g["edge%d" % (5 - k)]) < 1e-6
    for n, gn, gs in zip([str(s) for s in g["grad_names"]], g["grad_norm"], g["grad_sum"]):
        assert abs(grads[n].norm().item() - gn) <= 1e-8 * max(gn, 1e-12) + 1e-13, n
    for key in g.files:
        if key.startswith("stat:"):
            assert rel(ns[key[5:]], g[key]) < 1e-11, key
