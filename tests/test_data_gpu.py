"""GPU input pipeline (eel_unet_b200.data) against Pillow / torchvision goldens (tests/golden/resize_pil.npz, produced by
tests/golden/make_golden_resize.py with the third-party code the reference's dataset calls) and against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_pil.npz")
CASES = ["down2", "down_frac", "up", "square", "same_w", "tooth"]


@pytest.mark.parametrize("name", CASES)
def test_resize_and_tensor_match_pillow_bit_exact(name):
    from eel_unet_b200 import data

    g = np.load(GOLD)
    oh, ow = (int(v) for v in g[name + "_size"])
    img = torch.from_numpy(g[name + "_img"]).cuda()[None]
    mask = torch.from_numpy(g[name + "_mask"]).cuda()[None]
    assert np.array_equal(data.resize_u8(img, (oh, ow))[0].cpu().numpy(), g[name + "_img_resized"])
    assert np.array_equal(data.resize_u8(mask, (oh, ow))[0].cpu().numpy(), g[name + "_mask_resized"])
    t = data.preprocess_images(img, (oh, ow))[0].cpu().numpy()
    assert t.shape == g[name + "_img_tensor"].shape and np.abs(t - g[name + "_img_tensor"]).max() <= 1e-6
    m = data.preprocess_masks(mask, (oh, ow))[0].cpu().numpy()
    assert m.shape == g[name + "_mask_tensor"].shape and np.array_equal(m, g[name + "_mask_tensor"])


def test_batch_matches_oracle_and_feeds_the_model():
    from eel_unet_b200 import EELUnet, data
    from oracle import resize_np, synth

    imgs, masks, _ = (None, None, None)
    imgs = synth.tooth_images(3, 80, 112, seed=2)[0]
    d = torch.from_numpy(imgs).cuda()
    x = data.preprocess_images(d, (64, 96))
    ref = np.stack([resize_np.preprocess_image(imgs[i], (64, 96)) for i in range(3)])
    assert np.abs(x.cpu().numpy() - ref).max() <= 1e-6
    seg, edges = EELUnet(3, 1).cuda().eval()(x)            # the tensor is what train.py:38 hands the model
    assert tuple(seg.shape) == (3, 1, 64, 96) and len(edges) == 5


def test_rejects_cpu_and_bad_arguments():
    from eel_unet_b200 import _lib, data

    with pytest.raises(_lib.EelError):
        data.preprocess_images(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), (4, 4))
    with pytest.raises(_lib.EelError):
        data.preprocess_images(torch.zeros(1, 8, 8, 3, device="cuda"), (4, 4))
    with pytest.raises(_lib.EelError):
        data.preprocess_images(torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda"), (4, 4), mean=(0.5,), std=(0.5,))


def test_device_prefetcher_hands_over_the_batches_in_order():
    """data.DevicePrefetcher: double-buffered host -> device copies on a copy stream; every batch arrives intact and in order
    while the compute stream is busy, through the iterator and through put / get"""
    from eel_unet_b200 import data

    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(4, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 2, (4, 1, 32, 32), generator=g).float().pin_memory())
               for _ in range(7)]
    busy = torch.randn(2048, 2048, device=dev)
    seen = []
    for x, y in data.DevicePrefetcher(batches, dev):
        assert x.is_cuda and y.is_cuda
        busy = busy @ busy * 1e-3                # keep the compute stream occupied between hand-overs
        seen.append((x.clone(), y.clone()))
    torch.cuda.synchronize()
    assert len(seen) == len(batches)
    for (x, y), (hx, hy) in zip(seen, batches):
        assert torch.equal(x.cpu(), hx) and torch.equal(y.cpu(), hy)
    pf = data.DevicePrefetcher(device=dev)
    pf.put(*batches[0])
    pf.put(*batches[1])
    with pytest.raises(Exception):
        pf.put(*batches[2])                      # both slots in flight
    a = pf.get()
    pf.put(*batches[2])
    b = pf.get()
    c = pf.get()
    torch.cuda.synchronize()
    assert torch.equal(b[0].cpu(), batches[1][0]) and torch.equal(c[0].cpu(), batches[2][0])
    with pytest.raises(Exception):
        pf.get()
