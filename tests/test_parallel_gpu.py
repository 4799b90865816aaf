"""GPU, >= 2 devices: data-parallel EELUnet over NCCL against single-process shard runs (tests/nccl_worker.py).
On a one-GPU box the test skips; `gpurun --gpus 2 -- python -m pytest tests/test_parallel_gpu.py -m gpu` runs it
(log kept in profiles/r02_nccl_parity_2gpu.log)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_nccl_data_parallel_gradient_equals_mean_of_shard_gradients():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert len(lines) == world and all(ln["ok"] for ln in lines)
    for ln in lines:
        assert ln["fp32_worst_grad_rel"] < 1e-3 and ln["fp32_buckets"] >= 3
        assert ln["fp32_ranks_hold_same_gradient"] and ln["bf16_ranks_hold_same_gradient"]
        assert ln["fp32_weights_identical_after_adam"] and ln["bf16_weights_identical_after_adam"]


def test_model_runs_on_a_device_that_is_not_current():
    """a model on cuda:1 while the current device is cuda:0 must launch on cuda:1's context / stream (device guard
    in _lib.call, per-device kernel attributes)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from eel_unet_b200 import EELUnet, edge_BceDiceLoss

    torch.cuda.set_device(0)
    torch.manual_seed(0)
    x = torch.randn(2, 3, 64, 64)
    y = (torch.rand(2, 1, 64, 64) > 0.5).float()
    outs = []
    for d in (0, 1):
        torch.manual_seed(1)
        m = EELUnet(3, 1, precision="bf16").to("cuda:%d" % d).train()
        seg, edges = m(x.to("cuda:%d" % d))
        loss = edge_BceDiceLoss(1, 1)(edges, seg, y.to("cuda:%d" % d))
        loss.backward()
        torch.cuda.synchronize(d)
        assert seg.device.index == d and torch.isfinite(seg).all()
        outs.append(loss.item())
    assert torch.cuda.current_device() == 0
    assert abs(outs[0] - outs[1]) < 0.05 * abs(outs[0])
