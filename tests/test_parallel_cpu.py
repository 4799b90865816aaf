"""CPU (gloo, world_size 2): host logic of the data-parallel gradient buckets -- flat re-homing of parameters,
bucket cut, post-accumulate hooks, mean all-reduce, set_to_none cycle.  The EELUnet kernels need a GPU, so the
network here is a small torch module; what is under test is eel_unet_b200.parallel, not the model."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.Tanh(), torch.nn.Linear(13, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from eel_unet_b200.parallel import GradBuckets

        net = _net()
        before = [p.detach().clone() for p in net.parameters()]
        gb = GradBuckets(list(net.parameters()), bucket_mb=0.0003)      # tiny buckets -> several of them
        assert len(gb.buckets) >= 3
        for p, b in zip(net.parameters(), before):                      # re-homing keeps values, makes views of one buffer
            assert torch.equal(p.detach(), b)
            assert p.data_ptr() >= gb.flat_param.data_ptr() and p.data_ptr() < gb.flat_param.data_ptr() + gb.flat_param.numel() * 4
        torch.manual_seed(100 + rank)
        for step in range(2):
            x = torch.randn(4, 7)
            gb.zero_grad()
            net(x).pow(2).mean().backward()
            gb.finish()
            out[(rank, step)] = ([p.grad.clone() for p in net.parameters()], x)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_equals_mean_of_rank_gradients():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for step in range(2):
        xs = [out[(r, step)][1] for r in range(world)]
        ref = None
        for x in xs:
            net = _net()
            net(x).pow(2).mean().backward()
            g = [p.grad for p in net.parameters()]
            ref = g if ref is None else [a + b for a, b in zip(ref, g)]
        ref = [g / world for g in ref]
        for r in range(world):
            for a, b in zip(out[(r, step)][0], ref):
                assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_single_process_buckets_and_views():
    from eel_unet_b200.parallel import GradBuckets

    net = _net()
    gb = GradBuckets(list(net.parameters()), bucket_mb=25.0)
    assert gb.world == 1 and len(gb.buckets) == 1
    net(torch.randn(3, 7)).sum().backward()
    gb.finish()
    for p in net.parameters():
        o = (p.grad.data_ptr() - gb.flat_grad.data_ptr()) // 4
        assert 0 <= o < gb.flat_grad.numel()                             # gradients live in the flat buffer
    gb.flat_param.add_(1.0)                                              # an optimizer step on the flat buffer ...
    assert all((p.detach() - 1.0).abs().max() < 10 for p in net.parameters())   # ... is seen through the parameters
    gb.zero_grad()
    assert all(p.grad is None for p in net.parameters())


def test_fused_adam_optimizer_contract_on_cpu():
    """Host logic of FusedAdam without launching the kernel: it is a torch.optim.Optimizer whose single param group StepLR
    rewrites, and its state_dict()/load_state_dict() speak torch.optim.Adam's format."""
    from torch.optim.lr_scheduler import StepLR

    from eel_unet_b200.parallel import FusedAdam, GradBuckets

    net, ref = _net(), _net()
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-5)
    for _ in range(3):
        ropt.zero_grad()
        ref(torch.randn(4, 7)).pow(2).mean().backward()
        ropt.step()
    gb = GradBuckets(list(net.parameters()))
    opt = FusedAdam(gb, lr=1e-3, weight_decay=1e-5)
    assert isinstance(opt, torch.optim.Optimizer)
    sch = StepLR(opt, 30, 0.5)                                    # reference train.py:315
    assert opt.state_dict()["state"] == {}
    opt.load_state_dict(ropt.state_dict())
    assert opt.t == 3
    back = opt.state_dict()
    assert back["param_groups"][0]["params"] == list(range(6)) and back["param_groups"][0]["weight_decay"] == 1e-5
    for i, st in ropt.state_dict()["state"].items():
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"]) and torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(back["state"][i]["step"]) == 3.0
    fresh = torch.optim.Adam(_net().parameters(), lr=1.0)
    fresh.load_state_dict(back)                                   # torch.optim.Adam accepts our checkpoint
    assert fresh.param_groups[0]["lr"] == 1e-3
    # ranges the kernel would update when the middle layer got no gradient
    gb.missing = [list(net.parameters())[2]]
    r = opt._ranges()
    o, n = gb._slot[id(gb.missing[0])]
    assert all(hi <= o or lo >= o + n for lo, hi, _ in r) and sum(hi - lo for lo, hi, _ in r) == gb.flat_param.numel() - (n + 3) // 4 * 4
    assert all(t == 3 for _, _, t in r)
    gb.missing = []
    assert opt._ranges() == [(0, gb.flat_param.numel(), 3)]
    opt._lag = {id(list(net.parameters())[2]): 2}                 # a parameter that missed two steps runs two steps behind
    r = opt._ranges()
    assert sum(hi - lo for lo, hi, _ in r) == gb.flat_param.numel() and sorted({t for _, _, t in r}) == [1, 3]
    assert float(opt.state_dict()["state"][2]["step"]) == 1.0
    opt._lag = {}
    import pytest
    with pytest.raises(Exception):
        opt.step()                                                # no CPU fallback for the update kernel
