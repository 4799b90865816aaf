"""Worker of tests/test_parallel_gpu.py (one process per GPU, launched by torch.distributed.run, NCCL).

Every rank trains EELUnet on ITS shard of a global batch through eel_unet_b200.parallel (flat buffers, bucketed all-reduce
overlapped with backward).  Afterwards every rank recomputes, single-process and without any collective, the gradient of
EVERY shard with a plain model and averages them: the data-parallel flat gradient must equal that mean (SURVEY.md section 8e:
"N-rank result vs single-process run on each shard with gradients averaged"), the Adam step must leave identical weights on all
ranks, and BatchNorm running statistics must stay per replica (the reference has no SyncBN).
Prints one JSON line per rank; exit code != 0 on failure.
"""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    from eel_unet_b200 import EELUnet, edge_BceDiceLoss, synth
    from eel_unet_b200.parallel import DataParallel, FusedAdam

    per, size = 2, 128
    shards = []
    for r in range(world):
        xs, ys, _ = synth.batch(per, size, size, seed=50 + r)
        shards.append((torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)))
    crit = edge_BceDiceLoss(1, 1)
    report = {"rank": rank, "world": world}
    ok = True
    for precision in ("fp32", "bf16"):
        torch.manual_seed(0)
        model = EELUnet(3, 1, precision=precision).to(dev).train()
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        dp = DataParallel(model, bucket_mb=8.0)                 # several buckets -> several overlapped collectives
        opt = FusedAdam(dp.buckets, lr=1e-3, weight_decay=1e-5)
        x, y = shards[rank]
        dp.zero_grad()
        seg, edges = dp(x)
        crit(edges, seg, y).backward()
        dp.finish_backward()
        flat = dp.buckets.flat_grad.clone()
        # (1) all ranks hold the same averaged gradient
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same_grad = all(torch.equal(gathered[0], g) for g in gathered)
        # (2) it equals the mean of single-process shard gradients (plain model, fresh tensors, no flat buffer, no collective)
        if precision == "fp32":
            mean = None
            for (xs_, ys_) in shards:
                plain = EELUnet(3, 1, precision=precision).to(dev).train()
                plain.load_state_dict(sd0)
                s2, e2 = plain(xs_)
                crit(e2, s2, ys_).backward()
                g = {n: p.grad.detach().clone() for n, p in plain.named_parameters()}
                mean = g if mean is None else {n: mean[n] + g[n] for n in g}
                del plain
            mean = {n: v / world for n, v in mean.items()}
            gmax = max(v.norm().item() for v in mean.values())
            worst, worst_name, checked = 0.0, None, 0
            for n, p in model.named_parameters():
                if mean[n].norm().item() <= 1e-4 * gmax:
                    continue                                      # analytically-zero gradients: rounding noise only
                e = rel(p.grad, mean[n])
                checked += 1
                if e > worst:
                    worst, worst_name = e, n
            report["fp32_worst_grad_rel"] = worst
            report["fp32_worst_grad_name"] = worst_name
            report["fp32_tensors_checked"] = checked
            ok &= worst < 1e-3 and checked > 200
        else:
            ok &= bool(torch.isfinite(flat).all())
        ok &= same_grad
        report[precision + "_ranks_hold_same_gradient"] = same_grad
        report[precision + "_buckets"] = len(dp.buckets.buckets)
        # (3) the optimizer step leaves bit-identical weights everywhere; running statistics stay per replica
        opt.step()
        torch.cuda.synchronize()
        fp = dp.buckets.flat_param.clone()
        gathered = [torch.empty_like(fp) for _ in range(world)]
        dist.all_gather(gathered, fp)
        same_w = all(torch.equal(gathered[0], g) for g in gathered)
        rm = model.enc1[0][1].running_mean.clone()
        rms = [torch.empty_like(rm) for _ in range(world)]
        dist.all_gather(rms, rm)
        local_stats = not torch.equal(rms[0], rms[-1])
        report[precision + "_weights_identical_after_adam"] = same_w
        report[precision + "_batchnorm_stats_per_replica"] = local_stats
        ok &= same_w and local_stats
        dp.buckets.remove()
        del dp, opt, model
    report["ok"] = bool(ok)
    print(json.dumps(report), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
