#!/usr/bin/env python
"""bench.py -- EEL-Unet training throughput (BASELINE.json metric: train images/s at 256^2, % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision bf16|fp32] [--size 256] [--batch 64] [--no-cpu-baseline]

One step = one pass of the hot path over one batch of synthetic tooth-like images: EELUnet forward,
edge_BceDiceLoss, backward, (N > 1: bucketed NCCL gradient all-reduce overlapped with backward) and the
fused Adam update.  Workload at every N: BASELINE config 3 per GPU (bf16 training, batch 64 at 3x256x256),
i.e. weak scaling.  Prints ONE JSON line (rank 0).

`--impl reference` times the CPU oracle port of the same step (oracle/eelunet_torch.py -- the reference is
Python and cannot travel to the GPU box) on all host threads, on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner under NCCL_DEBUG=VERSION, warnings):
# keep a private handle on the real stdout for the result and point file descriptor 1 at stderr for everything else.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], d["bf16_tflops"], "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Read through NVML from a thread of this
    process (nvidia_ml_py): two cheap queries every 25 ms.  The earlier `nvidia-smi -lms 100` child process is the fallback --
    its NVML attach and its polling loop were seen to hold up kernel launches (a timed region 10 % slower than the
    end-to-end region measured right after it, in which the sampler no longer ran)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, mode="nvml"):
        self.ordinal = index
        # nvidia-smi / NVML number the PHYSICAL devices: translate the process-local ordinal through CUDA_VISIBLE_DEVICES
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if index < len(vis) and vis[index].isdigit():
            index = int(vis[index])
        self.index, self.proc, self.lines, self.mode = index, None, [], mode
        self.samples, self.nv, self.stop_flag, self.th = [], None, False, None

    # ---- NVML thread
    def _nvml_open(self):
        import pynvml

        pynvml.nvmlInit()
        h = None
        try:
            import torch

            pr = torch.cuda.get_device_properties(self.ordinal)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.nv = (pynvml, h, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

    def _nvml_loop(self):
        nv, h, _ = self.nv
        bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        # the other reasons NVML knows (a clock lock shows up as applications_clocks_setting), named when the binding has them
        for nm, attr in (("hw_power_brake_slowdown", "nvmlClocksEventReasonHwPowerBrakeSlowdown"),
                         ("applications_clocks_setting", "nvmlClocksEventReasonApplicationsClocksSetting"),
                         ("sync_boost", "nvmlClocksEventReasonSyncBoost"), ("display_clock_setting", "nvmlClocksEventReasonDisplayClockSetting")):
            if hasattr(nv, attr):
                bits.append((nm, getattr(nv, attr)))
        cfg = os.environ.get("EEL_BENCH_NVML", "25,cr").split(",")
        period, what = float(cfg[0]) / 1e3, cfg[1]
        while not self.stop_flag:
            try:
                t_q = time.time()
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM) if "c" in what else 0
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if "r" in what else 0
                self.samples.append((time.time(), float(mhz), [n for n, b in bits if r & b], time.time() - t_q))
            except Exception:
                pass
            time.sleep(period)

    def start(self):
        if self.mode == "none":
            return
        if self.mode == "nvml":
            try:
                self._nvml_open()
                self.th = threading.Thread(target=self._nvml_loop, daemon=True)
                self.th.start()
                return
            except Exception:
                self.nv, self.mode = None, "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            # nvidia-smi attaching to the driver (NVML initialisation) can hold up kernel launches for hundreds of milliseconds:
            # wait for its first sample so that this happens before the warm-up, never inside the timed region
            t0 = time.time()
            while not self.lines and time.time() - t0 < 15.0 and self.proc.poll() is None:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0=None, t1=None):
        """summary of the samples taken between wall-clock times t0 and t1 (the timed region); the sampler is started BEFORE the
        warm-up so that starting it cannot stall the launching thread inside the timed region"""
        if self.mode == "none":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler switched off (--clock-sampler none)"]}
        if self.nv is not None:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            inside = [x for x in self.samples if (t0 is None or x[0] >= t0) and (t1 is None or x[0] <= t1 + 0.05)] or self.samples[-3:]
            sm = [x[1] for x in inside]
            reasons = sorted({r for x in inside for r in x[2]})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.nv[2]), "reasons": reasons,
                    "samples": len(sm), "source": "nvml (in-process thread, 25 ms)",
                    "slowest_query_ms": round(1e3 * max([x[3] for x in inside] or [0.0]), 2)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        inside = [ln for (ts, ln) in self.lines if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.15)]
        for ln in (inside or [ln for _, ln in self.lines[-3:]]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def _workload(precision, world, B, S):
    """the `config` both arms print: the reference arm reports the SAME workload (its bounded sample is described in
    `cpu_baseline.sample`), so the driver compares like with like"""
    return {"workload": "EELUnet %s training (fwd + edge_BceDiceLoss + bwd + Adam%s), batch %d per GPU at 3x%dx%d"
                        % (precision, " + NCCL allreduce" if world > 1 else "", B, S, S),
            "global_batch": world * B, "parallelism": "dp%d" % world,
            "l2": "no flush needed: per-step activation working set is GBs >> 126 MB L2"}


def _oracle_params(device="cpu"):
    """seed-0 default-init weights in the reference's state_dict format from oracle/params.py -- NOT through the product
    package, whose import maps libeel.so"""
    import torch

    from oracle import params as P

    torch.manual_seed(0)
    sd = P.eelunet_state_dict(3, 1)
    return {k: v.to(device).requires_grad_(v.dtype.is_floating_point and "running_" not in k) for k, v in sd.items()}


def cpu_step_rate(batch, size, iters, warmup, seed=0):
    """images/s of the CPU oracle port (fp32, fwd + loss + bwd + Adam) on all host threads."""
    import torch

    from oracle import eelunet_torch as O
    from oracle import synth

    params = _oracle_params()
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-4, weight_decay=1e-5)
    xs, ys, _ = synth.batch(batch, size, size, seed)
    x, y = torch.from_numpy(xs), torch.from_numpy(ys)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        opt.zero_grad()
        seg, edges = O.forward(params, x, True, {})
        loss = O.edge_bce_dice_loss(edges, seg, y)
        loss.backward()
        opt.step()
        float(loss)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def gpu_eager_rate(dev, batch, size, mode, iters=4, warmup=2):
    """The bar SURVEY.md section 2a / 8d names: the same step as stock PyTorch eager on THIS GPU (cuDNN / cuBLAS / cuFFT
    Blackwell kernels) -- the oracle port on `cuda`, forward + edge_BceDiceLoss + backward + torch.optim.Adam(fused=True).
    mode 'fp32': plain eager.  mode 'bf16': torch.autocast(bfloat16) + channels_last input and conv weights (cuDNN's
    tensor-core NHWC path).  cudnn.benchmark on (the faster setting; the reference's train.py:32-33 turns it off).
    CUDA-event timed; halves the batch on out-of-memory.  Returns a dict or {'error': ...}."""
    import torch

    from oracle import eelunet_torch as O
    from oracle import synth

    torch.backends.cudnn.benchmark = True
    # 'fp32' is what a user gets from stock PyTorch: cuDNN convolutions run on TF32 tensor cores.  'fp32_strict' turns TF32 off
    # (IEEE fp32 FFMA everywhere): the like-for-like partner of this repo's fp32 mode, which is exact FFMA by design
    torch.backends.cudnn.allow_tf32 = mode != "fp32_strict"
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {"mode": mode, "size": size}
    while batch >= 1:
        params = opt = x = y = None
        try:
            params = _oracle_params(dev)
            if mode == "bf16":
                for k, v in params.items():
                    if v.dim() == 4:
                        v.data = v.data.contiguous(memory_format=torch.channels_last)
            opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-4, weight_decay=1e-5, fused=True)
            xs, ys, _ = synth.batch(min(batch, 16), size, size, 0)
            reps = (batch + xs.shape[0] - 1) // xs.shape[0]
            x = torch.from_numpy(xs).repeat(reps, 1, 1, 1)[:batch].to(dev)
            y = torch.from_numpy(ys).repeat(reps, 1, 1, 1)[:batch].to(dev)
            if mode == "bf16":
                x = x.contiguous(memory_format=torch.channels_last)

            def bce_dice(p, t):
                """utils/Loss.py:28-73 with the ops the reference itself calls (nn.BCELoss: its backward is finite at saturated
                probabilities, which bf16 produces; the oracle's hand-written log form is not)"""
                n = p.shape[0]
                p2, t2 = p.reshape(n, -1), t.reshape(n, -1)
                dice = 1 - ((2 * (p2 * t2).sum(1) + 1) / (p2.sum(1) + t2.sum(1) + 1)).sum() / n
                return dice + torch.nn.functional.binary_cross_entropy(p2, t2)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    seg, edges = O.forward(params, x, True, {})
                loss = bce_dice(seg.float(), y)
                for k, (e, wk) in enumerate(zip(edges, O.EDGE_WEIGHTS)):          # utils/Loss.py:97-113
                    sc = 16 >> k
                    loss = loss + wk * bce_dice(e.float(), torch.nn.functional.max_pool2d(y, sc, sc) if sc > 1 else y)
                loss.backward()
                opt.step()
                return loss

            for _ in range(warmup):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                loss = step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / iters
            out.update(batch=batch, ms_per_step=ms, value=batch / (ms / 1e3), unit="images/s", loss=float(loss.item()),
                       peak_mem_gb=round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2))
            return out
        except torch.OutOfMemoryError:
            batch //= 2
        except Exception as e:          # a baseline must never take the bench line down with it
            out["error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
            return out
        finally:
            del params, opt, x, y
            torch.cuda.empty_cache()
    out["error"] = "out of memory at batch 1"
    return out


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: this arm is rank 0 alone on the whole host
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # bounded sample: keep (K + W) CPU steps inside ~150 s at ~0.05 img/s/core
    per_step = 150.0 / max(1, args.steps + args.warmup)
    b = int(max(1, min(8, per_step * 0.055 * cores * (256.0 / args.size) ** 2)))
    rate, sec = cpu_step_rate(b, args.size, args.steps, args.warmup)
    line = {
        "metric": "EELUnet train images/sec @%d^2" % args.size, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": _workload(args.precision, world, args.batch, args.size),
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "oracle/eelunet_torch.py on host CPU, fp32 fwd+loss+bwd+Adam, batch %d per step at %d^2 (a bounded "
                                   "sample of the batch-%d workload; images/s does not depend on N: one host), %d timed steps"
                                   % (b, args.size, args.batch, args.steps)},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # self-check: nothing of the product may be mapped into this process (only oracle/ + torch CPU)
        "native_so_loaded": sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ROOT in ln and ".so" in ln}),
        "product_imported": "eel_unet_b200" in sys.modules,
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    from eel_unet_b200 import EELUnet, _lib, data, edge_BceDiceLoss, ops, profiling
    from eel_unet_b200.parallel import DataParallel, FusedAdam
    from eel_unet_b200 import synth  # numpy input generator (nothing under oracle/ is touched by the measured arm)

    hbm, tens_sus, tens_burst, peak_src = _peaks()
    ops.set_wgrad_stream(not args.no_wgrad_stream)
    B, S = args.batch, args.size
    torch.manual_seed(0)
    model = EELUnet(3, 1, precision=args.precision).to(dev).train()
    dp = DataParallel(model, bucket_mb=25.0)
    opt = FusedAdam(dp.buckets, lr=1e-4, weight_decay=1e-5)
    crit = edge_BceDiceLoss(1, 1)

    # synthetic data (seeded per rank): a few distinct images tiled to the batch keeps host prep short
    base = min(B, 16)
    xs, ys, _ = synth.batch(base, S, S, seed=rank)
    reps = (B + base - 1) // base
    x_host = torch.from_numpy(xs).repeat(reps, 1, 1, 1)[:B].contiguous().pin_memory()
    y_host = torch.from_numpy(ys).repeat(reps, 1, 1, 1)[:B].contiguous().pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        seg, edges = dp(x)
        loss = crit(edges, seg, y)
        loss.backward()
        dp.finish_backward()
        opt.step()
        # gradients are cleared at the END of a step (pure host work: 365 attribute resets): in the end-to-end loop, which
        # synchronises on loss.item() after every step, it then runs while the GPU is still busy instead of in front of the
        # next step's first launch
        dp.zero_grad()
        return loss

    dp.zero_grad()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, args.clock_sampler)
    if rank == 0:
        sampler.start()          # before the warm-up: the fork of nvidia-smi must not steal the launching thread's time
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    barrier()

    # ---- timed region: inputs resident in HBM --------------------------------------------------
    l0 = _lib.launch_count()
    ms0 = torch.cuda.memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.time()
    e0.record()
    marks = []
    for _ in range(args.steps):
        loss = step(x_dev, y_dev)
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()           # (an event record per step: no synchronisation, nothing waits on it)
    e1.record()
    barrier()
    t_region1 = time.time()
    ms = e0.elapsed_time(e1)
    per_step = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
    ms1 = torch.cuda.memory_stats(dev)
    # the caching allocator must be in its steady state inside the timed region: cudaMalloc / cudaFree calls there stall the
    # launching thread (and synchronise the device)
    alloc_delta = {k: ms1.get(k, 0) - ms0.get(k, 0) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries")}
    alloc_delta["reserved_gb"] = round(ms1.get("reserved_bytes.all.current", 0) / 2 ** 30, 2)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- end to end: pinned host buffers in, loss scalar out, copies inside the timed region -----
    # the public input hand-over (eel_unet_b200.data.DevicePrefetcher): every step's batch is copied from pinned host memory
    # inside the region -- batch i+1 on a copy stream while step i computes, the first one exposed -- and every step's loss is
    # read back before the next step is enqueued
    pf = data.DevicePrefetcher(device=dev)
    barrier()
    e0.record()
    pf.put(x_host, y_host)
    for i in range(args.steps):
        xd, yd = pf.get()
        if i + 1 < args.steps:
            pf.put(x_host, y_host)
        float(step(xd, yd).item())
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * B * args.steps / (float(t.item()) / 1e3)

    # ---- per-kernel CUDA-event profile of one more step (rank 0): roofline of the dominant kernel ----
    roof, breakdown = None, None
    rec = []
    if rank == 0:
        _lib.set_profiler(rec)
    # kernels are timed ALONE here: the weight-gradient stream (ops._Wgrad) is switched off for this one step, otherwise an
    # event pair around a launch would also span whatever the other stream runs next to it
    ops.set_wgrad_stream(False)
    step(x_dev, y_dev)            # every rank runs it: the step contains collectives
    ops.set_wgrad_stream(not args.no_wgrad_stream)
    _lib.set_profiler(None)
    barrier()
    if rank == 0:
        fam = profiling.summarize(rec)
        rows = profiling.table(fam, hbm, tens_sus)
        breakdown = [{"kernel": r[0], "calls": r[1], "ms": round(r[2], 3), "share": round(r[3], 2), "tflops": round(r[4], 2),
                      "gbs": round(r[6], 1)} for r in rows[:12]]
        top = rows[0]
        d = fam[top[0]]
        if top[0] in profiling.GEMM_CLASS:
            ach = d["flops"] / d["ms"] / 1e9
            roof = {"bound": "tensor", "kernel": top[0], "achieved": ach, "peak": tens_sus, "unit": "TFLOP/s", "frac": ach / tens_sus,
                    "traffic": None, "peak_source": peak_src + " (sustained cuBLAS bf16)", "share_of_step": top[3] / 100.0}
        else:
            ach = d["bytes"] / d["ms"] / 1e6
            roof = {"bound": "hbm", "kernel": top[0], "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": None, "peak_source": peak_src + " (copy)", "share_of_step": top[3] / 100.0}
        if top[0] == "tc_conv3x3":
            # the family holds two kinds of launch: plain implicit-GEMM convolutions, and data gradients whose epilogue also reads
            # the BatchNorm input z and accumulates that BatchNorm's backward sums (HBM work without FLOPs): both rates
            for tag, names in (("plain_conv_launches", ("eel_tc_conv3x3", "eel_tc_conv3x3_2src")), ("dgrad_with_bn_sums_launches", ("eel_tc_conv3x3_dgrad_bnsums", "eel_tc_conv3x3_dgrad_split"))):
                sel = [(profiling.cost(nm, a)[0], s0.elapsed_time(s1)) for nm, a, s0, s1 in rec if nm in names]
                if sel:
                    fl, tm = sum(v[0] for v in sel), sum(v[1] for v in sel)
                    roof[tag] = {"calls": len(sel), "ms": round(tm, 3), "achieved": fl / tm / 1e9, "frac": fl / tm / 1e9 / tens_sus}
        # DRAM traffic of the dominant family from the committed ncu --set full capture of the same step (bytes per launch,
        # like `achieved`); null when no capture of this family is on file
        tj = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tj):
            t = json.load(open(tj)).get(top[0])
            if t:
                roof["traffic"] = t["dram_bytes_per_launch"]
                roof["traffic_algorithmic"] = d["bytes"] / max(1, d["calls"])
                roof["traffic_source"] = t["source"]
        # the best bandwidth-bound kernel family, for the north_star's ">= 70 % of HBM roofline" target
        ew = [r for r in rows if r[0] not in profiling.GEMM_CLASS and fam[r[0]]["bytes"] > 1e8]
        if ew:
            tot_b = sum(fam[r[0]]["bytes"] for r in ew)
            tot_ms = sum(fam[r[0]]["ms"] for r in ew)
            roof["elementwise_hbm_frac"] = tot_b / tot_ms / 1e6 / hbm
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                f.write("kernel,calls,ms,share_pct,tflops,pct_tensor_peak,gbs,pct_hbm_peak\n")
                for r in rows:
                    f.write("%s,%d,%.3f,%.2f,%.2f,%.2f,%.1f,%.2f\n" % r)
        if args.profile_shapes:
            with open(args.profile_shapes, "w") as f:
                f.write("kernel,int_args,calls,ms,tflops,gbs\n")
                for r in profiling.shape_table(rec):
                    f.write("%s,%s,%d,%.3f,%.2f,%.1f\n" % r)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cores = torch.get_num_threads()
        b = 4 if cores >= 8 else 2
        rate, sec = cpu_step_rate(b, S, 1, 1)
        cpu = {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "oracle/eelunet_torch.py fp32 fwd+loss+bwd+Adam, batch %d at %d^2, 1 warm-up + 1 timed step" % (b, S)}

    # ---- the real bar: stock PyTorch eager (cuDNN / cuBLAS / cuFFT) on this same GPU, same step, same batch --------------
    eager = None
    final_loss = float(loss.item())
    peak_mem = round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)
    if rank == 0 and world == 1 and not args.no_eager_baseline:
        dp.buckets.remove()
        del dp, opt, model, x_dev, y_dev, loss
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        eager = {"what": "oracle/eelunet_torch.py (the reference's op sequence) as stock PyTorch eager on this GPU: fwd + "
                         "edge_BceDiceLoss + bwd + torch.optim.Adam(fused=True), cudnn.benchmark on, CUDA-event timed in this process",
                 "fp32": gpu_eager_rate(dev, B, S, "fp32"),
                 "bf16_autocast_channels_last": gpu_eager_rate(dev, B, S, "bf16")}
        if args.precision == "fp32":
            eager["fp32_strict"] = gpu_eager_rate(dev, B, S, "fp32_strict")
        for k in list(eager):
            if not isinstance(eager[k], dict):
                continue
            if eager[k].get("value"):
                eager[k]["ours_over_eager"] = round(value / eager[k]["value"], 3)

    if rank == 0:
        line = {
            "metric": "EELUnet train images/sec @%d^2" % S, "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": _workload(args.precision, world, B, S),
            "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 4,
                    "d2h_bytes_per_step": 4,
                    "pipeline": "data.DevicePrefetcher: pinned host batch i+1 copied on a copy stream during step i (first copy "
                                "exposed); loss.item() after every step"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager,
            "kernels": breakdown, "loss": final_loss, "peak_mem_gb": peak_mem,
            # spread of the timed steps on this rank (`value` is steps / the whole region, as the contract says)
            "step_ms": {"min": min(per_step), "median": statistics.median(per_step), "max": max(per_step),
                        "slow": [(i, v) for i, v in enumerate(per_step) if v > 1.15 * statistics.median(per_step)]},
            "allocator_in_timed_region": alloc_delta,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "none"])
    ap.add_argument("--no-wgrad-stream", action="store_true", help="weight gradients on the launching stream (A/B of ops._Wgrad)")
    ap.add_argument("--profile-out", default=None)
    ap.add_argument("--profile-shapes", default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
