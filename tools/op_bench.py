#!/usr/bin/env python
"""Micro-benchmark of single C-ABI kernels at the model's shapes (batch 64 at 256^2 by default).

    python tools/op_bench.py [--batch 64] [--iters 5] [--only conv,wgrad,bn,...] [--once]

Each case is timed with CUDA events on the launching stream after warm-up; between iterations the inputs are
larger than L2 or a 256 MB buffer is rewritten (L2 flush).  `--once` launches every case exactly once (for ncu).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from eel_unet_b200 import _lib, profiling
from eel_unet_b200._lib import call, ptr

BF16, F32 = torch.bfloat16, torch.float32
DEV = torch.device("cuda", 0)


def rnd(*shape, dtype=BF16):
    return torch.randn(*shape, device=DEV, dtype=torch.float32).to(dtype)


def cases(B, S, only):
    st = lambda: torch.cuda.current_stream().cuda_stream
    out = []

    def add(group, name, args_fn):
        if only and not any(group.startswith(o) or name.startswith(o) for o in only):
            return
        out.append((group, name, args_fn))

    conv_shapes = [(S, 64, 64), (S, 128, 64), (S // 2, 128, 128), (S // 2, 256, 128), (S // 2, 64, 128), (S // 4, 256, 256),
                   (S // 4, 512, 256), (S // 8, 512, 512), (S // 8, 1024, 512), (S // 16, 512, 1024)]
    for (s, ci, co) in conv_shapes:
        def mk(s=s, ci=ci, co=co):
            x, w, b, y = rnd(B, s, s, ci), rnd(3, 3, co, ci), rnd(co, dtype=F32), torch.empty(B, s, s, co, device=DEV, dtype=BF16)
            return "eel_tc_conv3x3", (ptr(x), ptr(w), ptr(b), ptr(y), B, s, s, ci, co, int(os.environ.get("EEL_RELU", "0")), 0, None, st()), (x, w, b, y)
        add("conv", "conv3x3 %dx%d %d->%d" % (s, s, ci, co), mk)
    for (s, ci, co) in conv_shapes:
        def mk(s=s, ci=ci, co=co):
            x, dy, dw = rnd(B, s, s, ci), rnd(B, s, s, co), torch.empty(3, 3, ci, co, device=DEV, dtype=F32)
            return "eel_tc_conv3x3_wgrad", (ptr(x), ptr(dy), ptr(dw), B, s, s, ci, co, st()), (x, dy, dw)
        add("wgrad", "wgrad3x3 %dx%d %d->%d" % (s, s, ci, co), mk)
    for (P, K, No) in [(B * (S // 4) ** 2, 256, 256), (B * (S // 4) ** 2, 256, 64), (B * (S // 4) ** 2, 64, 256), (B * (S // 8) ** 2, 512, 512),
                       (B * (S // 8) ** 2, 256, 512), (B * (S // 16) ** 2, 1024, 1024)]:
        def mk(P=P, K=K, No=No):
            x, w, b, y = rnd(P, K), rnd(No, K), rnd(No, dtype=F32), torch.empty(P, No, device=DEV, dtype=BF16)
            return "eel_tc_linear", (ptr(x), ptr(w), ptr(b), ptr(y), P, K, No, int(os.environ.get("EEL_RELU", "0")), None, 0, 0, st()), (x, w, b, y)
        add("linear", "linear P=%d %d->%d" % (P, K, No), mk)

        def mkw(P=P, K=K, No=No):
            x, dy, dw = rnd(P, K), rnd(P, No), torch.empty(No, K, device=DEV, dtype=F32)
            if No % 128 == 0:
                return "eel_tc_wgrad", (ptr(dy), ptr(x), ptr(dw), P, No, K, K, 1, dw.numel(), 0, st()), (x, dy, dw)
            return "eel_tc_wgrad", (ptr(x), ptr(dy), ptr(dw), P, K, No, 1, K, dw.numel(), 0, st()), (x, dy, dw)
        add("lwgrad", "lin-wgrad P=%d %d->%d" % (P, K, No), mkw)
    for (s, ci, co) in [(S // 2, 128, 64), (S // 4, 256, 128), (S // 8, 512, 256), (S // 16, 1024, 512)]:
        def mk(s=s, ci=ci, co=co):
            x, w, b, y = rnd(B, s, s, ci), rnd(2, 2, co, ci), rnd(co, dtype=F32), torch.empty(B, 2 * s, 2 * s, co, device=DEV, dtype=BF16)
            return "eel_tc_convt2x2_fwd", (ptr(x), ptr(w), ptr(b), ptr(y), B, s, s, ci, co, None, st()), (x, w, b, y)
        add("convt", "convT %dx%d %d->%d" % (s, s, ci, co), mk)
    for (s, c) in [(S, 64), (S // 2, 128), (S // 4, 256), (S // 8, 512)]:
        P = B * s * s

        def mk_stats(P=P, c=c):
            z, mean, rstd = rnd(P, c), torch.empty(c, device=DEV), torch.empty(c, device=DEV)
            rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
            n = _lib.lib.eel_reduce_workspace_bytes(c, 2)
            ws = torch.empty(n, dtype=torch.uint8, device=DEV)
            return "eel_bn_stats", (ptr(z), P, c, ptr(mean), ptr(rstd), ptr(rm), ptr(rv), 0.1, 1e-5, ptr(ws), n, 1, st()), (z, mean, rstd, rm, rv, ws)
        add("bn", "bn_stats P=%d C=%d" % (P, c), mk_stats)

        def mk_bwd(P=P, c=c):
            z, dy, dz = rnd(P, c), rnd(P, c), torch.empty(P, c, device=DEV, dtype=BF16)
            mean, rstd, g, b = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
            dg, db, dzs = torch.empty(c, device=DEV), torch.empty(c, device=DEV), torch.empty(c, device=DEV)
            n = _lib.lib.eel_reduce_workspace_bytes(c, 2) + 8 * c
            ws = torch.empty(n, dtype=torch.uint8, device=DEV)
            return "eel_bn_act_bwd", (ptr(dy), ptr(z), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(dz), ptr(dg), ptr(db), ptr(dzs), P, c, 1, 1,
                                      ptr(ws), n, 1, st()), (z, dy, dz, mean, rstd, g, b, dg, db, dzs, ws)
        add("bn", "bn_act_bwd P=%d C=%d" % (P, c), mk_bwd)

        def mk_pool_fwd(s=s, c=c):
            z, a_, pooled = rnd(B, s, s, c), torch.empty(B, s, s, c, device=DEV, dtype=BF16), torch.empty(B, s // 2, s // 2, c, device=DEV, dtype=BF16)
            amax = torch.empty(B * (s // 2) * (s // 2), c // 8, device=DEV, dtype=torch.int16)
            mean, rstd, g, b = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
            return "eel_bn_relu_pool_fwd", (ptr(z), ptr(a_), ptr(pooled), ptr(amax), ptr(mean), ptr(rstd), ptr(g), ptr(b), B, s, s, c, 1, st()), \
                (z, a_, pooled, amax, mean, rstd, g, b)
        add("pool", "bn_relu_pool_fwd %dx%d C=%d" % (s, s, c), mk_pool_fwd)

        def mk_pool_bwd(s=s, c=c):
            z, da, dz = rnd(B, s, s, c), rnd(B, s, s, c), torch.empty(B, s, s, c, device=DEV, dtype=BF16)
            dp = rnd(B, s // 2, s // 2, c)
            amax = torch.randint(0, 1 << 15, (B * (s // 2) * (s // 2), c // 8), device=DEV, dtype=torch.int16)
            mean, rstd, g, b = torch.zeros(c, device=DEV), torch.ones(c, device=DEV), torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
            dg, db, dzs = torch.empty(c, device=DEV), torch.empty(c, device=DEV), torch.empty(c, device=DEV)
            n = _lib.lib.eel_reduce_workspace_bytes(c, 2) + 8 * c
            ws = torch.empty(n, dtype=torch.uint8, device=DEV)
            return "eel_bn_relu_pool_bwd", (ptr(da), ptr(dp), ptr(z), ptr(amax), ptr(mean), ptr(rstd), ptr(g), ptr(b), ptr(dz), ptr(dg), ptr(db),
                                            ptr(dzs), B, s, s, c, 1, ptr(ws), n, 1, st()), (z, da, dz, dp, amax, mean, rstd, g, b, dg, db, dzs, ws)
        add("pool", "bn_relu_pool_bwd %dx%d C=%d" % (s, s, c), mk_pool_bwd)
    for (s, c) in [(S // 4, 256), (S // 8, 512), (S // 16, 1024)]:
        def mk_mlp(s=s, c=c):
            P = B * s * s
            u, w0, b0, wc, bc = rnd(P, 64), rnd(256, 64), rnd(256, dtype=F32), rnd(c, 256), rnd(c, dtype=F32)
            h, a_, z = (torch.empty(P, 256, device=DEV, dtype=BF16), torch.empty(P, 256, device=DEV, dtype=BF16),
                        torch.empty(P, c, device=DEV, dtype=BF16))
            return "eel_tc_capmlp_fwd", (ptr(u), ptr(w0), ptr(b0), ptr(wc), ptr(bc), ptr(h), ptr(a_), ptr(z), P, c, 0, None, st()), \
                (u, w0, b0, wc, bc, h, a_, z)
        add("mlp", "capmlp_fwd P=%d C=%d" % (B * s * s, c), mk_mlp)
    for (s_, c_) in [(S // 4, 256), (S // 8, 512)]:
        def mk_shift(s_=s_, c_=c_):
            x, y = rnd(B, s_, s_, c_), torch.empty(B, s_, s_, c_, device=DEV, dtype=BF16)
            return "eel_shift_channels", (ptr(x), ptr(y), B, s_, s_, c_, 0, 1, st()), (x, y)
        add("shift", "shift_channels %dx%d C=%d" % (s_, s_, c_), mk_shift)

    def mk_head(fwd):
        P = B * S * S
        x, dx = rnd(B, S, S, 64), torch.empty(B, S, S, 64, device=DEV, dtype=BF16)
        lnw, lnb, w, b = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV), torch.randn(64, device=DEV) / 8, torch.zeros(1, device=DEV)
        prob, dprob = torch.rand(P, device=DEV), torch.randn(P, device=DEV)
        g = [torch.empty(64, device=DEV) for _ in range(3)] + [torch.empty(1, device=DEV)]
        n = 4 * (4 * _lib.lib.eel_num_sms() + 1) * (3 * 64 + 1)
        ws = torch.empty(n, dtype=torch.uint8, device=DEV)
        keep = (x, dx, lnw, lnb, w, b, prob, dprob, g, ws)
        if fwd:
            return "eel_head_fwd", (ptr(x), ptr(lnw), ptr(lnb), ptr(w), ptr(b), ptr(prob), B, S * S, 1, 1, st()), keep
        return "eel_head_bwd", (ptr(x), ptr(lnw), ptr(lnb), ptr(w), ptr(b), ptr(prob), ptr(dprob), ptr(dx), ptr(g[0]), ptr(g[1]), ptr(g[2]),
                                ptr(g[3]), B, S * S, 1, ptr(ws), n, 1, st()), keep
    add("head", "head_fwd %dx%d" % (S, S), lambda: mk_head(True))
    add("head", "head_bwd %dx%d" % (S, S), lambda: mk_head(False))
    for (s, c) in [(S, 64), (S // 2, 128)]:
        def mk_hft(s=s, c=c, fwd=True):
            x, y = rnd(B, s, s, c), torch.empty(B, s, s, c, device=DEV, dtype=BF16)
            ph = torch.empty(_lib.lib.eel_hft_phase_elems(B, s, s, c, 20, 1), device=DEV, dtype=BF16)
            n = _lib.lib.eel_hft_workspace_bytes(B, s, s, c, 20)
            ws = torch.empty(n, dtype=torch.uint8, device=DEV)
            call("eel_hft_fwd", ptr(x), ptr(y), ptr(ph), B, s, s, c, 20, ptr(ws), n, 1, st())     # a valid phase for the backward
            if fwd:
                return "eel_hft_fwd", (ptr(x), ptr(y), ptr(ph), B, s, s, c, 20, ptr(ws), n, 1, st()), (x, y, ph, ws)
            return "eel_hft_bwd", (ptr(x), ptr(ph), ptr(y), B, s, s, c, 20, ptr(ws), n, 1, st()), (x, y, ph, ws)
        add("hft", "hft_fwd %dx%d C=%d" % (s, s, c), mk_hft)
        add("hft", "hft_bwd %dx%d C=%d" % (s, s, c), lambda s=s, c=c: mk_hft(s, c, False))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    only = [o for o in a.only.split(",") if o]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    print("%-34s %9s %9s %9s" % ("case", "us", "TFLOP/s", "GB/s"))
    for group, name, mk in cases(a.batch, a.size, only):
        fn, args, keep = mk()
        if a.once:
            call(fn, *args)
            torch.cuda.synchronize()
            continue
        call(fn, *args)
        call(fn, *args)
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call(fn, *args)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        f, b = profiling.cost(fn, args)
        print("%-34s %9.1f %9.1f %9.1f" % (name, ms * 1e3, f / ms / 1e9, b / ms / 1e6))
        del keep


if __name__ == "__main__":
    main()
