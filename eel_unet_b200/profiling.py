"""Algorithmic cost model of the C-ABI calls + aggregation of CUDA-event timings (bench.py roofline).

``cost(name, args)`` returns (flops, bytes): FLOPs of the mathematical op (2 per multiply-add) and the
ALGORITHMIC bytes -- every input tensor read once and every output written once at its storage dtype
(DESIGN.md "Kernels and rooflines").  Re-reads a kernel actually makes (e.g. the BatchNorm backward's
two passes) are deliberately not counted: they show up as a lower roofline fraction.
"""
from collections import defaultdict


def _esz(dtype):
    return 2 if dtype == 1 else 4


def cost(name, a):
    n = name[4:] if name.startswith("eel_") else name
    if n == "conv3x3_fwd":
        N, H, W, ci, co, dt = a[4], a[5], a[6], a[7], a[8], a[11]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + co) * _esz(dt) + 9 * ci * co * _esz(dt)
    if n == "tc_conv3x3":
        N, H, W, ci, co = a[4], a[5], a[6], a[7], a[8]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + co) * 2 + 9 * ci * co * 2
    if n == "tc_conv3x3_dgrad_bnsums":     # the data-gradient launch + one read of the BatchNorm input in its epilogue
        N, H, W, ci, co = a[3], a[4], a[5], a[6], a[7]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + 2 * co) * 2 + 9 * ci * co * 2
    if n == "tc_conv3x3_2src":
        N, H, W, c1, c2, co = a[5], a[6], a[7], a[8], a[9], a[10]
        P = N * H * W
        return 2.0 * P * 9 * (c1 + c2) * co, P * (c1 + c2 + co) * 2 + 9 * (c1 + c2) * co * 2
    if n == "bn_add_fwd":
        P, C, dt = a[3], a[4], a[9]
        return 3.0 * P * C, 3 * P * C * _esz(dt)
    if n == "tc_conv3x3_dgrad_split":      # the data-gradient launch (two output tensors) + one read of the BatchNorm input of the first half
        N, H, W, ci, co = a[4], a[5], a[6], a[7], a[8]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + co + (co // 2 if a[9] else 0)) * 2 + 9 * ci * co * 2
    if n == "bn_relu_pool_fwd":
        N, H, W, C, dt = a[8], a[9], a[10], a[11], a[12]
        return 6.0 * N * H * W * C, 2.25 * N * H * W * C * _esz(dt)
    if n == "bn_relu_pool_bwd":
        N, H, W, C, dt = a[12], a[13], a[14], a[15], a[19]
        return 16.0 * N * H * W * C, 3.25 * N * H * W * C * _esz(dt)
    if n == "gelu_bwd_colsum":
        return 12.0 * a[4], 3 * a[4] * _esz(a[6])
    if n == "tc_linear":
        P, K, No = a[4], a[5], a[6]
        return 2.0 * P * K * No, P * (K + No) * 2 + K * No * 2
    if n == "tc_capmlp_fwd":
        P, C = a[8], a[9]
        return 2.0 * P * (64 * 256 + 256 * C), P * (64 + (512 if a[5] else 0) + C) * 2 + (64 * 256 + 256 * C) * 2
    if n == "tc_convt2x2_fwd":
        N, h, w, ci, co = a[4], a[5], a[6], a[7], a[8]
        P = N * h * w
        return 2.0 * P * ci * 4 * co, P * (ci + 4 * co) * 2 + 4 * ci * co * 2
    if n == "tc_convt2x2_dgrad":
        N, h, w, ci, co = a[3], a[4], a[5], a[6], a[7]
        P = N * h * w
        return 2.0 * P * ci * 4 * co, P * (ci + 4 * co) * 2 + 4 * ci * co * 2
    if n in ("tc_conv3x3_wgrad",):
        N, H, W, ci, co = a[3], a[4], a[5], a[6], a[7]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + co) * 2 + 9 * ci * co * 4
    if n == "tc_wgrad":
        P, Ma, Nb = a[3], a[4], a[5]
        return 2.0 * P * Ma * Nb, P * (Ma + Nb) * 2 + Ma * Nb * 4
    if n == "shift_channels":
        tot = a[2] * a[3] * a[4] * a[5]
        return 0.0, 2 * tot * _esz(a[7])
    if n == "conv3x3_wgrad":
        N, H, W, ci, co, dt = a[3], a[4], a[5], a[6], a[7], a[8]
        P = N * H * W
        return 2.0 * P * 9 * ci * co, P * (ci + co) * _esz(dt) + 9 * ci * co * 4
    if n in ("convt2x2_fwd",):
        N, h, w, ci, co, dt = a[4], a[5], a[6], a[7], a[8], a[9]
        P = N * h * w
        return 2.0 * P * ci * 4 * co, P * (ci + 4 * co) * _esz(dt) + 4 * ci * co * _esz(dt)
    if n in ("convt2x2_dgrad", "convt2x2_wgrad"):
        N, h, w, ci, co, dt = a[3], a[4], a[5], a[6], a[7], a[8]
        P = N * h * w
        return 2.0 * P * ci * 4 * co, P * (ci + 4 * co) * _esz(dt) + 4 * ci * co * 4
    if n == "linear_fwd":
        P, K, No, dt = a[4], a[5], a[6], a[9]
        return 2.0 * P * K * No, P * (K + No) * _esz(dt) + K * No * _esz(dt)
    if n in ("linear_dgrad", "linear_wgrad"):
        P, K, No, dt = a[3], a[4], a[5], a[8]
        return 2.0 * P * K * No, P * (K + No) * _esz(dt) + K * No * 4
    if n == "bn_stats":
        P, C, dt = a[1], a[2], a[11]
        return 3.0 * P * C, P * C * _esz(dt)
    if n == "bn_act_fwd":
        P, C, dt = a[6], a[7], a[9]
        return 4.0 * P * C, 2 * P * C * _esz(dt)
    if n == "bn_act_shift_fwd":
        N, H, W, C, dt = a[6], a[7], a[8], a[9], a[11]
        return 4.0 * N * H * W * C, 2 * N * H * W * C * _esz(dt)
    if n == "bn_act_bwd":
        P, C, dt = a[10], a[11], a[16]
        return 12.0 * P * C, 3 * P * C * _esz(dt)
    if n in ("maxpool2_fwd", "maxpool2_bwd"):
        if n == "maxpool2_fwd":
            N, H, W, C, dt = a[2], a[3], a[4], a[5], a[6]
            return 0.75 * N * H * W * C, 1.25 * N * H * W * C * _esz(dt)
        N, H, W, C, dt = a[3], a[4], a[5], a[6], a[7]
        return 0.75 * N * H * W * C, 2.25 * N * H * W * C * _esz(dt)
    if n == "add_interleave_fwd":
        P, C, dt = a[4], a[5], a[10]
        return 1.0 * P * C, 5 * P * C * _esz(dt)
    if n == "add_interleave_bwd":
        P, C, dt = a[3], a[4], a[5]
        return 0.0, 4 * P * C * _esz(dt)
    if n == "add_interleave_bwd_bnsums":   # + one read of each BatchNorm input whose backward sums ride on the pass
        P, C, dt = a[3], a[4], a[21]
        nbn = 2 if a[12] else 1
        return 4.0 * nbn * P * C, (4 + nbn) * P * C * _esz(dt)
    if n == "pgr_fwd":
        P, C, dt = a[5], a[6], a[7]
        return 4.0 * P * C, 2 * P * C * _esz(dt) + 4 * P
    if n == "pgr_bwd":
        P, C, dt = a[8], a[9], a[12]
        return 8.0 * P * C, 3 * P * C * _esz(dt) + 8 * P
    if n == "bn_pgr_fwd":
        P, C, dt = a[9], a[10], a[11]
        return 8.0 * P * C, 2 * P * C * _esz(dt) + 4 * P
    if n == "bn_pgr_bwd":
        P, C, dt = a[13], a[14], a[17]
        return 14.0 * P * C, 3 * P * C * _esz(dt) + 8 * P
    if n == "bn_act_bwd_apply":
        P, C, dt = a[9], a[10], a[13]
        return 8.0 * P * C, 3 * P * C * _esz(dt)
    if n == "head_fwd":
        N, HW, O, dt = a[6], a[7], a[8], a[9]
        return (10.0 + 2 * O) * N * HW * 64, N * HW * (64 * _esz(dt) + 4 * O)
    if n == "head_bwd":
        N, HW, O, dt = a[12], a[13], a[14], a[17]
        return (30.0 + 6 * O) * N * HW * 64, N * HW * (2 * 64 * _esz(dt) + 8 * O)
    if n == "se_fwd":
        N, HW, C, dt = a[9], a[10], a[11], a[15]
        return 2.0 * N * HW * C, 2 * N * HW * C * _esz(dt)
    if n == "se_bwd":
        N, HW, C, dt = a[13], a[14], a[15], a[19]
        return 4.0 * N * HW * C, 3 * N * HW * C * _esz(dt)
    if n in ("gelu_fwd", "relu_fwd"):
        return 8.0 * a[2], 2 * a[2] * _esz(a[3])
    if n in ("gelu_bwd", "relu_bwd"):
        return 12.0 * a[3], 3 * a[3] * _esz(a[4])
    if n in ("hft_fwd", "hft_bwd"):
        if n == "hft_fwd":
            N, H, W, C, r, dt = a[3], a[4], a[5], a[6], a[7], a[10]
            io = 4
        else:
            N, H, W, C, r, dt = a[3], a[4], a[5], a[6], a[7], a[10]
            io = 4
        F = 2 * min(r, H // 2, W // 2)
        k1 = W if n == "hft_fwd" else 2 * W
        m4 = 2 * W if n == "hft_fwd" else W
        macs = N * C * (2 * F * k1 * H + 2 * F * 2 * H * F + 2 * H * 2 * F * F + m4 * 2 * F * H)
        return 2.0 * macs, io * N * H * W * C * _esz(dt)
    if n == "edge_loss_fwd":
        N, H, W = a[2], a[3], a[4]
        return 40.0 * N * H * W, 4 * N * H * W * (1 + 2.332)
    if n == "edge_loss_bwd":
        N, H, W = a[5], a[6], a[7]
        return 40.0 * N * H * W, 4 * N * H * W * (1 + 2 * 2.332)
    if n == "permute4":
        tot = a[4] * a[5] * a[6] * a[7]
        return 0.0, tot * (_esz(a[1]) + _esz(a[3]))
    if n == "colsum":
        return 1.0 * a[2] * a[3], a[2] * a[3] * _esz(a[6])
    if n == "nchw_to_nhwc":
        tot = a[2] * a[3] * a[4] * a[5]
        return 0.0, tot * (4 + _esz(a[6]))
    if n == "adam_step":
        return 12.0 * a[4], 28 * a[4]
    if n in ("canny_rgb", "canny_gray"):
        px = a[2] * a[3] * a[4]
        return 40.0 * px, px * ((3 if n == "canny_rgb" else 1) + 1)
    return 0.0, 0.0


# launches reported under another family's name
ALIAS = {"tc_conv3x3_dgrad_bnsums": "tc_conv3x3", "tc_conv3x3_dgrad_split": "tc_conv3x3", "tc_conv3x3_2src": "tc_conv3x3", "bn_add_fwd": "add_interleave_fwd", "gelu_bwd_colsum": "gelu_bwd", "add_interleave_bwd_bnsums": "add_interleave_bwd", "bn_act_shift_fwd": "bn_act_fwd"}

GEMM_CLASS = {"tc_conv3x3", "tc_linear", "tc_capmlp_fwd", "tc_convt2x2_fwd", "tc_convt2x2_dgrad", "tc_conv3x3_wgrad", "tc_wgrad", "conv3x3_fwd", "conv3x3_wgrad", "convt2x2_fwd", "convt2x2_dgrad", "convt2x2_wgrad", "linear_fwd",
              "linear_dgrad", "linear_wgrad", "hft_fwd", "hft_bwd"}


def summarize(records):
    """records from _lib.set_profiler -> {family: dict(calls, ms, flops, bytes)} (events must be complete)."""
    fam = defaultdict(lambda: {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
    for name, args, s, e in records:
        f, b = cost(name, args)
        k = ALIAS.get(name[4:], name[4:])
        d = fam[k]
        d["calls"] += 1
        d["ms"] += s.elapsed_time(e)
        d["flops"] += f
        d["bytes"] += b
    return dict(fam)


def table(fam, hbm_gbs, tensor_tflops):
    total = sum(d["ms"] for d in fam.values()) or 1.0
    rows = []
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
        ms = d["ms"] or 1e-9
        tf = d["flops"] / ms / 1e9
        gb = d["bytes"] / ms / 1e6
        rows.append((k, d["calls"], d["ms"], 100.0 * d["ms"] / total, tf, 100.0 * tf / tensor_tflops, gb, 100.0 * gb / hbm_gbs))
    return rows


def shape_table(records):
    """per (kernel, integer arguments) rows: calls, ms, TFLOP/s, GB/s -- finds the slow SHAPES inside a family."""
    acc = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for name, args, s, e in records:
        dims = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool) and 0 <= a < (1 << 28))
        f, b = cost(name, args)
        d = acc[(name[4:], dims)]
        d[0] += 1
        d[1] += s.elapsed_time(e)
        d[2] += f
        d[3] += b
    rows = []
    for (k, dims), (calls, ms, f, b) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        ms = ms or 1e-9
        rows.append((k, "x".join(str(v) for v in dims), calls, ms, f / ms / 1e9, b / ms / 1e6))
    return rows
