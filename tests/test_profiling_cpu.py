"""Host logic (no GPU): the roofline cost model reads kernel dimensions out of the C-ABI argument lists by POSITION
(eel_unet_b200/profiling.py).  Pin those positions to the parameter names in include/eel.h so that a changed prototype cannot
silently shift the algorithmic bytes / FLOPs bench.py reports."""
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _header_params():
    src = open(os.path.join(ROOT, "include", "eel.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(eel_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        args = [a.strip() for a in m.group(2).split(",")] if m.group(2).strip() not in ("", "void") else []
        out[m.group(1)] = [re.sub(r"^.*?(\w+)$", r"\1", a) for a in args]
    return out


# kernel -> {position: parameter name} exactly as profiling.cost() indexes them
EXPECTED = {
    "eel_tc_conv3x3": {4: "N", 5: "H", 6: "W", 7: "Cin", 8: "Cout"},
    "eel_tc_conv3x3_dgrad_bnsums": {3: "N", 4: "H", 5: "W", 6: "Cin", 7: "Cout"},
    "eel_tc_conv3x3_dgrad_split": {4: "N", 5: "H", 6: "W", 7: "Cin", 8: "Cout", 9: "z"},
    "eel_tc_conv3x3_2src": {5: "N", 6: "H", 7: "W", 8: "C1", 9: "C2", 10: "Cout"},
    "eel_bn_add_fwd": {3: "P", 4: "C", 9: "dtype"},
    "eel_tc_linear": {4: "P", 5: "K", 6: "Nout"},
    "eel_tc_convt2x2_fwd": {4: "N", 5: "h", 6: "w", 7: "Cin", 8: "Cout"},
    "eel_tc_convt2x2_dgrad": {3: "N", 4: "h", 5: "w", 6: "Cin", 7: "Cout"},
    "eel_tc_conv3x3_wgrad": {3: "N", 4: "H", 5: "W", 6: "Cin", 7: "Cout"},
    "eel_tc_wgrad": {3: "P", 4: "Ma", 5: "Nb"},
    "eel_bn_act_fwd": {6: "P", 7: "C", 9: "dtype"},
    "eel_bn_act_shift_fwd": {6: "N", 7: "H", 8: "W", 9: "C", 11: "dtype"},
    "eel_bn_act_bwd": {10: "P", 11: "C", 16: "dtype"},
    "eel_bn_act_bwd_apply": {9: "P", 10: "C", 13: "dtype"},
    "eel_bn_relu_pool_fwd": {8: "N", 9: "H", 10: "W", 11: "C", 12: "dtype"},
    "eel_bn_relu_pool_bwd": {12: "N", 13: "H", 14: "W", 15: "C", 19: "dtype"},
    "eel_bn_pgr_fwd": {9: "P", 10: "C", 11: "dtype"},
    "eel_bn_pgr_bwd": {13: "P", 14: "C", 17: "dtype"},
    "eel_gelu_bwd_colsum": {4: "n", 6: "dtype"},
    "eel_gelu_bwd": {3: "n", 4: "dtype"},
    "eel_se_fwd": {9: "N", 10: "HW", 11: "C", 15: "dtype"},
    "eel_se_bwd": {13: "N", 14: "HW", 15: "C", 19: "dtype"},
    "eel_add_interleave_fwd": {4: "P", 5: "C", 10: "dtype"},
    "eel_add_interleave_bwd": {3: "P", 4: "C", 5: "dtype"},
    "eel_add_interleave_bwd_bnsums": {3: "P", 4: "C", 12: "z1", 21: "dtype"},
    "eel_head_fwd": {6: "N", 7: "HW", 8: "O", 9: "dtype"},
    "eel_head_bwd": {12: "N", 13: "HW", 14: "O", 17: "dtype"},
    "eel_adam_step": {4: "n"},
    "eel_colsum": {2: "P", 3: "C", 6: "dtype"},
}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_cost_model_argument_positions_match_the_header(name):
    params = _header_params()
    assert name in params, "%s is not declared in include/eel.h" % name
    for pos, want in EXPECTED[name].items():
        assert pos < len(params[name]), (name, pos, params[name])
        assert params[name][pos].lower() == want.lower(), "%s argument %d is %r, the cost model expects %r" % (
            name, pos, params[name][pos], want)


def test_cost_model_covers_the_bench_families():
    """every family bench.py may name as dominant has a non-zero cost"""
    from eel_unet_b200 import profiling

    a = [0] * 24
    args = list(a)
    args[4], args[5], args[6], args[7], args[8] = 2, 16, 16, 64, 64
    f, b = profiling.cost("eel_tc_conv3x3", args)
    assert f == 2.0 * 2 * 16 * 16 * 9 * 64 * 64 and b > 0
    args = list(a)
    args[3], args[4], args[5], args[6], args[7] = 2, 16, 16, 64, 64
    f2, b2 = profiling.cost("eel_tc_conv3x3_dgrad_bnsums", args)
    assert f2 == f and b2 > b          # the same GEMM plus one read of the BatchNorm input
    assert profiling.ALIAS["tc_conv3x3_dgrad_bnsums"] == "tc_conv3x3"
