"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol eel.h declares."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from eel_unet_b200 import _lib

    protos = _lib.parse_header()
    assert len(protos) >= 40
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name, _, _ in protos:
        assert hasattr(raw, name), "libeel.so lacks %s declared in include/eel.h" % name


def test_no_undeclared_exports():
    from eel_unet_b200 import _lib

    declared = {p[0] for p in _lib.parse_header()}
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T eel_" in ln}
    assert exported == declared, (exported ^ declared)


def test_trivial_calls_without_gpu():
    from eel_unet_b200 import _lib

    assert _lib.lib.eel_version() >= 100
    assert _lib.lib.eel_reduce_workspace_bytes(64, 2) > 0
    assert _lib.lib.eel_canny_workspace_bytes(1, 16, 16) >= 16 + 8 * 256
    assert _lib.lib.eel_hft_workspace_bytes(1, 128, 128, 64, 20) > 0
    # argument validation happens before any CUDA work and reports through eel_last_error
    rc = _lib.lib.eel_gelu_fwd(None, None, 0, 0, None)
    assert rc == -1 and b"gelu_fwd" in _lib.lib.eel_last_error()


def test_no_cpu_fallback():
    import pytest
    import torch

    from eel_unet_b200 import EELUnet, _lib

    m = EELUnet(3, 1)
    with pytest.raises(_lib.EelError):
        m(torch.zeros(1, 3, 32, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eel_unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "/root/reference" not in src, f


def test_reference_arm_of_bench_runs_without_the_product():
    """bench.py --impl reference times the CPU oracle port only: the product package (and so libeel.so) must not even be
    imported into that process, also under torchrun's OMP_NUM_THREADS=1 it must use every host core."""
    import json
    import sys

    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--size", "32"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0
    assert line["native_so_loaded"] == [] and line["product_imported"] is False
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["workload"].startswith("EELUnet bf16 training")        # the same workload string as the measured arm
