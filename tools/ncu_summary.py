#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of roofline metrics kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/<name>.csv
"""
import csv
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
idx = [(k, hdr.index(k)) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow([k + (" [%s]" % units[i] if units[i] else "") for k, i in idx])
for r in rows[2:]:
    w.writerow([r[i] for _, i in idx])
