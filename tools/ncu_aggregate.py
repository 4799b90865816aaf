#!/usr/bin/env python
"""Condense the round's ncu outputs into the files bench.py and profiles/README.md cite.

    python tools/ncu_aggregate.py launches gpurun_out/r02_launches_bench.csv > profiles/r02_launches_bench_b64_agg.csv
        per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (kernel, launches, total_us, share_pct)
    python tools/ncu_aggregate.py traffic profiles/r02_tc_conv_ncu_full_b64.csv tc_conv3x3 > profiles/r02_traffic.json
        DRAM read + write bytes per launch of one kernel family from a condensed `ncu --set full` capture (tools/ncu_summary.py)
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("eel::", "")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) >= 15 and r[0].isdigit()]
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[12] != "gpu__time_duration.sum":
            continue
        v = float(r[14].replace(",", ""))
        us = v / 1e3 if r[13] in ("nsecond", "ns") else (v * 1e3 if r[13] in ("msecond", "ms") else v)
        t = tot[short(r[4])]
        t[0] += 1
        t[1] += us
    allus = sum(v[1] for v in tot.values()) or 1.0
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "share_pct"])
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, "%.1f" % us, "%.2f" % (100.0 * us / allus)])


def unit_scale(header, what):
    m = re.search(r"\[(\w+)\]", header)
    u = m.group(1) if m else "byte"
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def traffic(path, family):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ir = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum"))
    iw = next(i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum"))
    sr, sw = unit_scale(hdr[ir], "r"), unit_scale(hdr[iw], "w")
    body = rows[1:]
    total = sum(float(r[ir]) * sr + float(r[iw]) * sw for r in body)
    json.dump({family: {"launches": len(body), "dram_bytes_total": total, "dram_bytes_per_launch": total / max(1, len(body)),
                        "source": "%s (ncu --set full --clock-control none, the %d %s launches of one bench.py step, batch 64 at 256^2)"
                                  % (path, len(body), family)}}, sys.stdout, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        traffic(sys.argv[2], sys.argv[3])
