#!/bin/bash
# round-2 ncu evidence (never a bench number: timings under ncu are cold-cache and serialised).  The .ncu-rep files stay on the
# box: only the condensed CSVs (tools/ncu_summary.py) come back.
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
sum() { ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/$1.csv; wc -l gpurun_out/$1.csv; }
# 1. launch list of one bench step
timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# 2. full captures, condensed
timeout 600 $NCU --set full -k regex:"hft_tc" -o /tmp/ncu/r02_hft_ncu_full -f python tools/op_bench.py --only hft --once > /dev/null 2>&1; sum r02_hft_ncu_full
timeout 300 $NCU --set full -k regex:"canny" -o /tmp/ncu/r02_canny_ncu_full -f python tools/canny_bench.py --once > /dev/null 2>&1; sum r02_canny_ncu_full
timeout 900 $NCU --set full -k regex:"tc_conv_kernel" -c 36 -o /tmp/ncu/r02_tc_conv_ncu_full_b64 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-eager-baseline > /dev/null 2>&1; sum r02_tc_conv_ncu_full_b64
timeout 600 $NCU --set full -k regex:"bn_act_bwd|colreduce" -c 12 -o /tmp/ncu/r02_bn_bwd_ncu_full -f python tools/op_bench.py --only bn --once > /dev/null 2>&1; sum r02_bn_bwd_ncu_full
du -sh gpurun_out
