#!/bin/bash
# round-2 ncu evidence (never a bench number: timings under ncu are cold-cache and serialised).  The .ncu-rep files stay on the
# box: only the condensed CSVs (tools/ncu_summary.py, tools/ncu_aggregate.py) come back.
#   bash tools/gpu_ncu_evidence.sh [all]      "all" also re-captures the HFT / Canny / BatchNorm kernels
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --clock-control none"
BENCH="python bench.py --no-cpu-baseline --no-eager-baseline --clock-sampler none"
sum() { ncu -i /tmp/ncu/$1.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/$1.csv; wc -l gpurun_out/$1.csv; }
# 1. launch list of two bench steps (after the bench itself has run clean)
timeout 300 $BENCH --steps 2 --warmup 1 > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv $BENCH --steps 2 --warmup 1 > gpurun_out/r02_ncu_bench.log 2>&1
python tools/ncu_aggregate.py launches gpurun_out/r02_launches_bench.csv > gpurun_out/r02_launches_bench_b64_agg.csv; head -8 gpurun_out/r02_launches_bench_b64_agg.csv
# 2. full captures, condensed
timeout 900 $NCU --set full -k regex:"tc_conv_kernel" -c 36 -o /tmp/ncu/r02_tc_conv_ncu_full_b64 -f $BENCH --steps 1 --warmup 0 > /dev/null 2>&1; sum r02_tc_conv_ncu_full_b64
python tools/ncu_aggregate.py traffic gpurun_out/r02_tc_conv_ncu_full_b64.csv tc_conv3x3 > gpurun_out/r02_traffic.json
timeout 900 $NCU --set full -k regex:"add_interleave_bwd_bn|se_fwd_coop|se_bwd_coop|head_bwd_warp|head_fwd_tp|bn_act_bwd_kernel" -c 40 -o /tmp/ncu/r02_round2_kernels_ncu_full -f $BENCH --steps 1 --warmup 0 > /dev/null 2>&1; sum r02_round2_kernels_ncu_full
if [ "$1" = "all" ]; then
  timeout 600 $NCU --set full -k regex:"hft_tc" -o /tmp/ncu/r02_hft_ncu_full -f python tools/op_bench.py --only hft --once > /dev/null 2>&1; sum r02_hft_ncu_full
  timeout 300 $NCU --set full -k regex:"canny" -o /tmp/ncu/r02_canny_ncu_full -f python tools/canny_bench.py --once > /dev/null 2>&1; sum r02_canny_ncu_full
  timeout 600 $NCU --set full -k regex:"bn_act_bwd|colreduce" -c 12 -o /tmp/ncu/r02_bn_bwd_ncu_full -f python tools/op_bench.py --only bn --once > /dev/null 2>&1; sum r02_bn_bwd_ncu_full
fi
du -sh gpurun_out
