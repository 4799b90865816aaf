#!/bin/bash
# BASELINE config 4 (bf16 data-parallel training at 512^2, 64 images per GPU) at 1 / 2 / 4 / 8 GPUs of one box, plus fp32-mode
# and 512^2 single-GPU numbers with the eager baseline.  N = 1, 2, 4 run side by side on disjoint GPUs, then N = 8 alone.
mkdir -p gpurun_out
OUT=gpurun_out/r02_config_check.jsonl
: > $OUT
run() {  # run <visible devices> <nproc> <extra bench args...>
  local dev=$1 n=$2; shift 2
  if [ "$n" = "1" ]; then
    CUDA_VISIBLE_DEVICES=$dev python bench.py --gpus 1 "$@"
  else
    CUDA_VISIBLE_DEVICES=$dev python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"
  fi
}
A="--size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline"
run 0 1 $A > gpurun_out/sc_512_1.json 2> gpurun_out/sc_512_1.err &
run 1,2 2 $A > gpurun_out/sc_512_2.json 2> gpurun_out/sc_512_2.err &
run 3,4,5,6 4 $A > gpurun_out/sc_512_4.json 2> gpurun_out/sc_512_4.err &
wait
run 0,1,2,3,4,5,6,7 8 $A > gpurun_out/sc_512_8.json 2> gpurun_out/sc_512_8.err
run 0,1,2,3,4,5,6,7 8 --size 256 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/sc_256_8.json 2> gpurun_out/sc_256_8.err
# single-GPU extras, side by side: fp32 mode at 256^2, bf16 at 512^2 with the eager baseline, bf16 at 256^2 with both baselines
run 0 1 --size 256 --batch 64 --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sc_256_fp32.json 2> gpurun_out/sc_256_fp32.err &
run 1 1 --size 512 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sc_512_eager.json 2> gpurun_out/sc_512_eager.err &
run 2 1 --size 256 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/sc_256_eager.json 2> gpurun_out/sc_256_eager.err &
wait
for f in sc_512_1 sc_512_2 sc_512_4 sc_512_8 sc_256_8 sc_256_fp32 sc_512_eager sc_256_eager; do
  if [ -s gpurun_out/$f.json ]; then cat gpurun_out/$f.json >> $OUT; else echo "{\"run\": \"$f\", \"error\": \"no output\"}" >> $OUT; tail -5 gpurun_out/$f.err; fi
done
python - <<'PY'
import json
for ln in open("gpurun_out/r02_config_check.jsonl"):
    d = json.loads(ln)
    if "value" in d:
        e = d.get("gpu_eager_baseline") or {}
        print(d["config"]["workload"][:70], "| N", d["n_gpus"], "| %.1f img/s  %.2f ms/step | e2e %.1f | mem %.1f GB" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["peak_mem_gb"]),
              "| eager", {k: round(v.get("value", 0), 1) for k, v in e.items() if isinstance(v, dict)})
    else:
        print(d)
PY
