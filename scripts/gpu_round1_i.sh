#!/bin/bash
# GPU call 10: exact accumulation + bench contract tests; area-sorted linear scan; task-size sweep.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1i; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== cli"
for s in 1 5 6 7 8 9; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 --bps 8 | tail -1 | tee -a $OUT/cli_defaults.jsonl
echo "== task size sweep (cornell 1024 spp, scene 1 default, cfg1)"
for ts in 512 1024 4096 8192; do
  MORT_TASK_SAMPLES=$ts timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 | tail -1 | sed "s/^{/{\"task_samples\":$ts,/" | tee -a $OUT/task_sweep.jsonl
  MORT_TASK_SAMPLES=$ts timeout 120 mort_b200/mort 1 --frames 3 | tail -1 | sed "s/^{/{\"task_samples\":$ts,/" | tee -a $OUT/task_sweep.jsonl
  MORT_TASK_SAMPLES=$ts timeout 120 mort_b200/mort 1 --width 400 --spp 32 --depth 50 --frames 5 | tail -1 | sed "s/^{/{\"task_samples\":$ts,/" | tee -a $OUT/task_sweep.jsonl
done
