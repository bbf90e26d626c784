#!/bin/bash
# GPU call 23: does sampling the clocks with a spawned nvidia-smi slow the timed kernel?  (smi vs in-process NVML vs off);
# source-level ncu capture of scene 8 (BASELINE config 3's scene) for the next round.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1v; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for rep in 1 2; do for how in smi nvml off; do
  MORT_BENCH_CLOCKS=$how timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'clocks_via':'$how','rep':$rep,'value':d['value'],'ms_per_step':d['ms_per_step'],'kernel_ms':d['roofline']['kernel_ms_per_launch'],'e2e':d['e2e']['value'],'clocks':d['clocks']}))" | tee -a $OUT/clock_sampler_ab.jsonl
done; done
timeout 300 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 5 2>&1 | tail -1 | tee -a $OUT/clock_sampler_ab.jsonl
echo "== ncu full scene 8 (256 spp)"
timeout 300 python bench.py --steps 1 --warmup 1 --scene 8 --width 800 --spp 256 --depth 40 --no-cpu-baseline > $OUT/plain8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_scene8 python bench.py --steps 1 --warmup 1 --scene 8 --width 800 --spp 256 --depth 40 --no-cpu-baseline > $OUT/ncu8.log 2>&1
cp mort_b200/libmort_b200.so $OUT/libmort_b200.so
ls -la $OUT | tail -5
