#!/bin/bash
# GPU call 28: Perlin with the octave loop rolled and the 8 corners unrolled (scenes 4, 8, 9 against call 27's fully rolled numbers);
# reference arm twice (retry path); final bench line.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1z; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for s in 4 8 9 2 6; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl | cut -c1-140; done
timeout 300 mort_b200/mort 4 --spp 100 --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl | cut -c1-140
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 1 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_render.py -q -x --timeout 600 2>&1 | tail -2
for i in 1 2; do timeout 900 python bench.py --impl reference 2>$OUT/bench_ref_$i.err | tee -a $OUT/bench_reference.jsonl | cut -c1-160; tail -2 $OUT/bench_ref_$i.err | cut -c1-200; done
timeout 600 python bench.py 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-200
