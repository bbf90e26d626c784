#!/bin/bash
# GPU call 12: tile split + two-pass custom scene tests; ncu source-level profile of the current bench kernel.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1k; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
for s in 1 8; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
cp mort_b200/libmort_b200.so $OUT/libmort_b200.so
echo "== ncu full on the bench config (256 spp)"
timeout 300 python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/plain_for_ncu.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_cornell python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "== ncu full scene 1"
timeout 300 python bench.py --steps 1 --warmup 1 --scene 1 --width 1200 --aspect 1.7777778 --spp 100 --depth 20 --no-cpu-baseline > $OUT/plain_for_ncu1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_scene1 python bench.py --steps 1 --warmup 1 --scene 1 --width 1200 --aspect 1.7777778 --spp 100 --depth 20 --no-cpu-baseline > $OUT/ncu_full1.log 2>&1
ls -la $OUT | tail -5
