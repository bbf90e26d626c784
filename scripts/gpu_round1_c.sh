#!/bin/bash
# GPU call 3: re-test after the kernel restructuring (while-while traversal, block-aligned Philox, cold paths
# out of line), wavefront parity + timing, ncu of both variants.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1c; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 3 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== bench mort wave"; timeout 600 python bench.py --steps 2 --warmup 1 --mode wave --no-cpu-baseline 2>$OUT/bench_wave.err | tee $OUT/bench_wave.json; tail -3 $OUT/bench_wave.err
echo "== cli"
for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
for s in 1 6 8; do timeout 300 mort_b200/mort $s --frames 2 --mode wave 2>&1 | tail -1 | tee -a $OUT/cli_wave.jsonl; done
timeout 120 mort_b200/mort 1 --width 400 --spp 32 --depth 50 --frames 5 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 8 --width 800 --spp 1024 --depth 40 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 8 --width 800 --spp 1024 --depth 40 --frames 2 --stage 100000 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 1 --frames 2 --stage 100000 | tail -1 | tee -a $OUT/cli_configs.jsonl
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain_for_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full mega"
timeout 300 python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/plain_for_ncu2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_cornell python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -20
