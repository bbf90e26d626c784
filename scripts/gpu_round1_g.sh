#!/bin/bash
# GPU call 7: quad pre-filter + smem accumulators; new field / scene-file tests; 4K configs; ncu of the bench config.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1g; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== cli defaults"
for s in 1 5 6 7 8 9; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
echo "== variants"
for bps in 4 6 8; do
  timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 1 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
done
echo "== 4K (config 5, reduced spp)"
for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --width 3840 --aspect 1.7777778 --spp 64 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_4k.jsonl; done
echo "== ncu full on the bench config"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/plain_for_ncu.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_bench python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -6
