#!/bin/bash
# Round 2, GPU call L (closing): whole GPU suite, smoke(), both bench arms, ncu launch list + one --set full capture of the bench
# launch itself (roofline.traffic), the leaf-postponing A/B build, CLI lines of the ten scenes and of BASELINE configs 1-4.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2l; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=mort_b200/mort
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-900; tail -3 $OUT/bench_mort.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-500; tail -3 $OUT/bench_ref.err
echo "== postpone A/B + CLI lines"
for b in mort_b200/mort ab_postpone/mort; do
  for args in "8 --width 800 --spp 256 --depth 40 --frames 2" "1 --frames 3" "1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50 --frames 2" "6 --width 600 --spp 256 --frames 2"; do
    timeout 120 $b $args 2>&1 | tail -1 | cut -c1-240 | tee -a $OUT/final_ab.jsonl; echo "  # $b :: $args" | tee -a $OUT/final_ab.jsonl
  done
done
echo "== scene defaults"; for s in 1 2 3 4 5 6 7 8 9 10; do timeout 120 $M $s --frames 2 2>&1 | tail -1 | cut -c1-200 | tee -a $OUT/cli_defaults.jsonl; done
echo "== BASELINE configs"
timeout 120 $M 1 --width 400 --aspect 1.7777778 --spp 32 --depth 50 --frames 20 2>&1 | tail -1 | cut -c1-200 | tee -a $OUT/cli_configs.jsonl
timeout 120 $M 6 --width 600 --spp 1024 --depth 50 --frames 2 2>&1 | tail -1 | cut -c1-200 | tee -a $OUT/cli_configs.jsonl
timeout 120 $M 8 --width 800 --spp 1024 --depth 40 --frames 2 2>&1 | tail -1 | cut -c1-200 | tee -a $OUT/cli_configs.jsonl
timeout 120 $M 1 --field 500 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/cli_configs.jsonl
timeout 120 $M 1 --field 500 --fieldcam 1 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | cut -c1-330 | tee -a $OUT/cli_configs.jsonl
echo "== ncu launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > $OUT/ncu_list.log 2>&1
tail -2 $OUT/ncu_list.log | cut -c1-200
echo "== ncu full of the bench launch (1024-spp pass)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pool_kernel -s 3 -c 1 -o $OUT/pool_bench python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config > $OUT/ncu_full.log 2>&1
ncu -i $OUT/pool_bench.ncu-rep --page raw --csv > $OUT/ncu_pool_bench_raw.csv 2>/dev/null
rm -f $OUT/pool_bench.ncu-rep
ls -la $OUT | tail -16
