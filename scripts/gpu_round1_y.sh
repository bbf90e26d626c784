#!/bin/bash
# GPU call 27: round-1 closing validation: GPU suite, smoke(), both bench arms, CLI defaults + BASELINE configs with the shipped library.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1y; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-300
timeout 600 python bench.py --impl reference 2>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-200
echo "== CLI defaults"; for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl | cut -c1-140; done
echo "== BASELINE configs"
timeout 300 mort_b200/mort 1 --width 400 --aspect 1.7777778 --spp 32 --depth 50 --frames 20 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
timeout 300 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 1 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
for c in 0 1; do timeout 300 mort_b200/mort 1 --field 500 --fieldcam $c --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140; done
