#!/bin/bash
# Round 2, GPU call J: the whole GPU suite, smoke(), both bench arms.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2j; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-1500; tail -3 $OUT/bench_mort.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-600; tail -3 $OUT/bench_ref.err
