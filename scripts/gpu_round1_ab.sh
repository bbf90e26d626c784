#!/bin/bash
# GPU call 30: GPU suite with the fuzz parity tests and the flattening changes (wrapper boxes, degenerate / inconsistent reference-BVH boxes); bench.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1ab; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== fuzz"; timeout 600 python -m pytest tests/test_gpu_fuzz.py -q --timeout 300 2>&1 | tail -15
echo "== pytest gpu"; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-200
for s in 1 8 10; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | cut -c1-140; done
