#!/bin/bash
# GPU call 14: progressive accumulation / checkpoint / scene text tests + whole GPU suite; 4-frame progressive timing.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1m; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_progressive.py tests/test_gpu_scene_text.py -q --timeout 300 2>&1 | tail -30 | tee $OUT/pytest_new.txt
echo "== pytest gpu (all)"; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== progressive CLI"; timeout 300 mort_b200/mort 6 --frames 4 --accumulate --spp 256 --checkpoint $OUT/cornell.mckp --out $OUT/cornell_prog.ppm 2>&1 | tail -3 | tee $OUT/cli_progressive.txt; rm -f $OUT/cornell.mckp
timeout 300 mort_b200/mort 6 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_progressive.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-400
