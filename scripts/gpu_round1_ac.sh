#!/bin/bash
# GPU call 32: closing validation after the fuzz-driven flattening changes, 7-op instances and Q39.24 sums: GPU suite, smoke, both bench
# arms, occupancy variants on the tree scenes, and a fresh source-level capture of scene 8 (after the Perlin change).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1ac; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-200
timeout 900 python bench.py --impl reference 2>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-160
echo "== occupancy variants"
for s in 8 1; do for b in 4 6 8; do timeout 300 mort_b200/mort $s --frames 3 --bps $b 2>&1 | tail -1 | tee -a $OUT/variants.jsonl | cut -c1-150; done; done
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 1 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
echo "== ncu full scene 8 (256 spp)"
timeout 300 python bench.py --steps 1 --warmup 1 --scene 8 --width 800 --spp 256 --depth 40 --no-cpu-baseline > $OUT/plain8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_scene8_b python bench.py --steps 1 --warmup 1 --scene 8 --width 800 --spp 256 --depth 40 --no-cpu-baseline > $OUT/ncu8.log 2>&1
cp mort_b200/libmort_b200.so $OUT/libmort_b200.so
ls -la $OUT | tail -4
