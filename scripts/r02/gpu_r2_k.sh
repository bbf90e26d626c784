#!/bin/bash
# Round 2, GPU call K: first run of the GPU tree builder / refit / motion boxes (tests + timings), the whole GPU suite after the
# builder refactor, config 5's load-balance matrix on one GPU (virtual ranks), ncu --set full of the default kernel on scene 8.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2k; mkdir -p $OUT
export PYTHONUNBUFFERED=1
M=mort_b200/mort
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_build.py -q -s --timeout 300 2>&1 | tail -40 | tee $OUT/pytest_build.txt
echo "== builder timings (CLI)"
for b in host gpu; do for i in 1 2; do
  timeout 120 $M 1 --field 500 --width 320 --spp 4 --depth 8 --builder $b 2>&1 | tail -1 | tee -a $OUT/build_cli.jsonl | cut -c1-420
done; done
timeout 120 $M 1 --field 500 --width 320 --spp 4 --depth 8 --builder gpu --gpu-flags 1 2>&1 | tail -1 | tee -a $OUT/build_cli.jsonl | cut -c1-420
for k in 16 32 128 256; do timeout 120 $M 1 --field 500 --width 320 --spp 4 --depth 8 --builder gpu --gpu-small $k 2>&1 | tail -1 | tee -a $OUT/build_cli.jsonl | cut -c1-420; done
timeout 120 $M 8 --width 200 --spp 4 --builder gpu 2>&1 | tail -1 | tee -a $OUT/build_cli.jsonl | cut -c1-420
echo "== motion boxes A/B"
for mb in "" "--motion-bounds"; do
  timeout 120 $M 1 --frames 3 $mb 2>&1 | tail -1 | tee -a $OUT/motion_ab.jsonl | cut -c1-300
  timeout 120 $M 1 --width 400 --aspect 1.7777778 --spp 25 --depth 50 --frames 5 $mb 2>&1 | tail -1 | tee -a $OUT/motion_ab.jsonl | cut -c1-300
  timeout 120 $M 1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50 --frames 2 $mb 2>&1 | tail -1 | tee -a $OUT/motion_ab.jsonl | cut -c1-300
done
echo "== whole GPU suite"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --deselect tests/test_gpu_build.py 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== config 5 load balance on one GPU (virtual ranks)"
timeout 400 python scripts/r02/virtual_ranks.py --spp 64 --out $OUT/virtual_cfg5.jsonl 2>&1 | cut -c1-400 | tail -12
timeout 200 python scripts/r02/virtual_ranks.py --field 500 --width 1920 --spp 64 --out $OUT/virtual_cfg4.jsonl 2>&1 | cut -c1-400 | tail -3
echo "== ncu full: default kernel on scene 8 (128 spp frame)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pool_kernel -c 1 -o $OUT/pool_s8 $M 8 --width 800 --spp 128 --depth 40 --frames 1 > $OUT/ncu_s8.log 2>&1
ncu -i $OUT/pool_s8.ncu-rep --page raw --csv > $OUT/ncu_pool_s8_raw.csv 2>/dev/null
ls -la $OUT | tail -20
