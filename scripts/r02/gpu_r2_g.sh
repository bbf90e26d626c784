#!/bin/bash
# Round 2, GPU call G: class byte tables + cold-diffuse queue + register-resident stack top (vs the previous build in ab_base/),
# and fewer threads per pool (more rays per lane per trace phase).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2g; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -q -x --timeout 120 2>&1 | tail -4 | tee $OUT/pytest.txt
run() { tag=$1; bin=$2; shift; shift; echo -n "$tag: "; timeout 60 $bin "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-130; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S6="6 --width 600 --spp 256 --depth 50"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S1" "$SF" "$S6"; do
  echo "=== $cfg"
  run base_640x1_2048_rf16 ab_base/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 16
  run base_512x2_1024_rf16 ab_base/mort $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024 --refill 16
  for shape in "1024 1 2048" "640 1 2048" "512 2 1024" "512 1 2048" "448 1 2304" "384 1 2560" "320 1 2560" "256 1 2560" "384 2 1280" "256 2 1280"; do
    set -- $shape
    run new_${1}x${2}_${3}_rf16 mort_b200/mort $cfg --frames 2 --mode pool --tpb $1 --bps $2 --pool $3 --refill 16
  done
  run new_640x1_2048_rf8 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 8
  run new_384x1_2560_rf8 mort_b200/mort $cfg --frames 2 --mode pool --tpb 384 --bps 1 --pool 2560 --refill 8
  run new_mega mort_b200/mort $cfg --frames 2
  run base_mega ab_base/mort $cfg --frames 2
done
