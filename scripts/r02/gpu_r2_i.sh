#!/bin/bash
# Round 2, GPU call I: trace/classify overlap on (xflags 0) vs off (xflags 1), same build; ab_base = previous commit.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2i; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -q -x --timeout 120 2>&1 | tail -4 | tee $OUT/pytest.txt
run() { tag=$1; bin=$2; shift; shift; echo -n "$tag: "; timeout 60 $bin "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-130; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S6="6 --width 600 --spp 256 --depth 50"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S1" "$SF" "$S6"; do
  echo "=== $cfg"
  run base_640x1_2048 ab_base/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 16
  run base_512x2_1024 ab_base/mort $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024 --refill 16
  for x in 0 1; do
    run new_x${x}_640x1_2048 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 16 --xflags $x
    run new_x${x}_640x1_1792 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 1792 --refill 16 --xflags $x
    run new_x${x}_512x2_1024 mort_b200/mort $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024 --refill 16 --xflags $x
    run new_x${x}_384x2_1152 mort_b200/mort $cfg --frames 2 --mode pool --tpb 384 --bps 2 --pool 1152 --refill 16 --xflags $x
    run new_x${x}_768x1_2048 mort_b200/mort $cfg --frames 2 --mode pool --tpb 768 --bps 1 --pool 2048 --refill 16 --xflags $x
  done
  run new_x1_640x1_2048_norefill mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill -1 --xflags 1
  run new_mega mort_b200/mort $cfg --frames 2
done
