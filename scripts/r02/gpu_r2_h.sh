#!/bin/bash
# Round 2, GPU call H: traced-list overlap of the trace drain with classify, one generic kernel again, settle out of line, explicit fused
# shading forms; A/B builds: ab_base (previous commit), ab_stacktop (-DMORT_STACK_TOP), ab_settleinl (-DMORT_SETTLE_INLINE).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2h; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -q -x --timeout 120 2>&1 | tail -4 | tee $OUT/pytest.txt
run() { tag=$1; bin=$2; shift; shift; echo -n "$tag: "; timeout 60 $bin "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-130; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S6="6 --width 600 --spp 256 --depth 50"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S1" "$SF" "$S6"; do
  echo "=== $cfg"
  for b in ab_base mort_b200 ab_stacktop ab_settleinl; do
    run "${b}_640x1_2048" $b/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 16
    run "${b}_512x2_1024" $b/mort $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024 --refill 16
    run "${b}_384x2_1280" $b/mort $cfg --frames 2 --mode pool --tpb 384 --bps 2 --pool 1280 --refill 16
    run "${b}_mega" $b/mort $cfg --frames 2
  done
  run new_1024x1_2048 mort_b200/mort $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 2048 --refill 16
  run new_768x1_2048 mort_b200/mort $cfg --frames 2 --mode pool --tpb 768 --bps 1 --pool 2048 --refill 16
  run new_512x1_2048 mort_b200/mort $cfg --frames 2 --mode pool --tpb 512 --bps 1 --pool 2048 --refill 16
  run new_640x1_2304 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2304 --refill 16
  run new_640x1_1792 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 1792 --refill 16
  run new_640x1_2048_rf8 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 8
  run new_640x1_2048_rf24 mort_b200/mort $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 24
done
