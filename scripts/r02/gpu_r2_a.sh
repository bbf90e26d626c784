#!/bin/bash
# Round 2, GPU call A: first run of the block wavefront (pool.cu): parity tests, then same-box A/B against the megakernel
# over block shapes and pool sizes on the BASELINE configs' scenes.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2a; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== diag"; timeout 300 python scripts/diag_pool_vs_mega.py 2>&1 | tail -12
M=mort_b200/mort
run() { tag=$1; shift; echo -n "$tag: "; timeout 120 $M "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-200; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
for cfg in "8 --width 800 --spp 256 --depth 40" "6 --width 600 --spp 256 --depth 50" "1" "1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"; do
  echo "=== $cfg"
  run mega $cfg --frames 2
  run pool_512x2_1024 $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024
  run pool_512x2_1280 $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1280
  run pool_1024x1_2048 $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 2048
  run pool_1024x1_2560 $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 2560
  run pool_384x2_1024 $cfg --frames 2 --mode pool --tpb 384 --bps 2 --pool 1024
  run pool_256x3_768 $cfg --frames 2 --mode pool --tpb 256 --bps 3 --pool 768
  run pool_256x2_1024 $cfg --frames 2 --mode pool --tpb 256 --bps 2 --pool 1024
  run pool_512x1_2048 $cfg --frames 2 --mode pool --tpb 512 --bps 1 --pool 2048
  for rf in 4 8 16 24; do run pool_512x2_1024_refill$rf $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024 --refill $rf; done
  run pool_256x2_1024_refill8 $cfg --frames 2 --mode pool --tpb 256 --bps 2 --pool 1024 --refill 8
done
