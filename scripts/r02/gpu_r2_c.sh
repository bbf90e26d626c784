#!/bin/bash
# Round 2, GPU call C: barrier-free block wavefront (pool_async_kernel): parity tests, then A/B against the phased form and the megakernel.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2c; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_pool.py tests/test_gpu_group.py -m gpu -q -x --timeout 120 2>&1 | tail -8 | tee $OUT/pytest.txt
run() { tag=$1; shift; echo -n "$tag: "; timeout 60 mort_b200/mort "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-150; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S6="6 --width 600 --spp 256 --depth 50"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S6" "$S1" "$SF"; do
  echo "=== $cfg"
  run mega $cfg --frames 2
  run sync_1024x1_2048 $cfg --frames 2 --mode pool --pool-sync --tpb 1024 --bps 1 --pool 2048
  run sync_512x2_1024 $cfg --frames 2 --mode pool --pool-sync --tpb 512 --bps 2 --pool 1024
  for shape in "1024 1 2048" "768 1 2048" "640 1 2048" "512 1 2048" "512 1 2560" "384 2 1024" "512 2 1024" "256 2 1024" "256 3 768" "640 1 2560"; do
    set -- $shape
    run async_${1}x${2}_${3} $cfg --frames 2 --mode pool --tpb $1 --bps $2 --pool $3
  done
done
