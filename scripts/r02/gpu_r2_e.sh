#!/bin/bash
# Round 2, GPU call E: lean lane refill + classify pass (phased block wavefront) on the tree scenes.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2e; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -q -x --timeout 120 2>&1 | tail -5 | tee $OUT/pytest.txt
run() { tag=$1; shift; echo -n "$tag: "; timeout 60 mort_b200/mort "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-130; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S1" "$SF"; do
  echo "=== $cfg"
  for shape in "1024 1 2048" "640 1 2048" "512 1 2048" "512 2 1024"; do
    set -- $shape
    run sync_${1}x${2}_${3} $cfg --frames 2 --mode pool --pool-sync --tpb $1 --bps $2 --pool $3
    for rf in 4 8 12 16 24; do run sync_${1}x${2}_${3}_rf$rf $cfg --frames 2 --mode pool --pool-sync --tpb $1 --bps $2 --pool $3 --refill $rf; done
  done
done
