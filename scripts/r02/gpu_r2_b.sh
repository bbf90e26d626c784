#!/bin/bash
# Round 2, GPU call B: (1) does -fmad=false make the two exact schedulers bit-identical, and what does it cost;
# (2) block shape / pool size sweep of the block wavefront; (3) ncu --set full of the pool kernel on scene 8 and Cornell.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2b; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== diag, fmad off"; MORT_B200_LIB=$ROOT/ab_fmadoff/libmort_b200.so timeout 300 python scripts/diag_pool_vs_mega.py 2>&1 | tail -12 | tee $OUT/diag_fmadoff.txt
run() { tag=$1; bin=$2; shift; shift; echo -n "$tag: "; timeout 120 $bin "$@" 2>&1 | tail -1 | tee -a $OUT/ab.jsonl | cut -c1-150; echo "  # $tag :: $*" >> $OUT/ab.jsonl; }
S8="8 --width 800 --spp 256 --depth 40"; S6="6 --width 600 --spp 256 --depth 50"; S1="1"; SF="1 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50"
for cfg in "$S8" "$S6" "$S1" "$SF"; do
  echo "=== $cfg"
  for b in mort_b200/mort ab_fmadoff/mort; do
    run "mega[$b]" $b $cfg --frames 2
    run "pool_512x2_1024[$b]" $b $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 1024
    run "pool_1024x1_2048[$b]" $b $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 2048
  done
  M=mort_b200/mort
  run pool_1024x1_1024 $M $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 1024
  run pool_1024x1_1536 $M $cfg --frames 2 --mode pool --tpb 1024 --bps 1 --pool 1536
  run pool_768x1_1536 $M $cfg --frames 2 --mode pool --tpb 768 --bps 1 --pool 1536
  run pool_768x1_2048 $M $cfg --frames 2 --mode pool --tpb 768 --bps 1 --pool 2048
  run pool_640x1_2048 $M $cfg --frames 2 --mode pool --tpb 640 --bps 1 --pool 2048
  run pool_512x2_768 $M $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 768
  run pool_512x2_512 $M $cfg --frames 2 --mode pool --tpb 512 --bps 2 --pool 512
done
echo "== ncu scene 8"
CMD8="mort_b200/mort 8 --width 800 --spp 64 --depth 40 --mode pool --tpb 1024 --bps 1 --pool 2048"
$CMD8 > $OUT/plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_kernel -c 1 -o $OUT/prof_pool_s8 $CMD8 > $OUT/ncu8.log 2>&1
tail -3 $OUT/ncu8.log
echo "== ncu cornell"
CMD6="mort_b200/mort 6 --width 600 --spp 64 --depth 50 --mode pool --tpb 512 --bps 2 --pool 1024"
$CMD6 > $OUT/plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_kernel -c 1 -o $OUT/prof_pool_s6 $CMD6 > $OUT/ncu6.log 2>&1
tail -3 $OUT/ncu6.log
for t in s8 s6; do
  if [ -f $OUT/prof_pool_$t.ncu-rep ]; then
    ncu -i $OUT/prof_pool_$t.ncu-rep --page raw --csv > $OUT/prof_pool_${t}_raw.csv 2>/dev/null
    ncu -i $OUT/prof_pool_$t.ncu-rep --page source --csv > $OUT/prof_pool_${t}_source.csv 2>/dev/null
    ls -la $OUT/prof_pool_$t.ncu-rep
  fi
done
du -sh $OUT
