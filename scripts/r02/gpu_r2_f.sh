#!/bin/bash
# Round 2, GPU call F: ncu --set full + source of the phased block wavefront with lane refill on scene 8 (the library of this
# very snapshot is kept as gpurun_out/r2f/libmort_b200_r2f.so for the inline-chain attribution).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2f; mkdir -p $OUT
CMD8="mort_b200/mort 8 --width 800 --spp 64 --depth 40 --mode pool --tpb 640 --bps 1 --pool 2048 --refill 16"
$CMD8 > $OUT/plain8.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_kernel -c 1 -o $OUT/prof_pool_s8 $CMD8 > $OUT/ncu8.log 2>&1
tail -2 $OUT/ncu8.log
ncu -i $OUT/prof_pool_s8.ncu-rep --page raw --csv > $OUT/prof_pool_s8_raw.csv 2>/dev/null
ncu -i $OUT/prof_pool_s8.ncu-rep --page source --csv > $OUT/prof_pool_s8_source.csv 2>/dev/null
rm -f $OUT/prof_pool_s8.ncu-rep
ls -la $OUT
