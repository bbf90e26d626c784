#!/bin/bash
# Round 2, GPU call N (the last 2.9 GPU-minutes): the CLI's new default MORT_MODE_AUTO against --mode pool on the scenes where the rule
# picks the megakernel (2, 3, 5, 7) and two where it keeps the block wavefront (6, 8), same box; byte-equality of the two frames.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2n; mkdir -p $OUT
M=mort_b200/mort
for s in 3 2 5 7 6 8; do
  extra=""; [ $s = 7 ] && extra="--spp 256"; [ $s = 8 ] && extra="--spp 64"; [ $s = 6 ] && extra="--spp 256"
  for m in auto pool mega; do
    timeout 60 $M $s --frames 3 $extra --mode $m --out $OUT/s${s}_$m.ppm 2>&1 | tail -1 | cut -c1-170 | tee -a $OUT/auto_ab.jsonl; echo "  # $m :: scene $s $extra" | tee -a $OUT/auto_ab.jsonl
  done
  cmp $OUT/s${s}_auto.ppm $OUT/s${s}_pool.ppm && echo "scene $s: auto == pool frame" | tee -a $OUT/auto_ab.jsonl
  cmp $OUT/s${s}_mega.ppm $OUT/s${s}_pool.ppm && echo "scene $s: mega == pool frame" | tee -a $OUT/auto_ab.jsonl
  rm -f $OUT/*.ppm
done
