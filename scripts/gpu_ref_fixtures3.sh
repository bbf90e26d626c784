#!/bin/bash
# Runs ON THE GPU BOX (via gpurun), round 2, second part: what scripts/gpu_ref_fixtures2.sh could not finish in its time limit.
# The reference is thread-per-pixel: at 96 px only 36 of its 16x16 blocks exist and scene 8 runs at 0.02 Msamples/s, so the converged
# frames of the brute-force scenes are taken at 256 px (256 blocks, ~0.15 Msamples/s) with 256 spp — the same number of samples as
# 128 px / 1024 spp; tests compare per pixel AND on 8x8-binned frames (= 16 384 samples per bin).
set -u
ROOT=$(pwd)
OUT=$ROOT/gpurun_out/ref3
mkdir -p $OUT
cd $ROOT/oracle/_ref
LOG=$OUT/log.jsonl
: > $LOG
run() { echo "+ $*" >> $OUT/cmds.txt; timeout 900 ./mort_ref "$@" >> $LOG 2>> $OUT/stderr.txt || echo "{\"failed\":\"$*\",\"rc\":$?}" >> $LOG; }
run --scene 8 --width 256 --spp 256 --depth 40 --seed 69420 --hdr $OUT/convhdr_8_a.mimg
run --scene 8 --width 256 --spp 256 --depth 40 --seed 12345 --hdr $OUT/convhdr_8_b.mimg
echo '{"spp": 256}' > $OUT/conv_meta_8.json
run --scene 9 --width 256 --spp 1024 --seed 69420 --hdr $OUT/convhdr_9_a.mimg
run --scene 9 --width 256 --spp 1024 --seed 12345 --hdr $OUT/convhdr_9_b.mimg
echo '{"spp": 1024}' > $OUT/conv_meta_9.json
# config 3 timing with a warm-up frame (round 1 had one un-warmed 4-spp frame)
run --scene 8 --width 800 --spp 1 --depth 40 --frames 3 --warmup 1
run --scene 8 --width 800 --spp 4 --depth 40 --frames 2 --warmup 1
# crash characterisation: Cornell 600x600 16 spp (the round-1 reference arm), one un-warmed frame per process
for heap in 8 64 1024 8192; do
  for i in 1 2 3; do
    echo "{\"crash_probe\":\"heap\",\"heap_mb\":$heap,\"try\":$i}" >> $LOG
    run --scene 6 --width 600 --spp 16 --depth 50 --frames 1 --warmup 0 --heap-mb $heap
  done
done
for i in 1 2 3; do
  echo "{\"crash_probe\":\"stack\",\"stack\":32768,\"try\":$i}" >> $LOG
  run --scene 6 --width 600 --spp 16 --depth 50 --frames 1 --warmup 0 --stack 32768
done
cat $LOG | cut -c1-250
tail -5 $OUT/stderr.txt
