#!/bin/bash
# Runs ON THE GPU BOX (via gpurun), round 2: drives the prebuilt reference harness oracle/_ref/mort_ref for
#   (1) the four harness-built scenes 101-104 (isotropic media + directly sampled quad light, sphere light, lens
#       sampling, medium + visible top-level list + nested wrappers): scene dump, primary-hit records, noisy and
#       converged frames, two seeds each;
#   (2) converged frames of the brute-force scenes at sizes the reference finishes: scene 8 at 96 px / 1024 spp /
#       depth 40, scene 9 at 96 px / 4096 spp, two seeds each;
#   (3) the reference's timing on BASELINE config 3 at 1 and 4 spp per frame (with a warm-up frame);
#   (4) a characterisation of its "invalid program counter" crashes: device malloc heap size sweep.
# Everything lands in gpurun_out/ref2/; tests/golden/ is curated from it by scripts/curate_golden.py gpurun_out/ref2.
set -u
ROOT=$(pwd)
OUT=$ROOT/gpurun_out/ref2
mkdir -p $OUT
cd $ROOT/oracle/_ref
LOG=$OUT/log.jsonl
: > $LOG
run() { echo "+ $*" >> $OUT/cmds.txt; timeout 900 ./mort_ref "$@" >> $LOG 2>> $OUT/stderr.txt || echo "{\"failed\":\"$*\",\"rc\":$?}" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1

for s in 101 102 103 104; do
  run --scene $s --dump-scene $OUT/scene_$s.mscn --trace-grid 64 $OUT/grid_$s.mhit --trace-random 4096 7 $OUT/rnd_$s.mhit
  run --scene $s --width 96 --spp 64 --seed 69420 --img8 $OUT/small8_${s}_a.mimg --hdr $OUT/smallhdr_${s}_a.mimg
  run --scene $s --width 96 --spp 64 --seed 12345 --hdr $OUT/smallhdr_${s}_b.mimg
  run --scene $s --width 160 --spp 1024 --seed 69420 --img8 $OUT/conv8_${s}_a.mimg --hdr $OUT/convhdr_${s}_a.mimg
  run --scene $s --width 160 --spp 1024 --seed 12345 --hdr $OUT/convhdr_${s}_b.mimg
  echo '{"spp": 1024}' > $OUT/conv_meta_$s.json
done

run --scene 8 --width 96 --spp 1024 --depth 40 --seed 69420 --img8 $OUT/conv8_8_a.mimg --hdr $OUT/convhdr_8_a.mimg
run --scene 8 --width 96 --spp 1024 --depth 40 --seed 12345 --hdr $OUT/convhdr_8_b.mimg
echo '{"spp": 1024}' > $OUT/conv_meta_8.json
run --scene 9 --width 96 --spp 4096 --seed 69420 --img8 $OUT/conv8_9_a.mimg --hdr $OUT/convhdr_9_a.mimg
run --scene 9 --width 96 --spp 4096 --seed 12345 --hdr $OUT/convhdr_9_b.mimg
echo '{"spp": 4096}' > $OUT/conv_meta_9.json

# config 3 timing with a warm-up frame (round 1 had one un-warmed 4-spp frame)
run --scene 8 --width 800 --spp 1 --depth 40 --frames 3 --warmup 1
run --scene 8 --width 800 --spp 4 --depth 40 --frames 2 --warmup 1

# crash characterisation: Cornell 600x600 16 spp (the round-1 reference arm), one un-warmed frame per process
for heap in 8 64 1024 8192; do
  for i in 1 2 3 4; do
    echo "{\"crash_probe\":\"heap\",\"heap_mb\":$heap,\"try\":$i}" >> $LOG
    run --scene 6 --width 600 --spp 16 --depth 50 --frames 1 --warmup 0 --heap-mb $heap
  done
done
for i in 1 2 3 4; do
  echo "{\"crash_probe\":\"stack\",\"stack\":32768,\"try\":$i}" >> $LOG
  run --scene 6 --width 600 --spp 16 --depth 50 --frames 1 --warmup 0 --stack 32768
done
ls -la $OUT > $OUT/ls.txt
tail -60 $LOG
