#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): drives the prebuilt reference harness oracle/_ref/mort_ref to
# produce (1) golden fixtures (scene dumps, primary-hit records, small and converged images) and
# (2) the reference's own timings on the BASELINE.json configurations.
# Everything lands in gpurun_out/ref/; tests/golden/ is curated from it by scripts/curate_golden.py.
set -u
ROOT=$(pwd)
OUT=$ROOT/gpurun_out/ref
mkdir -p $OUT
cd $ROOT/oracle/_ref
LOG=$OUT/log.jsonl
: > $LOG
run() { echo "+ $*" >> $OUT/cmds.txt; timeout 600 ./mort_ref "$@" >> $LOG 2>> $OUT/stderr.txt || echo "{\"failed\":\"$*\",\"rc\":$?}" >> $LOG; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1

for s in 1 2 3 4 5 6 7 8 9 10; do
  extra=""
  [ $s = 3 ] && extra="--dump-image-rgb $OUT/earthmap.ppm"
  # scene dump at scene defaults + small primary-hit fixtures
  run --scene $s --dump-scene $OUT/scene_$s.mscn $extra --trace-grid 64 $OUT/grid_$s.mhit --trace-random 2048 7 $OUT/rnd_$s.mhit
  # larger hit sets for three scenes (kept out of git; used once to validate the oracle here; gpurun_out is capped at 64 MiB)
  case $s in 1|6|8) run --scene $s --trace-grid 200 $OUT/gridbig_$s.mhit --trace-random 30000 11 $OUT/rndbig_$s.mhit;; esac
  # small noisy images for the CPU oracle's statistical check: 96 wide, 64 spp, two seeds
  run --scene $s --width 96 --spp 64 --seed 69420 --img8 $OUT/small8_${s}_a.mimg --hdr $OUT/smallhdr_${s}_a.mimg
  run --scene $s --width 96 --spp 64 --seed 12345 --hdr $OUT/smallhdr_${s}_b.mimg
done

# converged references, 4096 spp (64x64 strata), two seeds -> noise floor
for s in 1 2 3 4 5 6 7 10; do
  run --scene $s --width 240 --spp 4096 --seed 69420 --img8 $OUT/conv8_${s}_a.mimg --hdr $OUT/convhdr_${s}_a.mimg
  run --scene $s --width 240 --spp 4096 --seed 12345 --hdr $OUT/convhdr_${s}_b.mimg
done
for s in 8 9; do
  run --scene $s --width 160 --spp 4096 --seed 69420 --img8 $OUT/conv8_${s}_a.mimg --hdr $OUT/convhdr_${s}_a.mimg
  run --scene $s --width 160 --spp 4096 --seed 12345 --hdr $OUT/convhdr_${s}_b.mimg
done

# timings: BASELINE.json configs (reduced spp where the reference would take too long) + scene defaults
run --scene 1 --width 400 --spp 32 --depth 50 --frames 8
run --scene 6 --width 600 --spp 64 --depth 50 --frames 4
run --scene 6 --width 600 --spp 1024 --depth 50 --frames 2
run --scene 8 --width 800 --spp 16 --depth 40 --frames 3
run --scene 8 --width 800 --spp 64 --depth 40 --frames 2
for s in 1 2 3 4 5 7 9 10; do
  run --scene $s --spp 64 --frames 3
done
run --scene 1 --frames 3
run --scene 1 --width 3840 --spp 16 --frames 2
run --scene 6 --width 3840 --aspect 1.7777778 --spp 16 --frames 2
ls -la $OUT > $OUT/ls.txt
cat $LOG
