#!/bin/bash
# GPU call 5: stage block at the top of the segment, quad pre-division reject, default variants; leaf-size sweep.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1e; mkdir -p $OUT $ROOT/gpurun_out/ref
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
echo "== bench wave"; timeout 600 python bench.py --steps 2 --warmup 1 --mode wave --no-cpu-baseline 2>$OUT/bench_wave.err | tee $OUT/bench_wave.json
echo "== cli defaults"
for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
echo "== configs"
timeout 120 mort_b200/mort 1 --width 400 --spp 32 --depth 50 --frames 5 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 1 --field 500 --width 1920 --spp 256 --depth 50 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 1 --field 500 --fieldcam 1 --width 1920 --spp 256 --depth 50 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
echo "== leaf size sweep"
for ml in 2 4 8; do
  MORT_MAX_LEAF=$ml timeout 120 mort_b200/mort 1 --frames 3 | tail -1 | tee -a $OUT/leaf_sweep.jsonl
  MORT_MAX_LEAF=$ml timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 | tail -1 | tee -a $OUT/leaf_sweep.jsonl
  MORT_MAX_LEAF=$ml timeout 300 mort_b200/mort 1 --field 500 --width 1920 --spp 64 --depth 50 --frames 2 | tail -1 | tee -a $OUT/leaf_sweep.jsonl
done
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain_for_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full mega cornell"
timeout 300 python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/plain_for_ncu2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_cornell python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -8
echo "== reference: scene 8 (48x48, 64 spp, two seeds)"
cd $ROOT/oracle/_ref; R=$ROOT/gpurun_out/ref
for seed in 69420 12345; do n=a; [ $seed = 12345 ] && n=b
  timeout 200 ./mort_ref --scene 8 --width 48 --spp 64 --seed $seed --hdr $R/convhdr_8_$n.mimg >> $R/log3.jsonl 2>>$R/stderr3.txt
done
echo '{"spp":64}' > $R/conv_meta_8.json
cat $R/log3.jsonl; tail -3 $R/stderr3.txt
