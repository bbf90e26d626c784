#!/bin/bash
# GPU call 2: product parity tests, smoke, bench (both arms), ncu launch list + full capture, CLI timings,
# and the reference fixtures the first call ran out of time for (scene 8/9 converged, reduced spp).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1b; mkdir -p $OUT $ROOT/gpurun_out/ref
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench mort"; timeout 600 python bench.py --steps 3 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json; tail -3 $OUT/bench_ref.err
echo "== cli all scenes (defaults)"
for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 2 --out $OUT/scene_$s.ppm 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
echo "== cli baseline configs"
timeout 120 mort_b200/mort 1 --width 400 --spp 32 --depth 50 --frames 5 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 --stage 100000 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --stage 100000 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/plain_for_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "== ncu full (mega_kernel, cornell 600x600 256spp)"
timeout 300 python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/plain_for_ncu2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_cornell python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -30
echo "== reference fixtures: scene 8/9 converged at reduced size"
cd $ROOT/oracle/_ref
R=$ROOT/gpurun_out/ref
for seed in 69420 12345; do n=a; [ $seed = 12345 ] && n=b
  timeout 400 ./mort_ref --scene 8 --width 64 --spp 1024 --seed $seed --hdr $R/convhdr_8_$n.mimg >> $R/log2.jsonl 2>>$R/stderr2.txt
  timeout 200 ./mort_ref --scene 9 --width 96 --spp 1024 --seed $seed --hdr $R/convhdr_9_$n.mimg >> $R/log2.jsonl 2>>$R/stderr2.txt
done
echo '{"spp":1024}' > $R/conv_meta_8.json; echo '{"spp":1024}' > $R/conv_meta_9.json
timeout 200 ./mort_ref --scene 1 --width 400 --spp 32 --depth 50 --frames 5 >> $R/log2.jsonl 2>>$R/stderr2.txt
timeout 200 ./mort_ref --scene 8 --width 800 --spp 4 --depth 40 --frames 1 --warmup 0 >> $R/log2.jsonl 2>>$R/stderr2.txt
cat $R/log2.jsonl
