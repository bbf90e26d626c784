#!/bin/bash
# GPU call 4: megakernel v3 (dynamic in-warp sample distribution, exact fixed-point accumulation, linear scan
# for small scenes), register-capped variants, ncu for Cornell (linear) and scene 8 (BVH).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1d; mkdir -p $OUT $ROOT/gpurun_out/ref
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 3 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== variants"
for bps in 4 6 8; do
  timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 1 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
done
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 2 --tpb 64 --bps 8 | tail -1 | tee -a $OUT/variants.jsonl
echo "== cli defaults"
for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
timeout 120 mort_b200/mort 1 --width 400 --spp 32 --depth 50 --frames 5 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 1 --field 500 --width 1920 --spp 64 --depth 50 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
timeout 300 mort_b200/mort 1 --field 500 --fieldcam 1 --width 1920 --spp 64 --depth 50 --frames 2 | tail -1 | tee -a $OUT/cli_configs.jsonl
echo "== ncu full mega cornell"
timeout 300 python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/plain_for_ncu2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_cornell python bench.py --steps 1 --warmup 1 --spp 256 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "== ncu full mega scene 8"
timeout 300 python bench.py --steps 1 --warmup 1 --scene 8 --width 400 --spp 64 --depth 40 --no-cpu-baseline > $OUT/plain_for_ncu3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 1 -c 1 -o $OUT/prof_mega_scene8 python bench.py --steps 1 --warmup 1 --scene 8 --width 400 --spp 64 --depth 40 --no-cpu-baseline > $OUT/ncu_full3.log 2>&1
ls -la $OUT | tail -12
echo "== reference: scene 8 converged-ish (48x48, 256 spp, two seeds)"
cd $ROOT/oracle/_ref; R=$ROOT/gpurun_out/ref
for seed in 69420 12345; do n=a; [ $seed = 12345 ] && n=b
  timeout 100 ./mort_ref --scene 8 --width 48 --spp 256 --seed $seed --hdr $R/convhdr_8_$n.mimg >> $R/log3.jsonl 2>>$R/stderr3.txt
done
echo '{"spp":256}' > $R/conv_meta_8.json
cat $R/log3.jsonl
