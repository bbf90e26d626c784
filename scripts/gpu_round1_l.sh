#!/bin/bash
# GPU call 13: native 32-bit shared atomics + guided task tail + leaner node step: parity suite, A/B on small frames, reference-arm repeatability.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1l; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
for s in 1 8 9; do timeout 300 mort_b200/mort $s --frames 3 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl; done
echo "== tail A/B (per-GPU share of the bench frame at N=8 is 128 spp)"
for spp in 128 64 16; do for t in 0 1; do
  echo -n "{\"spp\":$spp,\"MORT_TAIL\":$t,\"line\":" >> $OUT/tail_ab.jsonl
  MORT_TAIL=$t timeout 300 python bench.py --steps 8 --warmup 3 --spp $spp --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'value':d['value'],'ms':d['ms_per_step']})+'}')" >> $OUT/tail_ab.jsonl
done; done
cat $OUT/tail_ab.jsonl
echo "== configs 1 and 4"
timeout 300 mort_b200/mort 1 --width 400 --aspect 1.7777778 --spp 32 --depth 50 --frames 20 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl
for t in 0 1; do MORT_TAIL=$t timeout 300 mort_b200/mort 1 --field 500 --fieldcam 0 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl; done
timeout 300 mort_b200/mort 1 --field 500 --fieldcam 1 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl
echo "== reference arm x3"
for i in 1 2 3; do timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$OUT/ref_$i.err | tee -a $OUT/bench_reference.jsonl; tail -2 $OUT/ref_$i.err; done
