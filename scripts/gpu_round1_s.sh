#!/bin/bash
# GPU call 20: validation of the adopted kernel (nearest-only ordering, packed stack, unrolled Philox): whole GPU suite, bench both arms,
# CLI defaults + BASELINE configs, and a last A/B of the Philox unroll factor (new = 10, u2 = 2) against commit b73a777 (old).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1s; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu"; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-200
for rep in 1 2 3; do for v in old new u2; do
  exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  for s in 6 1 8; do
    extra=""; [ $s = 8 ] && extra="--spp 256"
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 $extra 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
for v in old new u2; do exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  echo -n "{\"v\":\"$v\",\"rep\":1,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe 11 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50 --frames 2 2>&1 | tail -1 | sed 's/"scene":1,/"scene":"field",/; s/$/}/' >> $OUT/ab.jsonl
done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1s/ab.jsonl'):
    j=json.loads(l); d[(str(j['r']['scene']),j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
echo "== CLI defaults"; for s in 1 2 3 4 5 6 7 8 9 10; do timeout 300 mort_b200/mort $s --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_defaults.jsonl | cut -c1-140; done
echo "== BASELINE configs"
timeout 300 mort_b200/mort 1 --width 400 --aspect 1.7777778 --spp 32 --depth 50 --frames 20 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
timeout 300 mort_b200/mort 8 --width 800 --spp 4096 --depth 40 --frames 1 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140
for c in 0 1; do timeout 300 mort_b200/mort 1 --field 500 --fieldcam $c --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | tee -a $OUT/cli_configs.jsonl | cut -c1-140; done
