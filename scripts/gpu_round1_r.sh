#!/bin/bash
# GPU call 19: Philox unroll factor (z2, z5, z10), approximate-quotient plane pre-test (new vs zA = z10 without it), nearest-only ordering adopted.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1r; mkdir -p $OUT
for rep in 1 2 3; do for v in old new z2 z5 z10 zA; do
  exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  for s in 6 1 8; do
    extra=""; [ $s = 8 ] && extra="--spp 256"
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 $extra 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
for v in old new z5 z10 zA; do exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  echo -n "{\"v\":\"$v\",\"rep\":1,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe 11 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50 --frames 2 2>&1 | tail -1 | sed 's/"scene":1,/"scene":"field",/; s/$/}/' >> $OUT/ab.jsonl
done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1r/ab.jsonl'):
    j=json.loads(l); d[(str(j['r']['scene']),j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_trace.py tests/test_gpu_field.py -q -x --timeout 600 2>&1 | tail -3
