#!/bin/bash
# GPU call 18: same-box A/B of kernel variants: packed stack entries (new), inline stage-block Philox (y1), nearest-only child ordering (y2),
# unrolled Philox block (y3), y2+y3 (y4), against commit b73a777 (old).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1q; mkdir -p $OUT
for rep in 1 2 3; do for v in old new y1 y2 y3 y4; do
  exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  for s in 6 1 8; do
    extra=""; [ $s = 8 ] && extra="--spp 256"
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 $extra 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
for v in old new y2 y3 y4; do exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  echo -n "{\"v\":\"$v\",\"rep\":1,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe 11 --field 500 --width 1920 --aspect 1.7777778 --spp 64 --depth 50 --frames 2 2>&1 | tail -1 | sed 's/"scene":1,/"scene":"field",/; s/$/}/' >> $OUT/ab.jsonl
done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1q/ab.jsonl'):
    j=json.loads(l); d[(str(j['r']['scene']),j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_trace.py tests/test_gpu_field.py -q -x --timeout 600 2>&1 | tail -3
