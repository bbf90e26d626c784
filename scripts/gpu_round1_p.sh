#!/bin/bash
# GPU call 17: linear/tree kernel specialisation (new) vs unspecialised (x1) vs inline stage-block Philox (x2) vs commit b73a777 (old).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1p; mkdir -p $OUT
for rep in 1 2 3; do for v in old new x1 x2; do
  exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  for s in 6 1 8; do
    extra=""; [ $s = 8 ] && extra="--spp 256"
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 $extra 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
for v in new x2; do exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
 for b in 4 8; do for s in 6 1; do
  echo -n "{\"v\":\"$v-bps$b\",\"rep\":1,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 --bps $b 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
 done; done; done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1p/ab.jsonl'):
    j=json.loads(l); d[(j['r']['scene'],j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_trace.py -q -x --timeout 600 2>&1 | tail -3
