#!/bin/bash
# GPU call 16: bisect the 3 % Cornell regression of the new megakernel build (variants built with -DMORT_EXP_*), same box.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1o; mkdir -p $OUT
for rep in 1 2 3; do for v in old new v1 v2 v3 v4; do
  exe=ab_$v/mort; [ $v = new ] && exe=mort_b200/mort
  for s in 6 1; do
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1o/ab.jsonl'):
    j=json.loads(l); d[(j['r']['scene'],j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
