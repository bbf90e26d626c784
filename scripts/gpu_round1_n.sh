#!/bin/bash
# GPU call 15: same-box A/B of the megakernel before (ab_old/, commit b73a777) and after the atomics / tail / node-step changes;
# L1 carve-out sweep; rerun of the scene-text oracle test.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1n; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== scene text test"; timeout 600 python -m pytest tests/test_gpu_scene_text.py -q --timeout 300 2>&1 | tail -4
echo "== A/B old vs new"
for rep in 1 2 3; do for v in old new; do
  exe=mort_b200/mort; [ $v = old ] && exe=ab_old/mort
  for s in 6 1 8; do
    extra=""; [ $s = 8 ] && extra="--spp 256"
    echo -n "{\"v\":\"$v\",\"rep\":$rep,\"r\":" >> $OUT/ab.jsonl; timeout 300 $exe $s --frames 3 $extra 2>&1 | tail -1 | sed 's/$/}/' >> $OUT/ab.jsonl
  done
done; done
python - <<'PY'
import json,collections
d=collections.defaultdict(list)
for l in open('gpurun_out/r1n/ab.jsonl'):
    j=json.loads(l); d[(j['r']['scene'],j['v'])].append(j['r']['msamples_per_s'])
for k in sorted(d): print(k, ['%.0f'%x for x in d[k]])
PY
echo "== carveout sweep"
for c in -1 0 25 50 100; do for s in 6 1 8; do
  extra=""; [ $s = 8 ] && extra="--spp 256"
  if [ $c = -1 ]; then r=$(timeout 300 mort_b200/mort $s --frames 3 $extra 2>&1 | tail -1); else r=$(MORT_CARVEOUT=$c timeout 300 mort_b200/mort $s --frames 3 $extra 2>&1 | tail -1); fi
  echo "{\"carveout\":$c,\"r\":$r}" | tee -a $OUT/carveout.jsonl | cut -c1-160
done; done
