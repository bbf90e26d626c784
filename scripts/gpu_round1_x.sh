#!/bin/bash
# GPU call 26: does raising the stack limit at mort_create (or at the first render) recover the kernel time inside a torch process?
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1x; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for how in off create render; do
  MORT_STACK_FIX=$how timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'stack_fix':'$how','value':d['value'],'ms_per_step':d['ms_per_step'],'kernel_ms':d['roofline']['kernel_ms_per_launch'],'e2e':d['e2e']['value']}))" | tee -a $OUT/stack_fix.jsonl
  MORT_STACK_FIX=$how timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 4 2>&1 | tail -1 | sed "s/^{/{\"stack_fix\":\"$how\",/" | tee -a $OUT/stack_fix.jsonl
  for s in 1 8; do MORT_STACK_FIX=$how timeout 120 mort_b200/mort $s --frames 3 2>&1 | tail -1 | sed "s/^{/{\"stack_fix\":\"$how\",/" | tee -a $OUT/stack_fix.jsonl | cut -c1-150; done
  MORT_STACK_FIX=$how timeout 300 python bench.py --steps 3 --warmup 2 --scene 8 --width 800 --spp 256 --depth 40 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'stack_fix':'$how','scene':8,'value':d['value'],'ms_per_step':d['ms_per_step']}))" | tee -a $OUT/stack_fix.jsonl
done
