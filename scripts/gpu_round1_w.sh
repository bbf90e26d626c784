#!/bin/bash
# GPU call 24: isolate the ~3 % gap between the kernel time inside bench.py and inside the CLI.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1w; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 4 2>&1 | tail -1 | tee $OUT/probe.jsonl
timeout 200 python scripts/probe_process_effect.py plain 2>&1 | tail -4 | tee -a $OUT/probe.jsonl
timeout 300 python scripts/probe_process_effect.py torch 2>&1 | tail -12 | tee -a $OUT/probe.jsonl
timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 4 2>&1 | tail -1 | tee -a $OUT/probe.jsonl
