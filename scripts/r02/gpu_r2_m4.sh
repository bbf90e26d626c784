#!/bin/bash
# Round 2, 4-GPU call (the GPU-minute budget left no room for the 8-GPU matrix of gpu_r2_m8.sh): BASELINE config 5's shape (3840x2160,
# 256 spp per frame) on scenes 8 and 1 and config 4 (1 M spheres, 1920x1080, 256 spp) at N = 1, 2, 4 through the C-ABI group path
# (`mort --gpus N`: one process, one host thread + one NCCL communicator per GPU, ONE ncclReduce of the exact partial frames per
# frame), plus bench.py under torchrun at N = 4.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2m4; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
M=mort_b200/mort
for s in 8 1; do
  timeout 120 $M $s --width 3840 --aspect 1.7777778 --spp 256 --frames 2 2>&1 | tail -1 | tee -a $OUT/cfg5.jsonl | cut -c1-200
  for n in 2 4; do for sp in sample tile; do
    timeout 120 $M $s --width 3840 --aspect 1.7777778 --spp 256 --frames 2 --gpus $n --split $sp 2>&1 | tail -1 | tee -a $OUT/cfg5.jsonl | cut -c1-260
  done; done
done
echo "== config 4"
timeout 120 $M 1 --field 500 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 2>&1 | tail -1 | tee -a $OUT/cfg4.jsonl | cut -c1-200
for n in 2 4; do
  timeout 120 $M 1 --field 500 --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 --gpus $n 2>&1 | tail -1 | tee -a $OUT/cfg4.jsonl | cut -c1-260
done
echo "== bench.py N=4"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 2>$OUT/bench4.err | tee $OUT/bench_n4.json | cut -c1-300
tail -2 $OUT/bench4.err
