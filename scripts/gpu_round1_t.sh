#!/bin/bash
# GPU call 21 (8 GPUs, one box): the N = 1, 2, 4, 8 strong-scaling run of the bench workload (sample split, exact partial frames,
# one NCCL SUM reduce), tile split at N = 8, and BASELINE config 5's shape (scene 8 at 3840x2160, tile split) at N = 1 and 8 with a
# reduced sample count.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1t; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== N=1"; timeout 200 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline 2>$OUT/n1.err | tee $OUT/scale_n1.json | cut -c1-220
p=29530
for n in 2 4 8; do p=$((p+1)); echo "== N=$n"; timeout 300 $TR --nproc-per-node $n --master-port $p bench.py --gpus $n --steps 4 --warmup 3 --no-cpu-baseline 2>$OUT/n$n.err | tee $OUT/scale_n$n.json | cut -c1-220; tail -2 $OUT/n$n.err | cut -c1-200; done
echo "== N=8 tile split"; timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 4 --warmup 3 --split tile --no-cpu-baseline 2>$OUT/n8t.err | tee $OUT/scale_n8_tile.json | cut -c1-220
echo "== config 5 shape: scene 8 at 3840x2160, 64 spp, tile split"
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 1 --scene 8 --width 3840 --aspect 1.7777778 --spp 64 --depth 40 --no-cpu-baseline 2>$OUT/c5n1.err | tee $OUT/cfg5_n1.json | cut -c1-220
timeout 300 $TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 2 --warmup 1 --scene 8 --width 3840 --aspect 1.7777778 --spp 64 --depth 40 --split tile --no-cpu-baseline 2>$OUT/c5n8.err | tee $OUT/cfg5_n8.json | cut -c1-220
tail -3 $OUT/c5n8.err | cut -c1-200
