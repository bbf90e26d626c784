#!/bin/bash
# Round 2, 8-GPU call: BASELINE config 5 (scenes 1-10 at 3840x2160, 256 spp per frame) and config 4 (1 M-sphere field, 1920x1080,
# 256 spp) at N = 1, 2, 4, 8 through the C-ABI group path (`mort --gpus N`: one process, one host thread + one NCCL communicator
# per GPU, ONE ncclReduce of the exact partial frames per frame), plus bench.py under torchrun at N = 2, 4, 8.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2m8; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
M=mort_b200/mort
echo "== config 5: sample split"
for s in 8 1 2 3 4 5 6 7 9 10; do for n in 1 2 4 8; do
  timeout 300 $M $s --width 3840 --aspect 1.7777778 --spp 256 --frames 2 --gpus $n --split sample 2>&1 | tail -1 | tee -a $OUT/cfg5_sample.jsonl | cut -c1-250
done; done
echo "== config 5: tile split (scenes 8, 1, 6)"
for s in 8 1 6; do for n in 2 4 8; do
  timeout 300 $M $s --width 3840 --aspect 1.7777778 --spp 256 --frames 2 --gpus $n --split tile 2>&1 | tail -1 | tee -a $OUT/cfg5_tile.jsonl | cut -c1-250
done; done
echo "== config 4: 1 M spheres"
for cam in 0 1; do for n in 1 2 4 8; do
  timeout 300 $M 1 --field 500 --fieldcam $cam --width 1920 --aspect 1.7777778 --spp 256 --depth 50 --frames 2 --gpus $n 2>&1 | tail -1 | tee -a $OUT/cfg4.jsonl | cut -c1-250
done; done
echo "== bench.py scaling"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-per-config 2>$OUT/bench1.err | tee $OUT/bench_n1.json | cut -c1-200
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 2>$OUT/bench$n.err | tee $OUT/bench_n$n.json | cut -c1-200
done
echo "== 4096-spp frame of config 3 on 8 GPUs (the frame the config names)"
timeout 300 $M 8 --width 800 --spp 4096 --depth 40 --frames 2 --gpus 8 2>&1 | tail -1 | tee -a $OUT/cfg3_8gpu.jsonl | cut -c1-250
