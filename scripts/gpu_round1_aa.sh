#!/bin/bash
# GPU call 29 (2 GPUs): stdout of bench.py under torchrun must be exactly one JSON line (NCCL banner goes to stderr); 2-rank NCCL tests.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1aa; mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --steps 4 --warmup 3 > $OUT/n2.stdout 2> $OUT/n2.stderr
echo "stdout lines: $(wc -l < $OUT/n2.stdout)"; cut -c1-200 $OUT/n2.stdout; grep -c "NCCL version" $OUT/n2.stderr
timeout 300 $TR --nproc-per-node 2 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/n2ref.stdout 2> $OUT/n2ref.stderr
echo "reference stdout lines: $(wc -l < $OUT/n2ref.stdout)"; cut -c1-160 $OUT/n2ref.stdout
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_bench.py -q --timeout 600 2>&1 | tail -3
