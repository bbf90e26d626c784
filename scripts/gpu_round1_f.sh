#!/bin/bash
# GPU call 6 (2 GPUs): the N > 1 path — NCCL sample split through bench.py under torchrun, and the 2-GPU parity test.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1f; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
echo "== 2-GPU parity test"; timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 300 2>&1 | tail -8 | tee $OUT/pytest_dist.txt
echo "== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --no-cpu-baseline 2>$OUT/b1.err | tee $OUT/bench_n1.json; tail -2 $OUT/b1.err
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 2>$OUT/b2.err | tee $OUT/bench_n2.json; tail -5 $OUT/b2.err
echo "== bench N=2 reference arm (rank 0 only)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>$OUT/b2r.err | tee $OUT/bench_n2_ref.json; tail -3 $OUT/b2r.err
