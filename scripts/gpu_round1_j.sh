#!/bin/bash
# GPU call 11 (2 GPUs): final validation of the N > 1 bench path with exact partial frames + full GPU test suite + bench.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1j; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu (all, 2 GPUs visible)" ; timeout 1100 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 2>$OUT/b1.err | tee $OUT/bench_n1.json; tail -2 $OUT/b1.err
echo "== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2>$OUT/b2.err | tee $OUT/bench_n2.json; tail -3 $OUT/b2.err
echo "== scene 8 / field at N=2 (sample split through the CLI-equivalent bench flags)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2 --warmup 1 --scene 8 --width 800 --spp 1024 --depth 40 2>$OUT/b2s8.err | tee $OUT/bench_n2_scene8.json
timeout 300 python bench.py --gpus 1 --steps 2 --warmup 1 --scene 8 --width 800 --spp 1024 --depth 40 --no-cpu-baseline 2>$OUT/b1s8.err | tee $OUT/bench_n1_scene8.json
