#!/bin/bash
# Round 2, 2-GPU call: NCCL paths behind the C ABI (mort_group_*, mort_comm_*), bench.py under torchrun, the CLI's --gpus.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r2m2; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_dist.py tests/test_gpu_pool.py -m gpu -q --timeout 300 2>&1 | tail -12 | tee $OUT/pytest.txt
echo "== CLI"; for split in sample tile; do for n in 1 2; do timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --gpus $n --split $split 2>&1 | tail -1 | tee -a $OUT/cli.jsonl | cut -c1-330; done; done
timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --gpus 2 --out $OUT/s8_2gpu.ppm | tail -1 | cut -c1-100; timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --out $OUT/s8_1gpu.ppm | tail -1 | cut -c1-100; cmp $OUT/s8_2gpu.ppm $OUT/s8_1gpu.ppm && echo "2-GPU PPM == 1-GPU PPM"; rm -f $OUT/*.ppm
echo "== bench N=1 short"; timeout 600 python bench.py --steps 3 --warmup 3 2>$OUT/bench1.err | tee $OUT/bench_n1.json | cut -c1-400
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>$OUT/bench2.err | tee $OUT/bench_n2.json | cut -c1-700; tail -3 $OUT/bench2.err
echo "== bench N=2 tile"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --split tile 2>$OUT/bench2t.err | tee $OUT/bench_n2_tile.json | cut -c1-300; tail -3 $OUT/bench2t.err
