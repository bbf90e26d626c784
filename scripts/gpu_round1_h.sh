#!/bin/bash
# GPU call 8: octant node test, smem accumulators (pre-filter reverted); variants; ncu of the wavefront kernels.
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/r1h; mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench mort mega"; timeout 600 python bench.py --steps 5 --warmup 3 2>$OUT/bench_mort.err | tee $OUT/bench_mort.json; tail -3 $OUT/bench_mort.err
echo "== variants"
for bps in 4 6 8; do
  timeout 120 mort_b200/mort 6 --width 600 --spp 1024 --depth 50 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 1 --frames 3 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 8 --width 800 --spp 256 --depth 40 --frames 2 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
  timeout 120 mort_b200/mort 1 --field 500 --fieldcam 1 --width 1920 --spp 64 --depth 50 --frames 2 --bps $bps | tail -1 | tee -a $OUT/variants.jsonl
done
echo "== wavefront: slots sweep + ncu"
for paths in 2000000 8000000; do timeout 120 python - <<PY | tee -a $OUT/wave_sweep.jsonl
import sys, json; sys.path.insert(0, '.')
from mort_b200.api import Renderer
r = Renderer(0); r.build_scene(6).override_camera(width=600, spp=1024, depth=50).commit()
for i in range(2):
    fr = r.render(want_rgba8=False, want_accum=False, mode=1, wavefront_paths=$paths, frame=i)
st = fr.stats; print(json.dumps({"paths": $paths, "ms": st["last_render_ms"], "msamples_per_s": st["last_samples"] / st["last_render_ms"] / 1e3, "launches": st["last_kernel_launches"]}))
PY
done
timeout 300 python bench.py --steps 1 --warmup 1 --mode wave --spp 64 --no-cpu-baseline > $OUT/plain_for_ncu_wave.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:wf_ -s 40 -c 24 -o $OUT/prof_wave_cornell python bench.py --steps 1 --warmup 1 --mode wave --spp 64 --no-cpu-baseline > $OUT/ncu_wave.log 2>&1
ls -la $OUT | tail -6
