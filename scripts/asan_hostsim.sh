#!/bin/bash
# AddressSanitizer + UBSan over the host build of the product's flattening / tree builder / per-ray code (tests/hostsim): the ten shipped
# scenes, a sphere field and N random scene files (trace, render, text dump).  compute-sanitizer is closed on the GPU pool, so this is the
# memory-safety evidence for the code both builds share.  Usage: scripts/asan_hostsim.sh [N=120]
set -eu
cd "$(dirname "$0")/.."
N=${1:-120}
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -DMORT_GENERAL_MEDIA -Imort_b200/csrc -Iinclude \
    tests/hostsim/hostsim.cpp mort_b200/csrc/scene.cpp mort_b200/csrc/scenes.cpp mort_b200/csrc/scene_text.cpp mort_b200/csrc/flatten.cpp \
    mort_b200/csrc/bvh_build.cpp -o /tmp/hostsim_asan
# the GPU tree builder's and the refit pass's per-thread bodies + level loop, run serially (tests/hostsim/buildsim.cpp)
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -Imort_b200/csrc -Iinclude \
    tests/hostsim/buildsim.cpp mort_b200/csrc/scene.cpp mort_b200/csrc/scenes.cpp mort_b200/csrc/scene_text.cpp mort_b200/csrc/flatten.cpp \
    mort_b200/csrc/bvh_build.cpp -o /tmp/buildsim_asan -lpthread
bad=0
for what in "1 64" "8 4" "8 64" "field:40 16" "dup:300 4" "rand:2000:3 1" "rand:9000:7 64"; do
  /tmp/buildsim_asan ${what% *} mort_b200/assets ${what#* } 2>&1 | grep -E "ERROR: AddressSanitizer|runtime error|DIFFERENT|violations" | grep -v " 0 violations; " | grep -v " 0 edit-violations" && bad=1
done
for sc in 1 2 3 4 5 6 7 8 9 10 field:60; do
  /tmp/hostsim_asan $sc mort_b200/assets render 48 4 0 3 /tmp/asan.mimg 2>&1 | grep -E "ERROR: AddressSanitizer|runtime error" && bad=1
  /tmp/hostsim_asan $sc mort_b200/assets checkbvh 2>&1 | grep -E "ERROR: AddressSanitizer|runtime error" && bad=1
done
python - "$N" <<'PY' || bad=1
import subprocess, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from mort_b200 import formats as F
import test_scene_text as T
hs, A, bad, ran = "/tmp/hostsim_asan", "mort_b200/assets", 0, 0
for seed in range(30000, 30000 + int(sys.argv[1])):
    rng = np.random.default_rng(seed)
    open("/tmp/asan.txt", "w").write(T._random_scene_text(rng, 0.7 if seed % 3 == 0 else 0.1))     # every third scene nests media in wrappers / lists
    n = 500
    o, tgt = rng.uniform(-8, 8, (n, 3)), rng.uniform(-5, 5, (n, 3))
    F.write_hits("/tmp/asan_in.mhit", np.concatenate([o, tgt - o, rng.random((n, 1))], 1).astype(np.float32), np.zeros(n, dtype=F.hit_dt))
    for args in (["trace", "/tmp/asan_in.mhit", "/tmp/asan_out.mhit"], ["render", "24", "4", "0", "5", "/tmp/asan.mimg"], ["dumptext", "/tmp/asan2.txt"]):
        p = subprocess.run([hs, "text:/tmp/asan.txt", A] + args, capture_output=True, text=True)
        ran += 1
        if "ERROR: AddressSanitizer" in p.stderr or "runtime error" in p.stderr:
            bad += 1; print(seed, args[0], p.stderr[-600:])
print("runs", ran, "sanitizer reports", bad)
sys.exit(1 if bad else 0)
PY
# the tree builder runs subtrees on several threads: ThreadSanitizer over a 40 k-sphere field
g++ -std=c++17 -O1 -g -fsanitize=thread -ffp-contract=off -Imort_b200/csrc -Iinclude \
    tests/hostsim/hostsim.cpp mort_b200/csrc/scene.cpp mort_b200/csrc/scenes.cpp mort_b200/csrc/scene_text.cpp mort_b200/csrc/flatten.cpp \
    mort_b200/csrc/bvh_build.cpp -o /tmp/hostsim_tsan
/tmp/hostsim_tsan field:100 mort_b200/assets checkbvh 2>&1 | grep -E "WARNING: ThreadSanitizer" && bad=1
[ $bad = 0 ] && echo "asan/ubsan/tsan: clean"
exit $bad
