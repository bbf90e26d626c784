#!/bin/bash
# compute-sanitizer memcheck over small renders of every code path (one tool per gpurun call).
set -u
ROOT=$(pwd); OUT=$ROOT/gpurun_out/sanitize; mkdir -p $OUT
for args in "6 --width 48 --spp 16" "8 --width 40 --spp 4" "1 --width 64 --spp 9" "7 --width 40 --spp 9" "3 --width 64 --spp 4" "4 --width 64 --spp 4" \
            "6 --width 48 --spp 16 --mode wave" "8 --width 40 --spp 4 --mode wave" "1 --width 64 --spp 9 --stage 100000" "1 --field 30 --width 64 --spp 4"; do
  echo "== mort $args" | tee -a $OUT/memcheck.txt
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 mort_b200/mort $args 2>&1 | tail -4 | tee -a $OUT/memcheck.txt
done
grep -c "ERROR SUMMARY: 0 errors" $OUT/memcheck.txt
