#!/bin/bash
# Round-2 (final kernel) profiling passes on the GPU box.  Output: gpurun_out/prof_r2b/*
#  1. ncu launch list of the benchmark command itself (timed region only, one bench step)
#  2. ncu --set full of one grouped W = 10000 conv launch of that step
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/prof_r2b; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp32-grade --profile-range"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv $B > $OUT/bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches_bench.csv)"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tpconv_umma -s 3 -c 1 -o $OUT/prof_umma $B > $OUT/ncu_umma.log 2>&1
echo "umma capture rc=$?"
ls -la $OUT
