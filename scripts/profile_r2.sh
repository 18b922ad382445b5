#!/bin/bash
# Round-2 profiling passes (GPU box).  Output: gpurun_out/prof_r2/*
#  1. ncu launch list of the benchmark command itself (timed region only, one bench step)
#  2. ncu --set full of one grouped W = 10000 conv launch of that step, and of the graph-construction / HBM-type kernels
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/prof_r2; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp32-grade --profile-range"
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv $B > $OUT/bench_under_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches_bench.csv)"
# full capture: the 4th conv launch of the step is a W = 10000 layer (layers 3-5); -s skips earlier ones
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tpconv_umma -s 3 -c 1 -o $OUT/prof_umma $B > $OUT/ncu_umma.log 2>&1
echo "umma capture rc=$?"
for k in radius_scan knn_scan_filter edge_embed_fold node_update_multi pose_update; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -s 4 -c 1 -o $OUT/prof_$k $B > $OUT/ncu_$k.log 2>&1
  echo "$k capture rc=$?"
done
ls -la $OUT
