#!/usr/bin/env bash
# ncu --set full of one launch of each DP kernel inside a short bench run -> gpurun_out/<tag>_dp.ncu-rep
TAG=${1:-dp}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:s3_dp_ -s 3 -c 3 \
    -f -o $OUT/${TAG}_dp python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_dp.log 2>&1
tail -2 $OUT/${TAG}_ncu_dp.log
