#!/usr/bin/env bash
# What is run on the B200 box (through gpurun) to produce the files under profiles/.
# usage: bash profiles/gpu_session.sh <tag>      (outputs -> gpurun_out/<tag>_*)
set -uo pipefail
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py --steps 5 --warmup 3 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
tail -12 $OUT/${TAG}_bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_reference.json
echo "== ncu launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:s3_ -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -3 $OUT/${TAG}_launches.csv
echo "== ncu full: search"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:s3_search -s 7 -c 4 \
    -f -o $OUT/${TAG}_search python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_search.log 2>&1
echo "== ncu full: dp"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:s3_dp_ -s 3 -c 3 \
    -f -o $OUT/${TAG}_dp python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_dp.log 2>&1
kill $SMI
ls -la $OUT
