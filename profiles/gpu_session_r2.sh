#!/usr/bin/env bash
# Round-2 session on the B200 box (through gpurun): tests, bench lines, the ncu launch list of the bench command and one
# `ncu --set full` capture of a whole chain step.  usage: bash profiles/gpu_session_r2.sh <tag> [quick]
set -uo pipefail
TAG=${1:-r2}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
if [ -z "$QUICK" ]; then
  echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.log
  echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
fi
echo "== bench (with the L2 access-policy experiment)"
S3_L2_EXPERIMENT=1 timeout 1500 python bench.py --steps 5 --warmup 3 2> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench.json
tail -4 $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
if [ -z "$QUICK" ]; then
  echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench_reference.json
  cut -c1-300 $OUT/${TAG}_bench_reference.json
fi
echo "== ncu launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:s3_ -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_launches.csv
echo "== ncu full: one chain step (the kernels of the third step: two steps' launches skipped)"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:s3_(search|dp_|pe_|heavy|isbad)' -s 52 -c 26 \
    -f -o /tmp/${TAG}_chain python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_chain.log 2>&1
tail -2 $OUT/${TAG}_ncu_chain.log
# the reports are tens of MB (gpurun brings back 64 MiB at most): their raw pages as CSV, and the source page of the DP sweep
ncu -i /tmp/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_chain.ncu-rep --page source --csv --kernel-name regex:s3_dp_sweep16 > $OUT/${TAG}_sweep_source.csv 2>/dev/null
echo "== ncu full: search launch of config 2 (k <= 4)"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:s3_(search|heavy|isbad)' -s 12 -c 6 \
    -f -o /tmp/${TAG}_search_k4 python bench.py --config se100_k4 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_search_k4.log 2>&1
tail -2 $OUT/${TAG}_ncu_search_k4.log
ncu -i /tmp/${TAG}_search_k4.ncu-rep --page raw --csv > $OUT/${TAG}_search_k4_raw.csv 2>/dev/null
kill $SMI
ls -la $OUT | tail -20
