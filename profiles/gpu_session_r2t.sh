#!/usr/bin/env bash
# Final round-2 session on the B200 box (through gpurun): GPU tests, smoke, the bench line of every config, the reference arm, the ncu
# launch list of the bench command and one `ncu --set full` capture of a whole step (chain + deep DP).  usage: bash profiles/gpu_session_r2t.sh <tag>
set -uo pipefail
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== shim_check"; [ -x oracle/_ref/shim_check ] && (cd oracle/_ref && timeout 300 ./shim_check shim_case 2>&1 | tail -12) | tee $OUT/${TAG}_shim_check.txt
echo "== bench (default: config 4 whole)"
timeout 1500 python bench.py --steps 12 --warmup 3 2> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench.json
tail -4 $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> $OUT/${TAG}_bench_reference.err > $OUT/${TAG}_bench_reference.json
cut -c1-300 $OUT/${TAG}_bench_reference.json
for CFG in pe100_chain se100_k4 se150_dp; do
  echo "== bench --config $CFG"
  timeout 1500 python bench.py --config $CFG --steps 12 --warmup 3 2> $OUT/${TAG}_$CFG.err > $OUT/${TAG}_$CFG.json
  cut -c1-300 $OUT/${TAG}_$CFG.json
done
echo "== ncu launch list (one batch in flight, so that the launches of a step follow each other)"
S3_IN_FLIGHT=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:s3_|Device" -c 3000 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_launches.csv
echo "== ncu full: one whole step (chain + deep DP)"
S3_IN_FLIGHT=1 timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:s3_(search|dp_|pe_|heavy|isbad|stage|pair|seed|csr)' -s 125 -c 125 \
    -f -o /tmp/${TAG}_chain python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_chain.log 2>&1
tail -2 $OUT/${TAG}_ncu_chain.log
ncu -i /tmp/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
kill $SMI
ls -la $OUT | tail -24
