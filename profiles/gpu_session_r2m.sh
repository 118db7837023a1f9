#!/usr/bin/env bash
# Session r2m: the whole GPU tier, smoke, every bench config, the reference arm, launch list + one full ncu capture of a chain step.
set -uo pipefail
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 1000 > $OUT/${TAG}_clocks.csv &
SMI=$!
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== shim_check"; (cd oracle/_ref && timeout 300 ./shim_check shim_case) 2>&1 | tail -6 | tee $OUT/${TAG}_shim_check.txt
echo "== reference arm"; timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/${TAG}_bench_reference.err > $OUT/${TAG}_bench_reference.json
cut -c1-300 $OUT/${TAG}_bench_reference.json
echo "== bench"; timeout 1500 python bench.py 2> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
for cfg in se100_k4 se150_dp pe100_deep; do
  echo "== bench --config $cfg"; timeout 1500 python bench.py --config $cfg --steps 5 --warmup 3 2> $OUT/${TAG}_$cfg.err > $OUT/${TAG}_$cfg.json
  tail -3 $OUT/${TAG}_$cfg.err; cut -c1-330 $OUT/${TAG}_$cfg.json
done
echo "== ncu launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:s3_ -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_launches.csv
echo "== ncu full: one chain step"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:s3_(search|dp_|pe_|heavy|isbad)' -s 52 -c 26 \
    -f -o /tmp/${TAG}_chain python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_chain.log 2>&1
tail -2 $OUT/${TAG}_ncu_chain.log
ncu -i /tmp/${TAG}_chain.ncu-rep --page raw --csv > $OUT/${TAG}_chain_raw.csv 2>/dev/null
kill $SMI
ls -la $OUT | tail -24
