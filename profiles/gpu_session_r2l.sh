#!/usr/bin/env bash
# Session r2l: long-read tests, the two widened bench configs, and the L2 fetch-granularity experiment.
set -uo pipefail
TAG=${1:-r2l}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest (long reads, stages, chains)"
timeout 1200 python -m pytest tests/test_long_reads_gpu.py tests/test_pe_chain_gpu.py tests/test_stages_gpu.py -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log
echo "== bench se150_dp"; timeout 1500 python bench.py --config se150_dp --steps 4 --warmup 3 2> $OUT/${TAG}_se150_dp.err > $OUT/${TAG}_se150_dp.json
tail -5 $OUT/${TAG}_se150_dp.err; cut -c1-700 $OUT/${TAG}_se150_dp.json
echo "== bench pe100_deep"; timeout 1500 python bench.py --config pe100_deep --steps 4 --warmup 3 2> $OUT/${TAG}_pe100_deep.err > $OUT/${TAG}_pe100_deep.json
tail -5 $OUT/${TAG}_pe100_deep.err; cut -c1-700 $OUT/${TAG}_pe100_deep.json
for g in 32 64; do
  echo "== L2 fetch granularity $g"
  S3_L2_FETCH_GRANULARITY=$g timeout 900 python bench.py --no-cpu-baseline --steps 4 --warmup 3 2> $OUT/${TAG}_fetch$g.err > $OUT/${TAG}_fetch$g.json
  tail -2 $OUT/${TAG}_fetch$g.err; cut -c1-330 $OUT/${TAG}_fetch$g.json
  python - <<PY
import json
d = json.load(open("$OUT/${TAG}_fetch$g.json"))
print("value", d["value"], "search ms", d["search"]["ms_per_launch"], "random loads/s", d["search"]["random_loads_per_s"], "stages", d["stages_ms_per_step"])
PY
done
ls -la $OUT | tail -8
