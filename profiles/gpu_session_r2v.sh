#!/usr/bin/env bash
# Last session of round 2 on the B200 box: the whole GPU tier and smoke on the final tree, the default bench line, the chain-only line and
# config 2 with four batches in flight.  usage: bash profiles/gpu_session_r2v.sh
set -uo pipefail
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/r2v_pytest.log
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/r2v_smoke.log
echo "== bench (default: config 4 whole)"
timeout 1500 python bench.py 2> $OUT/r2v_bench.err > $OUT/r2v_bench.json
tail -4 $OUT/r2v_bench.err; cut -c1-300 $OUT/r2v_bench.json
for CFG in pe100_chain se100_k4; do
  echo "== bench --config $CFG"
  timeout 1500 python bench.py --config $CFG 2> $OUT/r2v_$CFG.err > $OUT/r2v_$CFG.json
  cut -c1-300 $OUT/r2v_$CFG.json
done
