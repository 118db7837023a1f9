#!/usr/bin/env bash
# Short B200 session between full ones: GPU parity tests + one bench line.
# usage: bash profiles/gpu_quick.sh <tag> [extra bench args]     (outputs -> gpurun_out/<tag>_*)
set -uo pipefail
TAG=${1:-q}; shift || true
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 "$@" 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
tail -6 $OUT/${TAG}_bench.err
