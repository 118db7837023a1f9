#!/usr/bin/env bash
# Session r2k: the CPU arm on the reference's real CPU search (full-size index), the bench line with the redefined search
# roofline, and ncu of the random-sector probe itself (DRAM bytes per random 32-byte load).
set -uo pipefail
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
echo "== reference arm"; timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/${TAG}_bench_reference.err > $OUT/${TAG}_bench_reference.json
tail -3 $OUT/${TAG}_bench_reference.err; cut -c1-600 $OUT/${TAG}_bench_reference.json
echo "== bench"; timeout 1500 python bench.py 2> $OUT/${TAG}_bench.err > $OUT/${TAG}_bench.json
tail -6 $OUT/${TAG}_bench.err; cut -c1-400 $OUT/${TAG}_bench.json
echo "== ncu of the probe"
S3_PROBE_ONLY=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__sectors_read.sum,lts__t_sectors_op_read.sum --clock-control none \
    -k regex:s3_random_sector --csv --log-file $OUT/${TAG}_probe_ncu.csv python bench.py > $OUT/${TAG}_probe.log 2>&1
tail -3 $OUT/${TAG}_probe.log; tail -30 $OUT/${TAG}_probe_ncu.csv
ls -la $OUT | tail -8
