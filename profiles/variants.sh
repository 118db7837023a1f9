#!/usr/bin/env bash
# Tuning experiments: bench.py against variant builds of the library (soap3-dp_b200/build.py
# --variant=<name> -D...), one JSON line each -> gpurun_out/<tag>_variants.jsonl
# usage: bash profiles/variants.sh <tag> <variant> [<variant> ...]     ("base" = the product library)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_variants.jsonl
for v in "$@"; do
  if [ "$v" = base ]; then unset S3_LIB_PATH; else export S3_LIB_PATH=$PWD/soap3-dp_b200/libsoap3dp_b200.$v.so; fi
  line=$(timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>$OUT/${TAG}_variants.err | tail -1)
  echo "{\"variant\": \"$v\", \"bench\": $line}" >> $OUT/${TAG}_variants.jsonl
  echo "$v: $(echo "$line" | python -c 'import sys,json; b=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM search %.2f ms (frac %.3f) dp %.2f ms (%.0f GCUPS)" % (b["value"]/1e6, b["e2e"]["value"]/1e6, b["search"]["ms_per_launch"], b["roofline"]["frac"], b["dp"]["ms_per_step"], b["dp"]["gcups"]))')"
done
