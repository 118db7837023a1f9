# experiment: bench.py against variant builds / env settings; usage: bash profiles/exp_variants.sh <tag> name[:ENV=VAL] ...
# a name with a matching soap3-dp_b200/libsoap3dp_b200.<name>.so loads that library; "base" is the product build
OUT=gpurun_out; mkdir -p $OUT; TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_variants.jsonl
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=${spec#*:}
  lib=$PWD/soap3-dp_b200/libsoap3dp_b200.$name.so
  if [ -f "$lib" ]; then envs="$envs S3_LIB_PATH=$lib"; fi
  line=$(env $envs S3_X=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>$OUT/${TAG}_var.err | tail -1)
  echo "{\"variant\": \"$spec\", \"bench\": $line}" >> $OUT/${TAG}_variants.jsonl
  echo "$spec: $(echo "$line" | python -c 'import sys,json; b=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM search %.2f ms dp %.2f ms" % (b["value"]/1e6, b["e2e"]["value"]/1e6, b["search"]["ms_per_launch"], b["dp"]["ms_per_step"]))')"
done
