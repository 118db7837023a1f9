OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:s3_ -c 400 --csv --log-file $OUT/r01k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/r01k_ncu_bench.log 2>&1
for b in 64 128 512 -1; do
  line=$(S3_SPLIT_BUDGET=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>$OUT/r01k_var.err | tail -1)
  echo "{\"budget\": $b, \"bench\": $line}" >> $OUT/r01k_budget.jsonl
  echo "budget $b: $(echo "$line" | python -c 'import sys,json; b=json.loads(sys.stdin.read()); print("value %.1fM e2e %.1fM search %.2f ms dp %.2f ms" % (b["value"]/1e6, b["e2e"]["value"]/1e6, b["search"]["ms_per_launch"], b["dp"]["ms_per_step"]))')"
done
