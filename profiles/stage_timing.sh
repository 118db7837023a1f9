S3_STAGE_TIMING=1 timeout 900 python bench.py --config pe100_deep --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -22
