timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_long_reads_gpu.py -x -q 2>&1 | tail -4
S3_STAGE_TIMING=1 timeout 900 python bench.py --config se150_dp --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -14 | cut -c1-900
S3_STAGE_TIMING=1 timeout 900 python bench.py --config pe100_deep --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -24 | cut -c1-900
