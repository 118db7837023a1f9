timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_search_gpu.py tests/test_long_reads_gpu.py -x -q 2>&1 | tail -6
S3_STAGE_TIMING=1 timeout 900 python bench.py --config se150_dp --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -9 | cut -c1-600
