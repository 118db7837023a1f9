S3_IN_FLIGHT=4 timeout 900 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2s_bench_T4.err | tail -1 > gpurun_out/r2s_bench_T4.json
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_T4.json').read()); print('T=4', d['value'], d['ms_per_step'], d['value_one_batch_in_flight'], d['e2e']['value'], d['e2e']['pageable_value'])"
for T in 2 4; do
S3_IN_FLIGHT=$T timeout 900 python bench.py --config se150_dp --steps 12 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2s_se150_T$T.err | tail -1 > gpurun_out/r2s_se150_T$T.json
python -c "
import json; d=json.loads(open('gpurun_out/r2s_se150_T$T.json').read()); print('se150 T=$T', d['value'], d['ms_per_step'], d['value_one_batch_in_flight'], d['stages_ms_per_step'])"; done
S3_STAGE_TIMING=1 S3_IN_FLIGHT=1 timeout 600 python bench.py --config se150_dp --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "s3_single_dp_align\]" | tail -7
