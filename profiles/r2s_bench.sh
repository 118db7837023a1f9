for T in 4 6 8; do
S3_IN_FLIGHT=$T timeout 900 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2s_bench_T$T.err | tail -1 > gpurun_out/r2s_bench_T$T.json
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_T$T.json').read()); print('T=$T', d['value'], d['ms_per_step'], d['value_one_batch_in_flight'], d['e2e']['value'], d['e2e']['pageable_value'])"; done
nproc; free -g | head -2
