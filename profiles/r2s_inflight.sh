S3_STAGE_TIMING=1 S3_IN_FLIGHT=1 timeout 600 python bench.py --config pe100_deep --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep "s3_pe_deep_dp\]" | tail -12
for T in 2 4; do
S3_IN_FLIGHT=$T timeout 600 python bench.py --config pe100_deep --steps 12 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2s_deep_T$T.err | tail -1 > gpurun_out/r2s_deep_T$T.json
python -c "
import json; d=json.loads(open('gpurun_out/r2s_deep_T$T.json').read()); print('T=$T', d['value'], d['ms_per_step'], d['value_one_batch_in_flight'], d['stages_ms_per_step'], d['pipeline'])"
done
