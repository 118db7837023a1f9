# bench lines re-measured after two fixes (the probe's DRAM-bytes-per-load factor restored in profiles/ncu_traffic.json; threaded warm-up in the stage config)
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 12 --warmup 3 2> gpurun_out/r2u_bench.err > gpurun_out/r2u_bench.json
tail -3 gpurun_out/r2u_bench.err; cut -c1-300 gpurun_out/r2u_bench.json
timeout 1500 python bench.py --config se150_dp --steps 12 --warmup 3 2> gpurun_out/r2u_se150_dp.err > gpurun_out/r2u_se150_dp.json
cut -c1-400 gpurun_out/r2u_se150_dp.json
