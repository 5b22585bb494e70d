python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2n_bench_mb4.json 2>> gpurun_out/r2n.err
AAC_LIB=$PWD/multi_agent_aac_b200/libaac_env_mb3.so python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2n_bench_mb3.json 2>> gpurun_out/r2n.err
AAC_LIB=$PWD/multi_agent_aac_b200/libaac_env_mb3.so python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 --launches 1 > gpurun_out/r2n_bench_mb3_L1.json 2>> gpurun_out/r2n.err
AAC_LIB=$PWD/multi_agent_aac_b200/libaac_env_mb3.so python bench.py --workload c5 --steps 50 --warmup 5 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2n_bench_mb3_c5.json 2>> gpurun_out/r2n.err
