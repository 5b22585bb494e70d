python bench.py --gpus 8 --steps 200 --warmup 5 --no-cpu > gpurun_out/bench_r2_c3_8gpu.json 2> gpurun_out/r2x_8gpu.err; echo "rc=$?" >> gpurun_out/r2x_8gpu.err
python bench.py --gpus 8 --workload c5 --steps 100 --warmup 5 --no-cpu > gpurun_out/bench_r2_c5_8gpu.json 2>> gpurun_out/r2x_8gpu.err; echo "rc=$?" >> gpurun_out/r2x_8gpu.err
