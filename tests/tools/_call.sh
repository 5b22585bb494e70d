python bench.py > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_c3_k20.json 2>> gpurun_out/bench_r2_c3.err
python bench.py --workload c2 --no-cpu > gpurun_out/bench_r2_c2.json 2>> gpurun_out/bench_r2_c3.err
python bench.py --workload c4 --no-cpu > gpurun_out/bench_r2_c4.json 2>> gpurun_out/bench_r2_c3.err
python bench.py --workload c5 --steps 200 --no-cpu > gpurun_out/bench_r2_c5.json 2>> gpurun_out/bench_r2_c3.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_ref.json 2>> gpurun_out/bench_r2_c3.err
python bench.py --workload c1 > gpurun_out/bench_r2_c1.json 2>> gpurun_out/bench_r2_c3.err
