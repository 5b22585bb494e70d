python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_tests.log
for W in c4 c2; do for L in 1 2; do python bench.py --workload $W --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 --launches $L > gpurun_out/r2l_bench_${W}_L$L.json 2>> gpurun_out/r2l.err; done; done
python bench.py --workload c5 --steps 50 --warmup 5 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2l_bench_c5.json 2>> gpurun_out/r2l.err
