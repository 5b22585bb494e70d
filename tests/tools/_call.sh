python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_tests.log
python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2t_bench.json 2>> gpurun_out/r2t.err
