python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench.err
