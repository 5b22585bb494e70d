python -m pytest tests -m gpu -x -q -k "sensor" > gpurun_out/r2k_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
