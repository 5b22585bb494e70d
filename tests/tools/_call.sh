python tests/tools/sanitize.py env > gpurun_out/r2y_sanitize_plain.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/tools/sanitize.py env > gpurun_out/r2y_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2y_memcheck.log
