#!/bin/bash
# The GPU-side half of the evidence under profiles/ (run from the repo root under gpurun, ONE GPU):
#   gpurun --timeout 3000 -- 'bash tests/tools/capture.sh'
# then, here:  ncu -i gpurun_out/cap_phased.ncu-rep --page raw --csv > /tmp/raw.csv
#              python tests/tools/ncu_summary.py /tmp/raw.csv profiles/r2_env_kernel_v2_phased_ncu_full_summary.csv
#              (the same with cap_two.ncu-rep and the _step / _reset summary names; tests/tools/hotspots.py for the per-line view)
# Every ncu pass runs only after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --no-aux --e2e-steps 2"
timeout 300 $CMD > gpurun_out/cap_c3.json 2> gpurun_out/cap_c3.err; rc1=$?
timeout 300 $CMD --launches 2 > gpurun_out/cap_c3_two.json 2> gpurun_out/cap_c3_two.err; rc2=$?
if [ $rc1 -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/cap_launches_c3.csv $CMD > gpurun_out/cap_ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 30 -c 1 -f -o gpurun_out/cap_phased $CMD > gpurun_out/cap_ncu_phased.log 2>&1
fi
if [ $rc2 -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 30 -c 2 -f -o gpurun_out/cap_two $CMD --launches 2 > gpurun_out/cap_ncu_two.log 2>&1
fi
