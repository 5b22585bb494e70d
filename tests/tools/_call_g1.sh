#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g1_tests.log
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --no-aux --e2e-steps 2"
timeout 300 $CMD > gpurun_out/g1_c3.json 2> gpurun_out/g1_c3.err; rc1=$?
timeout 300 $CMD --launches 2 > gpurun_out/g1_c3_L2.json 2> gpurun_out/g1_c3_L2.err; rc2=$?
if [ $rc1 -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/g1_launches_c3.csv $CMD > gpurun_out/g1_ncu_list.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 30 -c 1 -f -o gpurun_out/g1_phased $CMD > gpurun_out/g1_ncu_phased.log 2>&1
fi
if [ $rc2 -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 30 -c 2 -f -o gpurun_out/g1_two $CMD --launches 2 > gpurun_out/g1_ncu_two.log 2>&1
fi
tail -3 gpurun_out/g1_tests.log
