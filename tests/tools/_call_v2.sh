#!/bin/bash
mkdir -p gpurun_out
for w in c4 c2; do for L in 1 3 2; do
timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu --no-aux --e2e-steps 2 --launches $L > gpurun_out/v2_${w}_L$L.json 2>&1
done; done
