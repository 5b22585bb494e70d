for L in 2 1; do python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 3 --launches $L > gpurun_out/r2h_bench_L$L.json 2>> gpurun_out/r2h.err; done
for T in 128 192; do python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 3 --threads $T > gpurun_out/r2h_bench_T$T.json 2>> gpurun_out/r2h.err; done
