for i in 1 2; do
python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2w_new$i.json 2>> gpurun_out/r2w.err
AAC_LIB=$PWD/multi_agent_aac_b200/libv_prev.so python bench.py --steps 300 --warmup 10 --no-aux --no-cpu --e2e-steps 2 > gpurun_out/r2w_prev$i.json 2>> gpurun_out/r2w.err
done
