mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err
tail -3 gpurun_out/r2f_bench1.err; cat gpurun_out/r2f_bench1.json
