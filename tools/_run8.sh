mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_configs.py -q -m gpu -x -k "om or OM" 2>&1 | tail -8
timeout 600 python tools/bench_om.py > gpurun_out/r2g_bench_om.json 2> gpurun_out/r2g_bench_om.err; tail -3 gpurun_out/r2g_bench_om.err; head -c 1500 gpurun_out/r2g_bench_om.json
