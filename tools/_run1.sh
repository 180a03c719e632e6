mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_env.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
timeout 600 python tools/sweep.py > gpurun_out/r2a_sweep.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2>&1
timeout 200 python tools/timeline.py 512 21841 1024 > gpurun_out/r2a_tl.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_sweep.log; tail -2 gpurun_out/r2a_bench.log
