mkdir -p gpurun_out
timeout 900 python tools/exp_sketch.py > gpurun_out/r2b_sketch.log 2>&1
tail -50 gpurun_out/r2b_sketch.log
