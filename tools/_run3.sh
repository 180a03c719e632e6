mkdir -p gpurun_out
HGR_TL_IMPL=12 timeout 300 python tools/timeline.py 512 21841 1024 > gpurun_out/r2c_tl.log 2>&1
grep -v "^  epi[0-9]" gpurun_out/r2c_tl.log | head -34
timeout 900 python tools/exp_sketch.py > gpurun_out/r2b_sketch.log 2>&1
grep -v "^ok" gpurun_out/r2b_sketch.log | tail -30; grep -c "^ok" gpurun_out/r2b_sketch.log; grep "^ok   B=512 C=21841\|^ok   B=4096 C=2731\|^ok   B=4096 C=21841" gpurun_out/r2b_sketch.log
