mkdir -p gpurun_out
HGR_TL_IMPL=12 timeout 300 python tools/timeline.py 512 21841 1024 > gpurun_out/r2c_tl.log 2>&1
grep "warp 2: cr\|warp 2: sel\|scan loop\|cycles" gpurun_out/r2c_tl.log | grep -v "MMA\|tmem ld" | head -12
HGR_TL_IMPL=12 timeout 300 python tools/timeline.py 4096 21841 1024 > gpurun_out/r2c_tl2.log 2>&1
grep "warp 2: cr\|warp 2: sel\|scan loop\|cycles" gpurun_out/r2c_tl2.log | grep -v "MMA\|tmem ld" | head -12
timeout 900 python tools/exp_sketch.py > gpurun_out/r2b_sketch.log 2>&1
grep -v "^ok" gpurun_out/r2b_sketch.log | tail -30; grep -c "^ok" gpurun_out/r2b_sketch.log
