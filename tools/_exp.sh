mkdir -p gpurun_out
L=gpurun_out/r2p_exp.log
: > $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py -q -m gpu -x -k "not cfg3" 2>&1 | tail -3
run() { env "$@" timeout 300 python tools/exp_resident.py >> $L 2>&1; }
run HGR_X=base
env HGR_X=base timeout 200 python tools/exp_resident.py 4096,21841,1024 >> $L 2>&1
grep "^\[" $L
timeout 200 python tools/timeline.py 512 21841 1024 2>&1 | tee gpurun_out/r2p_tl.log | grep -v "^  epi[0-9]" | grep -A30 "== null"
