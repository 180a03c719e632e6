mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/exp_hostfed.py > gpurun_out/r2i_hostfed8.log 2>&1
grep "^N=" gpurun_out/r2i_hostfed8.log; tail -3 gpurun_out/r2i_hostfed8.log | grep -v "^N="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2i_bench8.json 2> gpurun_out/r2i_bench8.err
tail -2 gpurun_out/r2i_bench8.err; cat gpurun_out/r2i_bench8.json
