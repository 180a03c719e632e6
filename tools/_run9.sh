mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py -q -m gpu -x 2>&1 | tail -5
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2h_bench$n.json 2> gpurun_out/r2h_bench$n.err
tail -3 gpurun_out/r2h_bench$n.err; cat gpurun_out/r2h_bench$n.json
done
