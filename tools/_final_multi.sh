mkdir -p gpurun_out
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/r2t_bench$n.err | grep '^{' > gpurun_out/r2t_bench$n.json
tail -2 gpurun_out/r2t_bench$n.err | grep -v "^\*\|OMP_NUM\|NCCL version"
python - <<PY
import json
d=json.load(open('gpurun_out/r2t_bench$n.json'))
print("N=%d value %.1f M  ms/step %.4f  e2e %.1f M  e2e-fp32 %s  frac %.3f  1gpu %.1f M  speedup %.2f  %s" % (d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e_fp32_features'] and round(d['e2e_fp32_features']['value']/1e6,1), d['roofline']['frac'], d['single_gpu_same_workload']['value']/1e6, d['speedup_vs_1gpu'], d['impl']['pipeline'][:50]))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 --impl reference | head -c 300
