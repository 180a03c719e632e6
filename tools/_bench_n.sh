# tools/_bench_n.sh N TAG : one class-sharded bench line at N GPUs -> gpurun_out/TAG_benchN.json (+ .err), summary on stdout
n=$1; tag=${2:-r2z}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 2> gpurun_out/${tag}_bench$n.err | grep '^{' > gpurun_out/${tag}_bench$n.json
tail -2 gpurun_out/${tag}_bench$n.err | grep -v "^\*\|OMP_NUM\|NCCL version"
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench$n.json'))
print("N=%d value %.1f M  ms/step %.4f  e2e %.1f M  e2e-fp32 %s  frac %.3f kernel %.2f us  1gpu %.1f M  speedup %.2f  hits %s" % (d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e_fp32_features'] and round(d['e2e_fp32_features']['value']/1e6,1), d['roofline']['frac'], d['roofline']['kernel_ms']*1e3, d['single_gpu_same_workload']['value']/1e6, d['speedup_vs_1gpu'], d['hits']))
print(d['impl']['lists']); print(d['clocks'])
PY
