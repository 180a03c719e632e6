# ncu captures of the production bench (1 GPU): launch list + full sections of the three kernels; shard shapes for traffic
mkdir -p gpurun_out
BENCH="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --streams 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_umma_pair|topk_select|topk_merge|normalize_rows' --launch-skip 12 -c 6 -o gpurun_out/r02_full -f $BENCH > gpurun_out/r02_ncu_full.log 2>&1
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full.csv 2>/dev/null
# shard shapes: DRAM traffic and tensor-pipe share of the scoring kernel at the N = 2 / 4 / 8 shards and cfg 5 / cfg 4
cat > /tmp/shard_once.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from hgrnet_b200 import ops, _cabi
B, C, D = (int(v) for v in sys.argv[1:4])
g = torch.Generator().manual_seed(1)
def emb(n):
    x = torch.randn(n, D, generator=g); return (x / x.norm(dim=-1, keepdim=True)).to(torch.bfloat16).cuda()
banks = [emb(C) for _ in range(3)]; xs = [emb(B) for _ in range(2)]
for i in range(6):
    ops.score_topk(xs[i % 2], banks[i % 3], K=20, impl=ops.HGR_IMPL_TCGEN05 | _cabi.HGR_IMPL_FLAG_NO_MERGE)
torch.cuda.synchronize()
PY
for shape in "4096 2731 1024" "4096 5461 1024" "4096 10921 1024" "4096 21841 1024" "1024 10450 512"; do
  tag=$(echo $shape | tr ' ' 'x')
  timeout 300 ncu --set full --clock-control none -k regex:'score_umma_pair' --launch-skip 4 -c 1 -o gpurun_out/r02_shard_$tag -f python /tmp/shard_once.py $shape > gpurun_out/r02_shard_$tag.log 2>&1
  ncu -i gpurun_out/r02_shard_$tag.ncu-rep --page raw --csv > gpurun_out/r02_shard_$tag.csv 2>/dev/null
done
ls -la gpurun_out | grep r02
