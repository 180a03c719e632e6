# ncu captures of the class-sharded head in global-certificate mode (1 GPU, G logical ranks in one process):
#  * scoring kernel + producer-side merge + owner-side certified merge at the N = 8 arrangement of cfg 5
#  * the scoring kernel alone (narrow lists of the global certificate) at the N = 8 / 4 / 2 shards: DRAM traffic, tensor pipe
mkdir -p gpurun_out
# per batch the filter sees (score, select) x 8 producers, then select x 8 owners: skip batch 1 (24), take rank 0's pair ...
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'score_umma_pair|topk_select' --launch-skip 24 -c 2 \
  -o gpurun_out/r02b_cert_prod -f python tools/cert_once.py 4096 21841 1024 8 > gpurun_out/r02b_cert_prod.log 2>&1
# ... and the first owner-side merge of batch 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'topk_select' --launch-skip 24 -c 1 \
  -o gpurun_out/r02b_cert_owner -f python tools/cert_once.py 4096 21841 1024 8 > gpurun_out/r02b_cert_owner.log 2>&1
for f in r02b_cert_prod r02b_cert_owner; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null; done
for G in 8 4 2; do
  timeout 300 ncu --set full --clock-control none -k regex:'score_umma_pair' --launch-skip 4 -c 1 -o gpurun_out/r02b_shard_n$G -f \
    python tools/cert_once.py 4096 21841 1024 $G lists-only > gpurun_out/r02b_shard_n$G.log 2>&1
  ncu -i gpurun_out/r02b_shard_n$G.ncu-rep --page raw --csv > gpurun_out/r02b_shard_n$G.csv 2>/dev/null
done
tail -1 gpurun_out/r02b_*.log
ls -la gpurun_out | grep r02b
