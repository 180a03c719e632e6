#!/bin/bash
# Run on the GPU box (under gpurun).  Produces, under gpurun_out/:
#   launches_<tag>.csv        ncu launch list of a short bench.py run (gpu__time_duration only)
#   full_<tag>.ncu-rep/.csv   one `--set full` capture of our three kernels (first launches after warm-up)
# Numbers printed by bench.py under ncu are never bench values.
tag=${1:-r01}
extra=${2:-}
out=gpurun_out
mkdir -p $out
BENCH="python bench.py --steps 40 --warmup 3 --no-cpu-baseline --streams 1 $extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv $BENCH > $out/ncu_list_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'score_umma|topk_merge|normalize' --launch-skip 12 -c 6 \
  -o $out/full_$tag -f $BENCH > $out/ncu_full_$tag.log 2>&1
ncu -i $out/full_$tag.ncu-rep --page raw --csv > $out/full_$tag.csv 2>/dev/null
ls -la $out | grep $tag
