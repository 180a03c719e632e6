mkdir -p gpurun_out
: > gpurun_out/r2d_var.log
for v in nogate minimal; do
  lib=$PWD/hgrnet_b200/lib/var_$v.so
  echo "=== variant ${v:-base}" >> gpurun_out/r2d_var.log
  HGR_LIB=$lib HGR_TL_IMPL=12 timeout 300 python tools/timeline.py 512 21841 1024 2>&1 | grep "scan loop\|cycles\|CTA lifetime\|warp 2" | head -22 >> gpurun_out/r2d_var.log
  HGR_LIB=$lib HGR_TL_IMPL=12 timeout 300 python tools/timeline.py 4096 21841 1024 2>&1 | grep "scan loop\|cycles\|CTA lifetime\|warp 2" | head -22 >> gpurun_out/r2d_var.log
done
cat gpurun_out/r2d_var.log
