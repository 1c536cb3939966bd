# ncu --set full on the 16 depthwise kernels and the 1x1-conv GEMMs of one EfficientNet forward (512 x 5 s)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwconv_kernel -s 48 -c 16 -o gpurun_out/dw_r2 \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dw_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 96 -c 33 -o gpurun_out/c1_r2 \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1_r2.log 2>&1
echo done
