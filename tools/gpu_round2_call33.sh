set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointwise_kernel -s 95 -c 1 -o gpurun_out/pw_r2e \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pw_r2e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma -s 60 -c 1 -o gpurun_out/dw_r2e \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dw_r2e.log 2>&1
cp avex_b200/_build/pointwise.o gpurun_out/pointwise_r2e.o; cp avex_b200/_build/effnet.o gpurun_out/effnet_r2e.o
echo done
