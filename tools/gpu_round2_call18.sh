set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointwise_kernel -s 93 -c 2 -o gpurun_out/pw_r2a \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pw_r2a.log 2>&1
echo done
