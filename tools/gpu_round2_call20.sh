set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t20.log
timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be20.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 280 -c 140 --csv \
  --log-file gpurun_out/launches_effnet_r2e.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pointwise_kernel -s 93 -c 2 -o gpurun_out/pw_r2c \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pw_r2c.log 2>&1
cp avex_b200/_build/pointwise.o gpurun_out/pointwise_r2c.o
echo done
