# pointwise (memory-bound 1x1 conv) kernel + slice-concurrent depthwise tiling: parity, A/B, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t17.log
timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be17_new.log 2>&1
AVEXK_PW=0 timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be17_nopw.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/launches_effnet_r2c.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2c.log 2>&1
echo done
