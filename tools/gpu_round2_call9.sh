set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t8.log
timeout 300 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be8.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_effnet_r2a.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be8n.log 2>&1
echo done
