set -x
mkdir -p gpurun_out
timeout 300 python tools/stress_pointwise.py 5 > gpurun_out/r2_stress38.log 2>&1
echo "rc=$?" >> gpurun_out/r2_stress38.log
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_t38.log
for i in 1 2; do
  timeout 60 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h38_$i.log 2>&1
  echo "run $i rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_h38_$i.log | head -3 | tr '\n' ' ')" >> gpurun_out/r2_h38_summary.log
done
cat gpurun_out/r2_h38_summary.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 280 -c 160 --csv \
  --log-file gpurun_out/launches_effnet_r2q.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2q.log 2>&1
echo done
