set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_t39.log
for i in 1 2; do
  timeout 60 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h39_$i.log 2>&1
  echo "slab run $i rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_h39_$i.log | head -3 | tr '\n' ' ')" >> gpurun_out/r2_h39_summary.log
done
AVEXK_DW_NOSLAB=1 timeout 60 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h39_3.log 2>&1
echo "noslab run rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_h39_3.log | head -3 | tr '\n' ' ')" >> gpurun_out/r2_h39_summary.log
cat gpurun_out/r2_h39_summary.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 280 -c 160 --csv \
  --log-file gpurun_out/launches_effnet_r2r.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2r.log 2>&1
echo done
