set -x
mkdir -p gpurun_out
AVEXK_DW_CH3=2 timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x -k "dwconv or forward or batch" 2>&1 | tail -4 > gpurun_out/r2_t41.log
for v in 4 2 4 2; do
  AVEXK_DW_CH3=$v timeout 60 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h41.log 2>&1
  echo "ch3=$v rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_h41.log | head -1)" >> gpurun_out/r2_h41_summary.log
done
cat gpurun_out/r2_h41_summary.log
AVEXK_DW_CH3=2 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 280 -c 160 --csv \
  --log-file gpurun_out/launches_effnet_r2s.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2s.log 2>&1
echo done
