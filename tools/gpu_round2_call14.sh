set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_effnet_gpu.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_t14.log
for o in 2 3 4; do
  AVEXK_DW_OCC=$o timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be_occ$o.log 2>&1
done
echo done
