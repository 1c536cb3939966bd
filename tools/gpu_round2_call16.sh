# TMA sliding-window depthwise kernel: parity, A/B against the per-output-row kernel, launch list, ncu of 5 dw + 6 GEMM launches
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t16.log
timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be16_new.log 2>&1
AVEXK_DW_OLD=1 timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be16_old.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/launches_effnet_r2b.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv_tma -s 48 -c 5 -o gpurun_out/dw_r2b \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dw_r2b.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 96 -c 6 -o gpurun_out/c1_r2b \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1_r2b.log 2>&1
ls -la gpurun_out
echo done
