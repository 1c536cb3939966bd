set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_beats_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t6.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 1 -o gpurun_out/attn_r2d python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_attn4.log 2>&1
echo done
