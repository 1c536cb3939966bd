set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_t3.log
timeout 120 python tools/bench_fbank.py > gpurun_out/r2_fbank3.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 1 -o gpurun_out/attn_r2a python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_attn.log 2>&1
echo done
