set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_t10.log
timeout 120 python tools/bench_fbank.py > gpurun_out/r2_fbank4.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b10.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -c 1 -o gpurun_out/fbank_r2d python tools/bench_fbank.py > gpurun_out/r2_ncu_fbank4.log 2>&1
echo done
