set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t0.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_b0.log 2>&1
timeout 120 python tools/bench_fbank.py > gpurun_out/r2_fbank0.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -c 1 -o gpurun_out/fbank_r2a python tools/bench_fbank.py > gpurun_out/r2_ncu_fbank.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_blocks_gpu.py -q -k "fused_layernorm or attention" -x 2>&1 | tail -30 > gpurun_out/r2_racecheck.log
echo done
