set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_t7.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b7.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_fbank_gpu.py tests/test_blocks_gpu.py -m gpu -q -x -k "not random_shapes and not 40000 and not 38017" 2>&1 | tail -12 > gpurun_out/r2_memcheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_beats_gpu.py -m gpu -q -x -k "L2_2x1s or L2_2x2s_mask or pooled_hooks or predictor" 2>&1 | tail -12 > gpurun_out/r2_memcheck2.log
echo done
