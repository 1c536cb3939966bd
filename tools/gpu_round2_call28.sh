# checkpoint: the whole GPU suite and the default bench line on the current build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gputest_r2_b.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_b.log 2>&1
tail -c 400 gpurun_out/bench_r2_b.log
echo done
