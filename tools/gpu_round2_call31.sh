# is the BEATs step slower on the final build, or was the GPU warm (bench after the 1-minute test suite)?  bench first on a fresh box
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_r2_c.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_r2_d.log 2>&1
echo done
