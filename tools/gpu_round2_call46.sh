set -x
mkdir -p gpurun_out
timeout 400 python bench.py --workload effnet --steps 10 --warmup 3 > gpurun_out/bench_effnet_r2_final3.log 2>&1
tail -c 1200 gpurun_out/bench_effnet_r2_final3.log
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final3.log 2>&1
tail -c 300 gpurun_out/bench_r2_final3.log
echo done
