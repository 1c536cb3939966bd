# Final round-2 evidence: the commands whose outputs are summarised under profiles/ (tools/summarize_profiles.py r2_final ...).
set -x
mkdir -p gpurun_out
# 0. parity suite on this build
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/gputest_r2_final2.log
# 1. the bench line itself (CUDA-event timing, no profiler attached) and the EfficientNet line
timeout 700 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final2.log 2>&1
timeout 300 python bench.py --workload effnet --steps 10 --warmup 3 > gpurun_out/bench_effnet_r2_final.log 2>&1
# 2. launch lists (time + DRAM bytes per launch; cold-cache and serialised: compare shares)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 276 -c 120 --csv \
  --log-file gpurun_out/launches_effnet_r2_final.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu_r2_final.log 2>&1
# 3. ncu --set full (no source import: small reports) of the EfficientNet kernels of one forward: 16 depthwise, 31 pointwise, mel, stem
timeout 600 ncu --set full --clock-control none -k regex:dwconv_tma -s 48 -c 16 -o gpurun_out/dw_r2_final \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dw_r2_final.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:pointwise_kernel -s 93 -c 8 -o gpurun_out/pw_r2_final \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_pw_r2_final.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"melspec_kernel|stem_kernel" -s 6 -c 2 -o gpurun_out/mel_r2_final \
  python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mel_r2_final.log 2>&1
# 4. parity suite on this build, stress test, compute-sanitizer on the new EfficientNet kernels
timeout 300 python tools/stress_pointwise.py 20 > gpurun_out/stress_r2_final.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_effnet_gpu.py -q -k "conv1x1 or dwconv or melspec" 2>&1 | tail -25 > gpurun_out/memcheck_effnet_r2_final.log
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_effnet_gpu.py -q -k "conv1x1 or dwconv" 2>&1 | tail -25 > gpurun_out/synccheck_effnet_r2_final.log
ls -la gpurun_out | tail -30
echo done
