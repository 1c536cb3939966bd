# Round-2 evidence run: the commands whose outputs are summarised under profiles/ (tools/summarize_profiles.py).
set -x
mkdir -p gpurun_out
# 1. the bench line itself (CUDA-event timing, no profiler attached)
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final.log 2>&1
# 2. ncu launch list of the same command (time + DRAM bytes per launch; cold-cache and serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu_r2.log 2>&1
# 3. ncu --set full: the GEMM flavours of one layer (qkv, out_proj+LN, fc1+GELU, fc2+LN), the attention kernel, the fbank kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 10 -c 4 -o gpurun_out/gemm_r2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_gemm_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 12 -c 1 -o gpurun_out/attn_r2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_attn_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -s 3 -c 1 -o gpurun_out/fbank_r2 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_fbank_r2.log 2>&1
# 4. EfficientNet workload: bench line and launch list
timeout 300 python bench.py --workload effnet --steps 10 --warmup 3 > gpurun_out/bench_effnet_r2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv \
  --log-file gpurun_out/launches_effnet_r2.csv python bench.py --workload effnet --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/be_ncu_r2.log 2>&1

# 5. parity suite on this build, then compute-sanitizer racecheck on the kernels with hand-rolled cross-CTA / TMEM synchronisation
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gputest_r2_final.log
timeout 700 compute-sanitizer --tool racecheck python -m pytest tests/test_blocks_gpu.py -q -k "fused_layernorm_pooled or test_attention_gated" 2>&1 | tail -40 > gpurun_out/racecheck_r2_final.log
echo done2
