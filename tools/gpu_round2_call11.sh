set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_effnet_gpu.py tests/test_blocks_gpu.py -m gpu -q -k "effnet or dwconv or conv1x1 or melspec or fused_layernorm" 2>&1 | tail -6 > gpurun_out/r2_t9.log
for q in 64 128 256 512 1024; do
  AVEXK_DW_QUADS=$q timeout 200 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_be_q$q.log 2>&1
done
echo done
