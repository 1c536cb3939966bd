# the default bench line on 2 GPUs of one box (weak + strong scaling step, EfficientNet at 256 clips per rank)
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_final2_n8.log 2>&1
tail -c 600 gpurun_out/bench_r2_final2_n8.log
echo done
