# last check of HEAD: build() + smoke() as the driver runs them, the GPU suite, one default bench line without extras
set -x
mkdir -p gpurun_out
timeout 900 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2_final.log 2>&1
tail -3 gpurun_out/smoke_r2_final.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/gputest_r2_head.log
cat gpurun_out/gputest_r2_head.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_head.log 2>&1
tail -c 300 gpurun_out/bench_r2_head.log
echo done
