# HEAD check after the raw-output change: GPU suite, memcheck + synccheck on the EfficientNet kernel tests, smoke
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gputest_r2_head2.log
cat gpurun_out/gputest_r2_head2.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_effnet_gpu.py -q -k "conv1x1 or dwconv or melspec" 2>&1 | tail -4 > gpurun_out/memcheck_effnet_r2_head2.log
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_effnet_gpu.py -q -k "conv1x1 or dwconv" 2>&1 | tail -4 > gpurun_out/synccheck_effnet_r2_head2.log
cat gpurun_out/memcheck_effnet_r2_head2.log gpurun_out/synccheck_effnet_r2_head2.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo done
