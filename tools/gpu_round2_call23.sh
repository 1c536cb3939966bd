set -x
mkdir -p gpurun_out
timeout 300 python tools/stress_pointwise.py 40 > gpurun_out/r2_stress23.log 2>&1
echo "rc=$?" >> gpurun_out/r2_stress23.log
timeout 500 compute-sanitizer --tool synccheck python tools/stress_pointwise.py 1 > gpurun_out/r2_stress23_sync.log 2>&1
echo "rc=$?" >> gpurun_out/r2_stress23_sync.log
tail -c 1500 gpurun_out/r2_stress23_sync.log
echo done
