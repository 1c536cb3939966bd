set -x
mkdir -p gpurun_out
timeout 200 python tools/e2e_probe_effnet.py > gpurun_out/e2e_probe_effnet.log 2>&1
cat gpurun_out/e2e_probe_effnet.log | tail -12
echo done
