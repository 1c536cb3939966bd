set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_effnet_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2_t44.log
timeout 300 python tools/stress_pointwise.py 3 > gpurun_out/r2_stress44.log 2>&1
for i in 1 2; do
  timeout 60 python bench.py --workload effnet --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_h44.log 2>&1
  echo "run $i rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_h44.log | head -1)" >> gpurun_out/r2_h44_summary.log
done
cat gpurun_out/r2_h44_summary.log
timeout 200 python - > gpurun_out/r2_hooks44.log 2>&1 <<'PY'
# EfficientNet forward with all 17 hooks requested (raw pre-BN outputs): time per 512 x 5 s batch
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from avex_b200 import plugin
from avex_b200.plugin import efficientnet_model  # noqa: F401
spec = plugin.ModelSpec(name="efficientnet", device="cuda", efficientnet_variant="b0",
                        audio_config=dict(sample_rate=16000, n_fft=800, hop_length=160, win_length=800, window="hann", n_mels=128,
                                          representation="mel_spectrogram", normalize=True, target_length_seconds=10, window_selection="random"))
model = plugin.build_model_from_spec(spec, "cuda", pretrained=False, return_features_only=True).eval()
wav = torch.randn(128, 80000, device="cuda") * 0.1
names = model.register_hooks_for_layers(["all"])
def run():
    with torch.no_grad():
        return model.extract_embeddings(wav, aggregation="mean")
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): e = run()
b.record(); torch.cuda.synchronize()
print(f"{len(names)} hooks, 128 x 5 s: {a.elapsed_time(b) / 5:.2f} ms per batch, embedding {tuple(e.shape)}")
PY
cat gpurun_out/r2_hooks44.log | tail -3
echo done
