"""Where the EfficientNet end-to-end step loses time against the device-only step (same model / inputs as bench.py --workload effnet)."""
import os, sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from avex_b200 import plugin
from avex_b200.plugin import efficientnet_model  # noqa: F401

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
spec = plugin.ModelSpec(name="efficientnet", device="cuda", efficientnet_variant="b0",
                        audio_config=dict(sample_rate=16000, n_fft=800, hop_length=160, win_length=800, window="hann", n_mels=128,
                                          representation="mel_spectrogram", normalize=True, target_length_seconds=10, window_selection="random"))
model = plugin.build_model_from_spec(spec, "cuda", pretrained=False, return_features_only=True).eval()
# (timing only: the module's own random initialisation; bench.py loads seeded weights with calibrated BatchNorm statistics)
B, T = 512, 80000
wav = torch.randn(B, T, device="cuda") * 0.1
host = [torch.empty(B, T).pin_memory() for _ in range(2)]
for h in host:
    h.copy_(wav.cpu())
feat_host = torch.empty(B, 1280).pin_memory()
dev_in = [torch.empty(B, T, device="cuda") for _ in range(2)]
cs = torch.cuda.Stream()


def timed(fn, n=10):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    print("forward only                    %.2f ms" % timed(lambda i: model(wav)))
    print("forward + mean                  %.2f ms" % timed(lambda i: model(wav).mean(dim=(2, 3))))
    print("forward + mean + D2H            %.2f ms" % timed(lambda i: feat_host.copy_(model(wav).mean(dim=(2, 3)), non_blocking=True)))
    print("H2D alone (164 MB)              %.2f ms" % timed(lambda i: dev_in[0].copy_(host[0], non_blocking=True)))

    def serial(i):
        dev_in[0].copy_(host[0], non_blocking=True)
        feat_host.copy_(model(dev_in[0]).mean(dim=(2, 3)), non_blocking=True)

    print("H2D then forward (one stream)   %.2f ms" % timed(serial))
    evr = [torch.cuda.Event(), torch.cuda.Event()]
    evc = [torch.cuda.Event(), torch.cuda.Event()]
    evr[0].record(); evc[1].record()

    def overlapped(i):
        cur = i & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(cs):
            cs.wait_event(evc[cur ^ 1])
            dev_in[cur ^ 1].copy_(host[cur ^ 1], non_blocking=True)
            evr[cur ^ 1].record(cs)
        main.wait_event(evr[cur])
        f = model(dev_in[cur])
        evc[cur].record(main)
        feat_host.copy_(f.mean(dim=(2, 3)), non_blocking=True)

    print("H2D on a copy stream, 2 buffers %.2f ms" % timed(overlapped))
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10):
        model(wav)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("host time to enqueue one forward %.2f ms (device %.2f ms)" % ((t1 - t0) / 10 * 1e3, (t2 - t0) / 10 * 1e3))
