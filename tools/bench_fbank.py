"""Time the fused fbank kernel alone (CUDA events, L2 flushed between iterations)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avex_b200.fbank import KaldiFbank

def main(B=256, T=160000, iters=20):
    fb = KaldiFbank().cuda()
    x = torch.randn(B, T, device="cuda") * 0.1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fb.run(x, prescale=32768.0, norm_mean=15.41663, norm_std2=13.11164)
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fb.run(x, prescale=32768.0, norm_mean=15.41663, norm_std2=13.11164); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    F = out.shape[1]
    byts = B * (4 * T + 4 * F * 128)
    med = ts[len(ts)//2]
    print(json.dumps({"kernel": "fbank", "B": B, "T": T, "ms_median": med, "ms_min": ts[0], "GBps_median": byts / med / 1e6, "GBps_best": byts / ts[0] / 1e6}))

if __name__ == "__main__":
    main()
    main(B=64, T=960000)
