"""Stress the memory-bound 1x1-conv kernel: many tiles per CTA, repeated, every result checked.  python tools/stress_pointwise.py [iters]"""
import os, sys, time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from avex_b200 import _lib

lib = _lib.load()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(516096, 240, 40, 1, 0), (300000, 144, 24, 1, 0), (400000, 96, 16, 1, 0), (131072, 480, 80, 1, 0), (32768, 1152, 192, 1, 0),
          (516096, 40, 240, 0, 1008), (400000, 24, 144, 0, 4000), (131072, 112, 672, 0, 256), (400000, 16, 32, 0, 16000)]
for M, N, K, silu, hw in shapes:
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).to(torch.float16)
    scale = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    shift = 0.3 * torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    if hw:
        clips = (M + hw - 1) // hw
        se = torch.rand(clips, K, device="cuda", generator=g)
        rows = torch.arange(M, device="cuda") // hw
        As = (A.float() * se[rows]).to(torch.float16).float()
    else:
        As = A.float()
    ref = (As @ W.float().T) * scale + shift
    if silu:
        ref = torch.nn.functional.silu(ref)
    tol = 3e-3 * max(1.0, ref.abs().max().item())
    t0 = time.time()
    for it in range(iters):
        out.fill_(7.0)
        if hw:
            rc = lib.avexk_conv1x1_se_f16(A.data_ptr(), se.data_ptr(), hw, W.data_ptr(), M, N, K, scale.data_ptr(), shift.data_ptr(), None, out.data_ptr(), st)
        else:
            rc = lib.avexk_conv1x1_f16(A.data_ptr(), W.data_ptr(), M, N, K, scale.data_ptr(), shift.data_ptr(), silu, None, None, out.data_ptr(), 1, st)
        assert rc == 0, lib.avexk_last_error().decode()
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        assert err <= tol, f"M={M} N={N} K={K} iter {it}: err {err} > {tol}"
    print(f"M={M} N={N} K={K} silu={silu} hw={hw}: {iters} iterations ok, {(time.time() - t0) / iters * 1e3:.2f} ms each (with check)", flush=True)
print("stress ok")
