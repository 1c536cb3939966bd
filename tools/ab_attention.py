"""A/B: time the attention kernel(s) at the config #2 shape and check against the fp32 torch formulation (GPU only)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from avex_b200 import _lib
from oracle import relpos as OR

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
names = ["avexk_attention_gated"] + (["avexk_attention_gated_mma"] if hasattr(raw, "avexk_attention_gated_mma") else [])


def ref_attn(qkv, B, N, H, gw, gb, ga, bias_vec, key_pad):
    d = 64
    q, k, v = [t.reshape(B, N, H, d).permute(0, 2, 1, 3).float() for t in qkv.split(H * d, dim=1)]
    gate = torch.sigmoid(q @ gw.T + gb)
    g1 = gate[..., 0:1] * (gate[..., 1:2] * ga.view(1, H, 1, 1) - 1.0) + 2.0
    idx = (torch.arange(N, device="cuda")[None, :] - torch.arange(N, device="cuda")[:, None]) + N - 1
    s = q @ k.transpose(-1, -2) * 0.125 + g1 * bias_vec[:, idx][None]
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :].bool(), float("-inf"))
    return (torch.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * N, H * d)


def run(B, N, H, pad, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(N)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda", generator=g).to(torch.bfloat16)
    gw = torch.randn(2, 64, device="cuda", generator=g) * 0.2
    gb = torch.randn(2, device="cuda", generator=g) * 0.2
    ga = 1.0 + 0.2 * torch.randn(H, device="cuda", generator=g)
    table = torch.randn(320, H, generator=torch.Generator().manual_seed(1))
    bias_vec = torch.from_numpy(OR.bias_vector(table.numpy(), N)).cuda()
    key_pad = None
    if pad:
        key_pad = torch.zeros(B, N, dtype=torch.uint8, device="cuda")
        key_pad[-1, N // 2 + 3:] = 1
        if B > 8:  # timing shape: suffix padding of 0 / 60 / 120 / 180 tokens, as a padded evaluation batch has
            for i in range(B):
                if i % 4:
                    key_pad[i, N - (i % 4) * 60:] = 1
    st = torch.cuda.current_stream().cuda_stream
    refB = min(B, 4)
    ref = ref_attn(qkv[: refB * N], refB, N, H, gw, gb, ga, bias_vec, key_pad[:refB] if pad and refB == B else None) if not (pad and refB != B) else None
    for name in names:
        fn = getattr(raw, name)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] * 1 + [C.c_int] * 3 + [C.c_void_p] * 7
        out = torch.zeros(B * N, H * 64, device="cuda", dtype=torch.bfloat16)
        rc = fn(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(), key_pad.data_ptr() if pad else None, out.data_ptr(), st)
        torch.cuda.synchronize()
        msg = f"{name:28s} B={B} N={N} H={H} pad={pad} rc={rc}"
        if rc != 0:
            print(msg, lib.avexk_last_error()); continue
        if ref is not None:
            o = out[: refB * N].float()
            err = (o - ref).abs().max().item()
            cos = torch.nn.functional.cosine_similarity(o.flatten(), ref.flatten(), dim=0).item()
            msg += f" max_err={err:.3e} cos={cos:.6f} finite={bool(torch.isfinite(out.float()).all())}"
        if time_it:
            for _ in range(3):
                fn(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(), None if not pad else key_pad.data_ptr(), out.data_ptr(), st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(qkv.data_ptr(), B, N, H, gw.data_ptr(), gb.data_ptr(), ga.data_ptr(), bias_vec.data_ptr(), None if not pad else key_pad.data_ptr(), out.data_ptr(), st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            msg += f"  {ms:.3f} ms  {4.0*B*H*N*N*64/ms/1e9:.0f} TFLOP/s"
        print(msg, flush=True)


if __name__ == "__main__":
    if "--profile" in sys.argv:
        names[:] = names[:1]
        run(256, 496, 12, False)
        sys.exit(0)
    for (B, N, H, pad) in [(1, 128, 1, False), (2, 48, 12, False), (1, 248, 12, False), (3, 96, 4, True), (2, 496, 2, False), (1, 700, 1, True), (2, 256, 3, False), (1, 2992, 2, False)]:
        run(B, N, H, pad)
    run(256, 496, 12, False, time_it=True)
    run(256, 496, 12, True, time_it=True)
    run(64, 2992, 12, False, time_it=True)
