"""Time avexk_gemm_bf16 / avexk_gemm_bf16_ln on the BEATs shapes (CUDA events; inputs exceed L2).

    GEMM_M=126976 GEMM_ITERS=10 GEMM_PAIR=1,0 python tools/bench_gemm.py
"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avex_b200 import _lib


def main():
    lib = _lib.load()
    M = int(os.environ.get("GEMM_M", 126976))
    iters = int(os.environ.get("GEMM_ITERS", 10))
    only = os.environ.get("GEMM_ONLY")
    # name, N, K, gelu, residual, bf16 out, fused LN
    shapes = [("qkv", 2304, 768, 0, False, True, False), ("out_proj", 768, 768, 0, True, False, False),
              ("out_proj+ln", 768, 768, 0, True, False, True), ("fc1_gelu", 3072, 768, 1, False, True, False),
              ("fc2", 768, 3072, 0, True, False, False), ("fc2+ln", 768, 3072, 0, True, False, True)]
    st = torch.cuda.current_stream().cuda_stream
    for pair in [int(p) for p in os.environ.get("GEMM_PAIR", "1,0").split(",")]:
        lib.avexk_gemm_config(pair)
        for name, N, K, gelu, res, obf, ln in shapes:
            if only and name not in only.split(","):
                continue
            A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
            bias = torch.randn(N, device="cuda")
            R = torch.randn(M, N, device="cuda") if res else None
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
            if ln:
                gam, bet = torch.randn(N, device="cuda"), torch.randn(N, device="cuda")
                xb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
                nb = lib.avexk_gemm_ln_scratch_bytes(M)
                scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")

            def run():
                if ln:
                    rc = lib.avexk_gemm_bf16_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), None, R.data_ptr(), 2.2133638,
                                                gam.data_ptr(), bet.data_ptr(), 1e-5, None if os.environ.get("GEMM_NOF32") else out.data_ptr(),
                                                None if os.environ.get("GEMM_NOXB") else xb.data_ptr(), scratch.data_ptr(), nb, st)
                else:
                    rc = lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), gelu, None,
                                             R.data_ptr() if res else None, 2.2133638, out.data_ptr(), N, int(obf), st)
                assert rc == 0, lib.avexk_last_error()

            for _ in range(2):
                run()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                run()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            print(json.dumps({"gemm": name, "pair": pair, "M": M, "N": N, "K": K, "ms": round(ms, 4),
                              "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}), flush=True)
            del A, W, R, out


if __name__ == "__main__":
    main()
