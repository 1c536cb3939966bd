"""Time avexk_gemm_bf16 on the BEATs shapes (CUDA events; inputs exceed L2)."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avex_b200 import _lib

def main():
    lib = _lib.load()
    M = int(os.environ.get("GEMM_M", 126976))
    iters = int(os.environ.get("GEMM_ITERS", 10))
    shapes = [("qkv", 2304, 768, 0, False, True), ("out_proj", 768, 768, 0, True, False),
              ("fc1_gelu", 3072, 768, 1, False, True), ("fc2", 768, 3072, 0, True, False)]
    st = torch.cuda.current_stream().cuda_stream
    for name, N, K, gelu, res, obf in shapes:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda") if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
        def run():
            rc = lib.avexk_gemm_bf16(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), gelu, None,
                                     R.data_ptr() if res else None, 2.2133638, out.data_ptr(), N, int(obf), st)
            assert rc == 0, lib.avexk_last_error()
        for _ in range(2): run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "ms": round(ms, 4), "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}), flush=True)
        del A, W, R, out

if __name__ == "__main__":
    main()
