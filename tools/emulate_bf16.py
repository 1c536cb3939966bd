"""No-GPU precision study: the numpy oracle with bf16 (or fp16) rounding injected at exactly the sites where the
CUDA path rounds (GEMM operands, stored qkv / attention output / GELU(fc1), P before PV).  Used to decide where
reduced precision is affordable and to check that GPU-measured errors are the inherent
operand-rounding noise, not bugs.  Test infrastructure only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import beats_encoder as OE, kaldi_fbank as OF, relpos
from oracle.weights import make_beats_weights
from tests.golden import cases

def bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)

def fp16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)

def split3(a, q):  # hi/lo split GEMM helper: returns list of (a_part) such that sum ~ a
    hi = q(a); lo = q(a - hi)
    return hi, lo

def forward(W, wav, mask, dims, q=bf16, front_split=False, split_all=False):
    dt = np.float32
    g = lambda n: W[n].astype(dt)
    def mm(a, w):  # a @ w.T with operand rounding
        if split_all:
            ah, al = split3(a, q); wh, wl = split3(w, q)
            return ah @ wh.T + al @ wh.T + ah @ wl.T
        return q(a) @ q(w).T
    def mm_front(a, w):
        if front_split:
            ah, al = split3(a, q); wh, wl = split3(w, q)
            return ah @ wh.T + al @ wh.T + ah @ wl.T
        return mm(a, w)
    fb = OF.beats_preprocess(wav, dims.fbank_mean, dims.fbank_std)
    key_pad = OE.token_padding_mask(mask, dims) if mask is not None else None
    a = OE.patchify(fb)
    x = mm_front(a, g("backbone.patch_embedding.weight").reshape(512, -1))
    x = OE.layer_norm(x, g("backbone.layer_norm.weight"), g("backbone.layer_norm.bias"))
    x = mm_front(x, g("backbone.post_extract_proj.weight")) + g("backbone.post_extract_proj.bias")
    if key_pad is not None: x = np.where(key_pad[:, :, None], dt(0), x)
    out = {"hook0": x.copy(), "fc2": []}
    B, N, C = x.shape
    Wq = dict(W)
    # pos conv with rounded operands
    Wc = OE.pos_conv_weight(W, dt)
    Wq2 = dict(W)
    xq = q(x) if not split_all else x
    K, G = 128, 16; cg = 48
    xp = np.zeros((B, N + K, C), dt); xp[:, 64:64 + N] = xq
    conv = np.empty((B, N, C), dt)
    wq = q(Wc) if not split_all else Wc
    for gi in range(G):
        win = np.lib.stride_tricks.sliding_window_view(xp[:, :, gi*cg:(gi+1)*cg], K, axis=1)[:, :N]
        conv[:, :, gi*cg:(gi+1)*cg] = (win.reshape(B*N, cg*K) @ wq[gi*cg:(gi+1)*cg].reshape(cg, cg*K).T).reshape(B, N, cg)
    x = x + OE.gelu(conv + g("backbone.encoder.pos_conv.0.bias"))
    x = OE.layer_norm(x, g("backbone.encoder.layer_norm.weight"), g("backbone.encoder.layer_norm.bias"))
    table = g("backbone.encoder.layers.0.self_attn.relative_attention_bias.weight")
    bias_vec = relpos.bias_vector(table, N)
    idx = (np.arange(N)[None, :] - np.arange(N)[:, None]) + N - 1
    pos_bias = bias_vec[:, idx]
    alpha = dt(dims.alpha); H, d = 12, 64
    for li in range(dims.layers):
        p = f"backbone.encoder.layers.{li}"; sa = p + ".self_attn"
        qq = q(mm(x, g(sa + ".q_proj.weight")) + g(sa + ".q_proj.bias"))
        kk = q(mm(x, g(sa + ".k_proj.weight")) + g(sa + ".k_proj.bias"))
        vv = q(mm(x, g(sa + ".v_proj.weight")) + g(sa + ".v_proj.bias"))
        qh, kh, vh = [t.reshape(B, N, H, d).transpose(0, 2, 1, 3) for t in (qq, kk, vv)]
        gl = (qh @ g(sa + ".grep_linear.weight").T + g(sa + ".grep_linear.bias")).reshape(B, H, N, 2, 4).sum(-1)
        gate = 1 / (1 + np.exp(-gl))
        g1 = gate[..., 0:1] * (gate[..., 1:2] * g(sa + ".grep_a").reshape(1, H, 1, 1) - 1) + 2
        s = (qh @ kh.transpose(0, 1, 3, 2)) * dt(0.125) + g1 * pos_bias[None]
        if key_pad is not None: s = np.where(key_pad[:, None, None, :], dt(-np.inf), s)
        s = s - s.max(-1, keepdims=True); pr = np.exp(s); l = pr.sum(-1, keepdims=True)
        o = (q(pr) @ vh) / l
        att = q(o.transpose(0, 2, 1, 3).reshape(B, N, C))
        ao = mm(att, g(sa + ".out_proj.weight")) + g(sa + ".out_proj.bias")
        x = OE.layer_norm(x * alpha + ao, g(p + ".self_attn_layer_norm.weight"), g(p + ".self_attn_layer_norm.bias"))
        h = q(OE.gelu(mm(x, g(p + ".fc1.weight")) + g(p + ".fc1.bias")))
        f2 = mm(h, g(p + ".fc2.weight")) + g(p + ".fc2.bias")
        out["fc2"].append(f2)
        x = OE.layer_norm(x * alpha + f2, g(p + ".final_layer_norm.weight"), g(p + ".final_layer_norm.bias"))
    out["x"] = x
    return out

def report(tag, got, ref):
    e = np.abs(got.astype(np.float64) - ref).max()
    c = (got.astype(np.float64) * ref).sum() / np.linalg.norm(got) / np.linalg.norm(ref)
    return f"{tag}: max {e:.4f} cos {c:.6f}"

if __name__ == "__main__":
    G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    which = sys.argv[1:] or list(cases.beats_cases())
    for cname in which:
        case = cases.beats_cases()[cname]
        gold = np.load(os.path.join(G, f"beats_{cname}.npz"))
        dims = OE.BeatsDims(layers=case["layers"])
        W = make_beats_weights(dims, seed=case["wseed"], init=case.get("init", "perturbed"))
        for tag, kw in [("bf16", {}), ("bf16+front3", dict(front_split=True)), ("fp16", dict(q=fp16)), ("bf16 split3 all", dict(split_all=True))]:
            o = forward(W, case["wav"], case.get("mask"), dims, **kw)
            hooks = [o["hook0"]] + o["fc2"]
            line = [report("final", o["x"], gold["final"])] + [report(f"h{li}", hooks[li], gold[f"hook{li}"]) for li in case["keep_hooks"]]
            print(f"{cname:20s} {tag:16s} " + " | ".join(line), "| ref absmax %.2f" % np.abs(gold["final"]).max(), flush=True)
