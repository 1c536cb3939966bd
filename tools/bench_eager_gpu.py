"""The incumbent GPU implementation: the reference's op mix in stock torch eager on the B200 itself -- fp32 and
under autocast(bf16) -- next to the fused path, on the same clips.  oracle/beats_torch.py is the same restatement bench.py's
CPU baseline times (materialised [B,H,N,N] gate*bias mask + SDPA, like backbone.py:544-568); here its tensors live on cuda:0."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import beats_encoder as OE
from oracle import beats_torch as OT
from oracle.weights import make_beats_weights

B = int(os.environ.get("EAGER_B", 32))
dims = OE.BeatsDims(layers=12)
W = make_beats_weights(dims, seed=0, init="reference")
Wt = {k: v.cuda() for k, v in OT.to_torch(W).items()}
wav = torch.randn(B, 160000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1234)) * 0.1


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run():
    return OT.beats_forward(Wt, wav, None, dims)["x"].mean(dim=1)


res = {}
ms = timeit(run)
res["eager_fp32"] = {"ms_per_batch": ms, "audio_h_per_s": B * 10 / 3600 / (ms / 1e3)}
with torch.autocast("cuda", dtype=torch.bfloat16):
    ms = timeit(run)
res["eager_autocast_bf16"] = {"ms_per_batch": ms, "audio_h_per_s": B * 10 / 3600 / (ms / 1e3)}

from avex_b200 import plugin
from avex_b200.plugin import beats_model  # noqa: F401

plugin.register_model("eager_cmp", plugin.ModelSpec(name="beats", device="cuda", init_config=dict(encoder_layers=12)))
model = plugin.load_model("eager_cmp", device="cuda", return_features_only=True).eval()
model.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=False)
with torch.no_grad():
    ms = timeit(lambda: model.backbone.run(wav, None, want_features=False, want_pooled=True))
    res["avex_b200_bf16"] = {"ms_per_batch": ms, "audio_h_per_s": B * 10 / 3600 / (ms / 1e3)}
    ours = model.backbone.run(wav, None, want_features=True)["features"].mean(dim=1)
    ref = run()
res["pooled_max_abs_vs_eager_fp32"] = float((ours - ref).abs().max())
res["batch"] = B
print(json.dumps(res))
