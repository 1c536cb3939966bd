"""Cost of materialising per-layer hook outputs: forward without hooks vs extract_embeddings over all 13 layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avex_b200 import plugin
from avex_b200.plugin import beats_model  # noqa: F401

init_config = dict(encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12, deep_norm=True,
                   dropout=0.0, attention_dropout=0.0, finetuned_model=False, layer_wise_gradient_decay_ratio=0.6)
plugin.register_model("probe_beats", plugin.ModelSpec(name="beats", device="cuda", init_config=init_config))
torch.manual_seed(0)
model = plugin.load_model("probe_beats", device="cuda", return_features_only=True).eval()
B = int(os.environ.get("PROBE_B", 128))
wav = torch.randn(B, 160000, device="cuda") * 0.1


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


with torch.no_grad():
    print("forward, no hooks            %.2f ms" % timeit(lambda: model(wav)))
    model.register_hooks_for_layers(["last_layer"])
    print("extract last layer (mean)    %.2f ms" % timeit(lambda: model.extract_embeddings(wav, aggregation="mean")))
    model.deregister_all_hooks()
    model.register_hooks_for_layers(["all"])
    print("extract all 13 layers (mean) %.2f ms" % timeit(lambda: model.extract_embeddings(wav, aggregation="mean")))
    print("extract all 13 layers (none) %.2f ms" % timeit(lambda: model.extract_embeddings(wav, aggregation="none")))
