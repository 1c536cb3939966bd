"""Where does the end-to-end step (pinned host input -> H2D -> forward -> D2H) lose time against the device-only step?
Times, with CUDA events: the forward alone, the H2D copy alone, and the overlapped pipeline under a few copy strategies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avex_b200 import plugin
from avex_b200.plugin import beats_model  # noqa: F401

dev = torch.device("cuda", 0)
init_config = dict(encoder_layers=12, encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_attention_heads=12, deep_norm=True,
                   dropout=0.0, attention_dropout=0.0, finetuned_model=False, layer_wise_gradient_decay_ratio=0.6)
plugin.register_model("probe_beats", plugin.ModelSpec(name="beats", device="cuda", init_config=init_config))
torch.manual_seed(0)
bk = plugin.load_model("probe_beats", device="cuda", return_features_only=True).eval().backbone
B, T = 256, 160000
wav = torch.randn(B, T, device=dev) * 0.1
host = [wav.cpu().pin_memory() for _ in range(2)]
dev_in = [torch.empty(B, T, device=dev) for _ in range(2)]
pooled_host = torch.empty(B, 768).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)


def timeit(fn, n=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print("forward only            %.2f ms" % timeit(lambda i: bk.run(wav, None, want_features=False, want_pooled=True)))
print("H2D only (164 MB)       %.2f ms" % timeit(lambda i: dev_in[i & 1].copy_(host[i & 1], non_blocking=True)))


def pipeline(chunks):
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_used = [torch.cuda.Event(), torch.cuda.Event()]

    def enqueue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[slot])
            n = B // chunks
            for c in range(chunks):
                dev_in[slot][c * n:(c + 1) * n].copy_(host[slot][c * n:(c + 1) * n], non_blocking=True)
            ev_ready[slot].record(copy_stream)

    enqueue_copy(0)

    def step(i):
        cur = i & 1
        main = torch.cuda.current_stream()
        enqueue_copy(cur ^ 1)  # BEFORE this batch's kernels are enqueued
        main.wait_event(ev_ready[cur])
        res = bk.run(dev_in[cur], None, want_features=False, want_pooled=True)
        ev_used[cur].record(main)
        pooled_host.copy_(res["pooled"], non_blocking=True)

    return step


for chunks in (1, 8):
    print("pipeline, copy enqueued first, %d chunk(s)  %.2f ms" % (chunks, timeit(pipeline(chunks))))
