"""GPU: the fused BEATs forward behind the plugin surface, against the reference goldens (bf16 tolerances of
BASELINE.json north_star: per-layer cosine >= 0.999 and max-abs <= 2e-2) and the numpy oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import beats_encoder as OE
from oracle.weights import make_beats_weights
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _build(layers, wseed, return_features_only=True, num_classes=None, init="perturbed"):
    from avex_b200 import plugin
    from avex_b200.plugin import beats_model  # noqa: F401  (registers the "beats" class)

    spec = plugin.ModelSpec(name="beats", device="cuda", init_config=dict(encoder_layers=layers, finetuned_model=True, dropout=0.0))
    plugin.register_model(f"gpu_test_L{layers}", spec)
    kw = {}
    model = plugin.load_model(f"gpu_test_L{layers}", device="cuda", return_features_only=return_features_only, **kw).eval()
    W = make_beats_weights(OE.BeatsDims(layers=layers), seed=wseed, init=init)
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=False)
    assert not unexpected
    assert all(k.startswith(("backbone.fbank.", "backbone.predictor.")) for k in missing), missing
    return model, W


def _cmp(name, got, ref, atol=2e-2, cos_min=0.999):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    err = np.abs(got - ref).max()
    cos = float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-30))
    print(f"{name}: max_abs {err:.4e} cos {cos:.6f} (ref absmax {np.abs(ref).max():.3f})")
    assert np.isfinite(got).all()
    assert cos >= cos_min, (name, cos)
    assert err <= atol, (name, err)


@pytest.mark.parametrize("cname", list(cases.beats_cases()))
def test_against_reference_golden(cname):
    case = cases.beats_cases()[cname]
    g = np.load(os.path.join(G, f"beats_{cname}.npz"))
    model, W = _build(case["layers"], case["wseed"], init=case.get("init", "perturbed"))
    wav = torch.from_numpy(case["wav"]).cuda()
    mask = torch.from_numpy(case["mask"]).cuda() if "mask" in case else None
    with torch.no_grad():
        feats = model(wav, mask)
    _cmp(f"{cname}/final", feats.cpu().numpy(), g["final"])
    names = model.register_hooks_for_layers(["all"])
    assert names[0] == "backbone.post_extract_proj" and len(names) == case["layers"] + 1
    hooks = model.extract_embeddings(wav, padding_mask=mask, aggregation="none")
    assert len(hooks) == case["layers"] + 1
    for li in case["keep_hooks"]:
        _cmp(f"{cname}/hook{li}", hooks[li].cpu().numpy(), g[f"hook{li}"])
    pooled = model.extract_embeddings(wav, padding_mask=mask, aggregation="mean")
    _cmp(f"{cname}/pooled_hooks_mean", pooled.cpu().numpy(), g["pooled_hooks_mean"], atol=1e-2, cos_min=0.9999)
    # hooks on a subset, by index, still work after the "all" registration
    model.register_hooks_for_layers([0, -1])
    two = model.extract_embeddings({"raw_wav": wav, "padding_mask": mask}, aggregation="mean")
    assert two.shape == (wav.shape[0], 2 * 768)


@pytest.mark.parametrize("cname", list(cases.beats_cases()))
def test_fp32_mode_against_reference_golden(cname):
    """north_star fp32 mode: per-layer cosine >= 0.999 and max-abs <= 1e-3 against the reference run in fp32."""
    case = cases.beats_cases()[cname]
    g = np.load(os.path.join(G, f"beats_{cname}.npz"))
    model, W = _build(case["layers"], case["wseed"], init=case.get("init", "perturbed"))
    model.backbone.precision = "fp32"
    wav = torch.from_numpy(case["wav"]).cuda()
    mask = torch.from_numpy(case["mask"]).cuda() if "mask" in case else None
    with torch.no_grad():
        feats = model(wav, mask)
    _cmp(f"fp32 {cname}/final", feats.cpu().numpy(), g["final"], atol=1e-3, cos_min=0.99999)
    model.register_hooks_for_layers(["all"])
    hooks = model.extract_embeddings(wav, padding_mask=mask, aggregation="none")
    for li in case["keep_hooks"]:
        _cmp(f"fp32 {cname}/hook{li}", hooks[li].cpu().numpy(), g[f"hook{li}"], atol=1e-3, cos_min=0.99999)
    # switching back re-packs the bf16 engine
    model.backbone.precision = "bf16"
    model.deregister_all_hooks()
    with torch.no_grad():
        feats16 = model(wav, mask)
    _cmp(f"bf16 again {cname}/final", feats16.cpu().numpy(), g["final"])


def test_pooled_hooks_are_formed_on_device():
    """`extract_embeddings(aggregation="mean")` gets its per-layer [B,768] means from the fc2 epilogues (nothing of
    size [B,N,768] is written); they must equal the means of the materialised hooks, also under a padding mask (the reference
    pools hooks over ALL tokens) -- and a foreign forward hook on a hooked module switches the shortcut off."""
    case = cases.beats_cases()["L2_2x2s_mask"]
    model, _ = _build(2, case["wseed"])
    wav, mask = torch.from_numpy(case["wav"]).cuda(), torch.from_numpy(case["mask"]).cuda()
    model.register_hooks_for_layers(["all"])
    for m in (mask, None):
        full = model.extract_embeddings(wav, padding_mask=m, aggregation="none")
        want = torch.cat([h.mean(dim=1) for h in full], dim=1)
        before = torch.cuda.max_memory_allocated()
        got = model.extract_embeddings(wav, padding_mask=m, aggregation="mean")
        assert got.shape == (2, 3 * 768)
        assert (got - want).abs().max().item() <= 2e-5, (got - want).abs().max().item()
        # no [B,N,768] hook tensors were allocated (3 layers x 2 x 96 x 768 fp32 = 1.7 MB would show; the pooled outputs are 18 KB)
        assert torch.cuda.max_memory_allocated() - before < 256 * 1024
    res = model.backbone.run(wav, None, want_features=False, hook_layers=[0, 1, 2], hook_pool=True)
    assert all(t.shape == (2, 768) for t in res["hooks"].values())
    seen = []
    h = model.backbone.encoder.layers[1].fc2.register_forward_hook(lambda mod, inp, out: seen.append(tuple(out.shape)))
    got2 = model.extract_embeddings(wav, aggregation="mean")
    h.remove()
    assert seen == [(96, 2, 768)]  # the user's hook saw the reference's (T, B, C) tensor, not a pooled one
    assert (got2 - got).abs().max().item() <= 2e-5


@pytest.mark.parametrize("seconds,precision", [(0.5, "bf16"), (1.0, "fp32"), (0.5, "fp32")])
def test_pooled_hooks_fallback_paths(seconds, precision):
    """The epilogue pooling needs >= 32 tokens per clip and the bf16 engine; shorter clips (0.5 s -> N = 24) and the fp32 mode go
    through the materialise-then-pool path -- same numbers either way."""
    model, _ = _build(2, 7)
    model.backbone.precision = precision
    g = torch.Generator(device="cuda").manual_seed(int(seconds * 10))
    wav = torch.randn(3, int(16000 * seconds), device="cuda", generator=g) * 0.1
    model.register_hooks_for_layers(["all"])
    full = model.extract_embeddings(wav, aggregation="none")
    want = torch.cat([h.mean(dim=1) for h in full], dim=1)
    got = model.extract_embeddings(wav, aggregation="mean")
    assert got.shape == (3, 3 * 768) and torch.isfinite(got).all()
    assert (got - want).abs().max().item() <= 2e-5
    res = model.backbone.run(wav, None, want_features=True, want_pooled=True)
    assert (res["pooled"] - res["features"].mean(dim=1)).abs().max().item() <= 2e-5


def test_classifier_mode_masked_mean_pool():
    case = cases.beats_cases()["L2_2x2s_mask"]
    from avex_b200 import plugin

    model, W = _build(2, case["wseed"])
    spec = plugin.get_model_spec("gpu_test_L2")
    clf = plugin.build_model_from_spec(spec, "cuda", num_classes=10).to("cuda").eval()
    clf.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=False)
    wav = torch.from_numpy(case["wav"]).cuda()
    mask = torch.from_numpy(case["mask"]).cuda()
    with torch.no_grad():
        logits = clf(wav, mask)
        feats = model(wav, mask)
    assert logits.shape == (2, 10) and torch.isfinite(logits).all()
    orc = OE.beats_forward(W, case["wav"], case["mask"], OE.BeatsDims(layers=2))
    pooled_ref = OE.mean_pool(orc["x"], orc["key_pad"])
    want = pooled_ref @ clf.classifier.weight.detach().cpu().numpy().T + clf.classifier.bias.detach().cpu().numpy()
    _cmp("classifier logits", logits.cpu().numpy(), want, atol=1e-2, cos_min=0.9999)
    # all-False mask == no mask, bit-identical (verified property of the reference)
    none = model(wav[:1], None)
    allf = model(wav[:1], torch.zeros_like(mask[:1]))
    assert torch.equal(none, allf)
    assert feats.shape == (2, 96, 768)


def test_batch_independence_and_full_size():
    """BASELINE config #2 shape (256 x 10 s, N=496): a clip's embedding does not depend on its batch neighbours."""
    model, W = _build(12, 3)
    g = torch.Generator(device="cuda").manual_seed(1234)
    wav = torch.randn(64, 160000, device="cuda", generator=g) * 0.1
    with torch.no_grad():
        full = model(wav)
        one = model(wav[17:18])
    assert full.shape == (64, 496, 768) and torch.isfinite(full).all()
    assert torch.equal(full[17], one[0])
    # oracle on one clip of the big batch (CPU, ~10 s)
    orc = OE.beats_forward(W, wav[17:18].cpu().numpy(), None, OE.BeatsDims(layers=12))
    _cmp("10s clip vs oracle", one.cpu().numpy(), orc["x"])


def test_60s_unmasked_against_reference_golden():
    """BASELINE config #5 shape against the REFERENCE (not the kernel itself): one unmasked 60 s clip, N = 2992, bias offsets
    up to +-2991 -- the log-bucket region and the max_distance saturation go through `relative_bias_vector` and the kernel's
    in-tile Toeplitz indexing.  The fixture keeps every 8th token."""
    case = cases.beats_long_case()
    g = np.load(os.path.join(G, "beats_L2_1x60s.npz"))
    model, W = _build(case["layers"], case["wseed"])
    wav = torch.from_numpy(case["wav"]).cuda()
    st = case["stride"]
    with torch.no_grad():
        feats = model(wav)
    assert feats.shape == (1, 2992, 768)
    _cmp("60s/final", feats[:, ::st].cpu().numpy(), g["final"])
    _cmp("60s/final_pooled", feats.mean(dim=1).cpu().numpy(), g["final_pooled"], atol=5e-3, cos_min=0.9999)
    model.register_hooks_for_layers(["all"])
    hooks = model.extract_embeddings(wav, aggregation="none")
    for li in case["keep_hooks"]:
        _cmp(f"60s/hook{li}", hooks[li][:, ::st].cpu().numpy(), g[f"hook{li}"])
    pooled = model.extract_embeddings(wav, aggregation="mean")
    _cmp("60s/pooled_hooks_mean", pooled.cpu().numpy(), g["pooled_hooks_mean"], atol=1e-2, cos_min=0.9999)
    model.backbone.precision = "fp32"
    with torch.no_grad():
        f32 = model(wav)
    _cmp("fp32 60s/final", f32[:, ::st].cpu().numpy(), g["final"], atol=1e-3, cos_min=0.99999)


def test_predictor_logits_against_reference_golden():
    """The AudioSet predictor branch of fine-tuned checkpoints (beats.py:369-380) with and without a padding mask."""
    from oracle.weights import make_predictor_weights

    case = cases.predictor_case()
    g = np.load(os.path.join(G, "beats_predictor.npz"))
    model, W = _build(case["layers"], case["wseed"])
    P = make_predictor_weights(case["pseed"])
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=False)
    assert not unexpected and model.backbone.predictor is not None
    wav, mask = torch.from_numpy(case["wav"]).cuda(), torch.from_numpy(case["mask"]).cuda()
    with torch.no_grad():
        lg_m, km = model.backbone.extract_features(wav, mask, feature_only=False)
        lg_n, _ = model.backbone.extract_features(wav, None, feature_only=False)
    assert km is not None and km.any() and lg_m.shape == (2, 527)
    _cmp("predictor logits (mask)", lg_m.cpu().numpy(), g["logits_mask"], atol=1e-2, cos_min=0.9999)
    _cmp("predictor logits (no mask)", lg_n.cpu().numpy(), g["logits_nomask"], atol=1e-2, cos_min=0.9999)


def test_bench_shape_256x10s():
    """The bench configuration itself (256 x 10 s, M = 126 976 token rows, 496 CTA-pair tiles per GEMM): every clip of the big
    batch must equal the same clip run alone (bit-identical: tiles never mix clips' rows in any reduction), and the fused
    mean-pool must equal the mean of the features."""
    model, W = _build(12, 3)
    g = torch.Generator(device="cuda").manual_seed(4321)
    wav = torch.randn(256, 160000, device="cuda", generator=g) * 0.1
    with torch.no_grad():
        res = model.backbone.run(wav, None, want_features=True, want_pooled=True)
        full, pooled = res["features"], res["pooled"]
        assert full.shape == (256, 496, 768) and torch.isfinite(full).all()
        for i in (0, 100, 255):
            one = model(wav[i : i + 1])
            assert torch.equal(full[i], one[0]), i
        sub = model(wav[128:192])
        assert torch.equal(full[128:192], sub)
    assert torch.allclose(pooled, full.mean(dim=1), atol=2e-5, rtol=1e-5)
    orc = OE.beats_forward(W, wav[255:256].cpu().numpy(), None, OE.BeatsDims(layers=12))
    _cmp("clip 255 of 256 vs oracle", full[255:256].cpu().numpy(), orc["x"])


def test_long_clip_equals_short_clip_under_padding_mask():
    """BASELINE config #5 shape (60 s clips, N = 2992 tokens), checked through a size-independent property instead of a
    60 s oracle run: a 10 s clip followed by 50 s of padding (samples >= 992 frames * 160 masked, so tokens >= 496 are
    padded) must give, on its first 496 tokens, the features of the 10 s clip alone -- padded tokens are zeroed before
    the pos-conv (== its zero padding) and excluded as attention keys, and fbank / LayerNorm are per frame / per token."""
    model, _ = _build(12, 3)
    g = torch.Generator(device="cuda").manual_seed(77)
    short = torch.randn(2, 160000, device="cuda", generator=g) * 0.1
    long = torch.zeros(2, 960000, device="cuda")
    long[:, :160000] = short
    long[:, 160000:] = torch.randn(2, 800000, device="cuda", generator=g) * 0.1  # content under the mask must not matter
    mask = torch.zeros(2, 960000, dtype=torch.bool, device="cuda")
    mask[:, 992 * 160 :] = True
    with torch.no_grad():
        f_short = model(short)
        f_long = model(long, mask)
    assert f_short.shape == (2, 496, 768) and f_long.shape == (2, 2992, 768)
    assert torch.isfinite(f_long).all()
    _cmp("60 s masked vs 10 s", f_long[:, :496].cpu().numpy(), f_short.cpu().numpy())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """the dynamic-shared-memory opt-in is a per-DEVICE function attribute -- a model on cuda:1 after one on cuda:0 in
    the same process must launch (the flags were once per process)."""
    model0, _ = _build(2, 1)
    wav = torch.randn(2, 16000, generator=torch.Generator().manual_seed(3)) * 0.1
    with torch.no_grad():
        f0 = model0(wav.cuda(0))
    import copy

    model1 = copy.deepcopy(model0).to("cuda:1")
    with torch.cuda.device(1), torch.no_grad():
        f1 = model1(wav.cuda(1))
    assert torch.equal(f0.cpu(), f1.cpu())
