"""GPU parity of the EfficientNet path (mel front end, 1x1 / depthwise conv kernels, whole feature extractor) through
the C ABI, against golden vectors produced by the reference itself (tests/golden/make_golden_effnet.py) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import effnet as OEF
from oracle import melspec as OM
from oracle.weights import make_effnet_weights

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    from avex_b200 import _lib

    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(rc, lib):
    assert rc == 0, lib.avexk_last_error().decode()


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))


# ---------------------------------------------------------------------------------------------------------------------
# mel spectrogram (audio_utils.py:106-172)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["noise_2x1s", "tones_2x1s", "noise_1x5s", "ragged_1x8123"])
def test_melspec_vs_reference(case):
    from avex_b200.melspec import MelSpectrogram

    z = np.load(os.path.join(G, "effnet_mel.npz"))
    wav, ref = z[case + "__wav"], z[case + "__mel"]
    got = MelSpectrogram().run(torch.from_numpy(wav).cuda(), normalize=True).cpu().numpy()
    assert got.shape == ref.shape
    f64 = OM.mel_spectrogram(wav)
    err_ref, err_64, ref_64 = np.abs(got - ref), np.abs(got - f64), np.abs(ref - f64)
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"melspec {case}: vs reference max {err_ref.max():.2e} rel-L2 {rel:.2e}; vs f64 max {err_64.max():.2e} (reference's own {ref_64.max():.2e})")
    assert rel <= 5e-5  # the fp32 reference itself is up to 1.8e-5 (rel-L2) from the float64 evaluation
    assert (err_ref > 1e-4 + 1e-4 * np.abs(ref)).mean() <= 1e-4
    assert err_64.max() <= max(2.0 * ref_64.max(), 1e-5)  # no further from float64 than the reference itself (x2)


def test_melspec_unnormalised_and_minmax():
    from avex_b200.melspec import MelSpectrogram

    wav = (np.random.RandomState(7).standard_normal((3, 24000)) * 0.05).astype(np.float32)
    mel = MelSpectrogram()
    y, mm = mel.run(torch.from_numpy(wav).cuda(), normalize=False, return_minmax=True)
    ref = OM.log_mel(wav)
    assert np.abs(y.cpu().numpy() - ref).max() <= 2e-3
    yn = mel.run(torch.from_numpy(wav).cuda(), normalize=True).cpu().numpy()
    assert yn.min() >= 0.0 and yn.max() <= 1.0
    for b in range(3):
        assert yn[b].min() == 0.0 and abs(yn[b].max() - 1.0) < 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K,silu,res", [(1000, 96, 16, 1, False), (777, 24, 96, 0, True), (4032, 144, 24, 1, False),
                                            (300, 40, 240, 0, False), (64 * 5, 1280, 320, 1, False), (129, 112, 672, 0, True)])
def test_conv1x1(lib, M, N, K, silu, res):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).to(torch.float16)
    scale = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    shift = 0.3 * torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).to(torch.float16) if res else None
    raw = torch.empty(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    _check(lib.avexk_conv1x1_f16(A.data_ptr(), W.data_ptr(), M, N, K, scale.data_ptr(), shift.data_ptr(), silu,
                                  R.data_ptr() if res else None, raw.data_ptr(), out.data_ptr(), 1, _stream()), lib)  # fmt: skip
    acc = A.float() @ W.float().T
    ref = acc * scale + shift
    if silu:
        ref = torch.nn.functional.silu(ref)
    if res:
        ref = ref + R.float()
    assert (raw - acc).abs().max().item() <= 1e-4 * max(1.0, acc.abs().max().item())
    assert (out.float() - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())  # fp16 output rounding


# the MBConv shapes of the 5 s bench (rows = a few 128-row tiles + a ragged tail), project convs with the SE scale fused
@pytest.mark.parametrize("rows_per_clip,clips,N,K,res", [(16000, 2, 16, 32, False), (4000, 3, 24, 96, False), (1008, 5, 40, 144, False),
                                                         (1008, 5, 40, 240, True), (256, 9, 80, 480, True), (256, 9, 112, 672, True),
                                                         (64, 37, 192, 1152, True), (64, 37, 320, 1152, False), (100, 3, 24, 144, True),
                                                         (1008, 60, 40, 240, True), (4000, 20, 24, 96, False)])  # 3-4 tiles per CTA
def test_conv1x1_se(lib, rows_per_clip, clips, N, K, res):
    g = torch.Generator(device="cuda").manual_seed(N * K)
    M = rows_per_clip * clips
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    se = torch.rand(clips, K, device="cuda", generator=g)
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).to(torch.float16)
    scale = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    shift = 0.3 * torch.randn(N, device="cuda", generator=g)
    R = torch.randn(M, N, device="cuda", generator=g).to(torch.float16) if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    A0 = A.clone()
    _check(lib.avexk_conv1x1_se_f16(A.data_ptr(), se.data_ptr(), rows_per_clip, W.data_ptr(), M, N, K, scale.data_ptr(), shift.data_ptr(),
                                     R.data_ptr() if res else None, out.data_ptr(), _stream()), lib)  # fmt: skip
    assert torch.equal(A, A0)  # the rescale happens on the staged copy
    # the kernel rounds se * A to fp16 before the tensor core, as the in-place se_apply pass did
    As = (A.float().view(clips, rows_per_clip, K) * se[:, None, :]).to(torch.float16).float().view(M, K)
    ref = (As @ W.float().T) * scale + shift
    if res:
        ref = ref + R.float()
    assert (out.float() - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K,silu", [(128 * 70 + 5, 96, 16, 1), (5000, 144, 24, 1), (3000, 240, 40, 1), (1500, 480, 80, 1),
                                        (700, 672, 112, 1), (400, 1152, 192, 1), (333, 8, 8, 0), (1, 1280, 320, 1),
                                        # several tiles per CTA: the accumulator ring (2 stages at N = 240, 5 at N = 96) wraps
                                        (148 * 128 * 3 + 77, 240, 40, 1), (148 * 128 * 6 + 5, 96, 16, 1), (148 * 128 * 2 + 1, 480, 80, 1)])
def test_conv1x1_wide_and_tiled(lib, M, N, K, silu):
    """expand-conv shapes: N up to 1152 (several N tiles, W streamed), K from one partial k-block to three."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.float16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).to(torch.float16)
    scale = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    shift = 0.3 * torch.randn(N, device="cuda", generator=g)
    out = torch.full((M + 1, N), 7.0, device="cuda", dtype=torch.float16)  # one guard row behind the output
    _check(lib.avexk_conv1x1_f16(A.data_ptr(), W.data_ptr(), M, N, K, scale.data_ptr(), shift.data_ptr(), silu, None, None,
                                  out.data_ptr(), 1, _stream()), lib)  # fmt: skip
    ref = (A.float() @ W.float().T) * scale + shift
    if silu:
        ref = torch.nn.functional.silu(ref)
    assert (out[:M].float() - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())
    assert bool((out[M] == 7.0).all())  # nothing written past row M


@pytest.mark.parametrize("B,H,W,C,k,stride", [(2, 64, 51, 32, 3, 1), (1, 64, 101, 96, 3, 2), (3, 32, 26, 144, 5, 2),
                                              (2, 8, 7, 480, 5, 1), (2, 4, 4, 1152, 3, 1), (1, 9, 13, 240, 3, 2),
                                              # the 5 s bench shapes whose tiling spans several column tiles / bands / slices
                                              (2, 64, 250, 32, 3, 1), (1, 64, 250, 96, 3, 2), (2, 32, 125, 144, 5, 2),
                                              (2, 16, 63, 240, 5, 1), (3, 8, 32, 672, 5, 2), (5, 4, 16, 1152, 5, 1),
                                              (1, 70, 300, 40, 3, 1), (1, 33, 17, 24, 5, 1)])
def test_dwconv(lib, B, H, W, C, k, stride):
    g = torch.Generator(device="cuda").manual_seed(C + k)
    x = torch.randn(B, H, W, C, device="cuda", generator=g).to(torch.float16)
    w = torch.randn(C, 1, k, k, device="cuda", generator=g) / k
    scale = 1.0 + 0.2 * torch.randn(C, device="cuda", generator=g)
    shift = 0.3 * torch.randn(C, device="cuda", generator=g)
    p = (k - 1) // 2
    Ho, Wo = (H + 2 * p - k) // stride + 1, (W + 2 * p - k) // stride + 1
    out = torch.empty(B, Ho, Wo, C, device="cuda", dtype=torch.float16)
    se = torch.empty(B, C, device="cuda")
    ws = torch.empty(2 * B * C + C * k * k, device="cuda")
    _check(lib.avexk_dwconv_nhwc(x.data_ptr(), B, H, W, C, k, stride, w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                 out.data_ptr(), se.data_ptr(), ws.data_ptr(), _stream()), lib)  # fmt: skip
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=p, groups=C)
    ref = torch.nn.functional.silu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    got = out.float().permute(0, 3, 1, 2)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() <= 3e-3 * max(1.0, ref.abs().max().item())
    ref_se = ref.sum(dim=(2, 3))
    assert (se - ref_se).abs().max().item() <= 1e-3 * max(1.0, ref_se.abs().max().item())


# ---------------------------------------------------------------------------------------------------------------------
# whole network through the plugin surface
# ---------------------------------------------------------------------------------------------------------------------
def _model(num_classes=None):
    from avex_b200 import plugin
    from avex_b200.plugin import efficientnet_model  # noqa: F401  (registers the class)

    stats = dict(np.load(os.path.join(G, "effnet_bn_stats.npz")))
    W = make_effnet_weights(seed=3, num_classes=num_classes or 0, bn_stats=stats)
    spec = plugin.ModelSpec(name="efficientnet", device="cuda", efficientnet_variant="b0",
                            audio_config=dict(sample_rate=16000, n_fft=800, hop_length=160, win_length=800, window="hann", n_mels=128,
                                              representation="mel_spectrogram", normalize=True, target_length_seconds=10, window_selection="random"))
    plugin.register_model("effnet_test", spec)
    kw = dict(pretrained=False, return_features_only=num_classes is None)
    if num_classes is not None:
        kw["num_classes"] = num_classes
    model = plugin.build_model_from_spec(spec, "cuda", **kw).eval()
    res = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=num_classes is not None)
    assert not res.unexpected_keys
    return model, W


@pytest.mark.parametrize("case", ["noise_2x1s", "tones_1x2s"])
def test_effnet_forward_and_hooks_vs_reference(case):
    z = np.load(os.path.join(G, f"effnet_fwd_{case}.npz"))
    model, W = _model()
    wav = torch.from_numpy(z["wav"]).cuda()
    assert model.get_model_layers() == OEF.hook_layer_names() == list(z["layer_names"])
    with torch.no_grad():
        feats = model(wav)
    assert tuple(feats.shape) == z["features"].shape
    names = model.register_hooks_for_layers(["all"])
    embs = model.extract_embeddings(wav, aggregation="none")
    agg = model.extract_embeddings(wav, aggregation="mean")
    assert tuple(agg.shape) == z["agg_mean"].shape
    report = []
    for n, e in zip(names, embs):
        ref = z["hook__" + n].astype(np.float32)
        assert tuple(e.shape) == ref.shape, n
        got = e.cpu().numpy()
        report.append((n, _cos(got, ref), float(np.abs(got - ref).max()), float(np.abs(ref).max())))
    fcos = _cos(feats.cpu().numpy(), z["features"])
    acos = _cos(agg.cpu().numpy(), z["agg_mean"])
    for n, c, e, m in report:
        print(f"  {n:36s} cos {c:.5f}  max|err| {e:.3e}  (|ref| max {m:.2f})")
    print(f"  features cos {fcos:.5f}; aggregated-mean embedding cos {acos:.5f}")
    # stem hook: fp32 math on the normalised image -> tight.  Every deeper layer, the features and the aggregated embedding:
    # north_star's per-layer cosine >= 0.999.  The network is random-init and BN-calibrated, i.e. it amplifies perturbations
    # (the reference itself under bf16 autocast scores 0.958 against its own fp32 run at the head; this
    # path with bf16 storage reached 0.997 / 0.98) -- fp16 storage and operands are what meet the gate, end to end, without
    # a separate fp32 mode.
    assert report[0][1] >= 0.99999 and report[0][2] <= 1e-3
    for n, c, e, m in report[1:]:
        assert c >= 0.999, (n, c)
    assert fcos >= 0.999 and acos >= 0.999


def test_effnet_vs_oracle_same_rounding_free_path():
    """Oracle (float64) on the GPU path's own mel image: isolates the CNN kernels from the front end."""
    model, W = _model()
    wav = (np.random.RandomState(21).standard_normal((2, 12000)) * 0.1).astype(np.float32)
    with torch.no_grad():
        feats = model(torch.from_numpy(wav).cuda()).cpu().numpy()
        img = model.process_audio(torch.from_numpy(wav).cuda())[:, 0].cpu().numpy()
    ora = OEF.forward(W, img)
    c = _cos(feats, ora["features"])
    print(f"features vs oracle: cos {c:.5f}, shape {feats.shape}")
    assert feats.shape == ora["features"].shape and c >= 0.999


def test_effnet_logits_vs_reference():
    z = np.load(os.path.join(G, "effnet_logits.npz"))
    model, _ = _model(num_classes=10)
    with torch.no_grad():
        logits = model(torch.from_numpy(z["wav"]).cuda()).cpu().numpy()
    assert logits.shape == z["logits"].shape
    c = _cos(logits, z["logits"])
    print(f"logits cos {c:.5f} max|err| {np.abs(logits - z['logits']).max():.3e} (|ref| max {np.abs(z['logits']).max():.2f})")
    assert c >= 0.999


def test_effnet_errors():
    from avex_b200 import _lib, plugin
    from avex_b200.plugin import efficientnet_model  # noqa: F401

    model, _ = _model()
    with pytest.raises(ValueError):
        model.extract_embeddings(torch.zeros(1, 16000, device="cuda"))  # no hooks registered
    with pytest.raises(_lib.AvexkError):
        model.train()
        model(torch.zeros(1, 16000, device="cuda"))
    model.eval()
    with pytest.raises(RuntimeError):
        plugin.build_model_from_spec(plugin.ModelSpec(name="efficientnet"), "cuda", pretrained=True)
    with pytest.raises(ValueError):
        model.register_hooks_for_layers([99])


def test_effnet_batch_independence_5s():
    """config #3 shape per clip (5 s -> [1280, 4, 16]); a clip's features do not depend on its batch neighbours."""
    model, _ = _model()
    wav = torch.randn(6, 80000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 0.1
    with torch.no_grad():
        a = model(wav)
        b = model(wav[2:3])
    assert tuple(a.shape) == (6, 1280, 4, 16)
    assert torch.allclose(a[2:3], b, atol=1e-5, rtol=1e-5)
