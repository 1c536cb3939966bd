#!/usr/bin/env python
"""Golden vectors for the EfficientNet path, produced by running the UNMODIFIED reference (earthspecies/avex from
/root/reference through tools/ref_shim.py; torchvision 0.26.0 / torchaudio 2.11.0 as pinned by its uv.lock).

    python tests/golden/make_golden_effnet.py     # writes effnet_*.npz + REPORT_effnet.json next to this script

* effnet_bn_stats.npz  BatchNorm running statistics from one calibration pass (train mode, momentum 1) of the reference
                       module holding `oracle.weights.make_effnet_weights(seed=3)`: a random-init EfficientNet with
                       identity statistics collapses to 1e-13 at the head, so parity would be noise.
* effnet_mel.npz       `Model.process_audio` (AudioProcessor mel, audio_utils.py:106-172) on seeded waveforms.
* effnet_fwd_*.npz     `Model.forward` features, the 17 hooked pre-BN conv outputs (`register_hooks_for_layers(["all"])`,
                       `extract_embeddings(aggregation="none")`), aggregated embeddings, classifier logits.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from tests.golden import cases  # noqa: E402,F401  (import before the shim: the reference has its own `tests` package)

import ref_shim  # noqa: E402

avex = ref_shim.install()

import torch  # noqa: E402

from oracle import effnet as OEF  # noqa: E402
from oracle import melspec as OM  # noqa: E402
from oracle.weights import make_effnet_weights  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
REPORT: dict = {"reference": "earthspecies/avex v1.2.0 @ /root/reference", "torch": torch.__version__, "cases": {}}


def diff(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    cos = float((a * b).sum() / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
    return {"max_abs": float(np.abs(a - b).max()), "rel_l2": float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)),
            "cos": cos, "ref_abs_max": float(np.abs(b).max())}


def wave(kind: str, b: int, t: int, seed: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return (rs.standard_normal((b, t)) * 0.1).astype(np.float32)
    n = np.arange(t, dtype=np.float64) / 16000.0
    out = []
    for i in range(b):
        f0 = 300.0 * (i + 1)
        x = 0.3 * np.sin(2 * np.pi * f0 * n) + 0.1 * np.sin(2 * np.pi * (2500.0 + 700 * i) * n * (1 + 0.2 * n))
        x = x * (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * n)) + 0.01 * rs.standard_normal(t)
        out.append(x)
    return np.stack(out).astype(np.float32)


def build_reference(num_classes=None):
    from avex.models.utils.factory import build_model_from_spec
    from avex.models.utils.registry import get_model_spec

    spec = get_model_spec("esp_aves2_effnetb0_all").model_copy(deep=True)
    kw = dict(pretrained=False, return_features_only=num_classes is None)
    if num_classes is not None:
        kw["num_classes"] = num_classes
    return build_model_from_spec(spec, "cpu", **kw)


def main():
    # ---- weights + BatchNorm calibration with the reference module ---------------------------------------------
    W = make_effnet_weights(seed=3)
    ref = build_reference()
    missing = ref.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=False)
    assert not [k for k in missing.missing_keys if "classifier" not in k], missing
    assert not missing.unexpected_keys, missing
    calib = torch.from_numpy(np.concatenate([wave("noise", 3, 32000, 11), wave("tones", 3, 32000, 12)]))
    ref.train()
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    with torch.no_grad():
        ref(calib)
    ref.eval()
    stats = {k: v.numpy().astype(np.float32) for k, v in ref.state_dict().items() if k.endswith(("running_mean", "running_var"))}
    np.savez_compressed(os.path.join(HERE, "effnet_bn_stats.npz"), **stats)
    W = make_effnet_weights(seed=3, bn_stats=stats)
    ref.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}, strict=False)
    ref.eval()
    assert ref.get_model_layers() == OEF.hook_layer_names(), (ref.get_model_layers(), OEF.hook_layer_names())

    # ---- mel front end ---------------------------------------------------------------------------------------------
    mel_cases = {"noise_2x1s": wave("noise", 2, 16000, 1), "tones_2x1s": wave("tones", 2, 16000, 2),
                 "noise_1x5s": wave("noise", 1, 80000, 3), "ragged_1x8123": wave("tones", 1, 8123, 4)}
    mel_out = {}
    for name, w in mel_cases.items():
        with torch.no_grad():
            img = ref.process_audio(torch.from_numpy(w))  # [B, 3, 128, frames]
        assert torch.equal(img[:, 0], img[:, 1]) and torch.equal(img[:, 0], img[:, 2])
        mel_out[name + "__wav"] = w
        mel_out[name + "__mel"] = img[:, 0].numpy()
        REPORT["cases"]["mel_" + name] = {"oracle_f64_vs_ref": diff(OM.mel_spectrogram(w), img[:, 0].numpy()),
                                          "shape": list(img.shape)}
    np.savez_compressed(os.path.join(HERE, "effnet_mel.npz"), **mel_out)

    # ---- forward: features + hooks -----------------------------------------------------------------------------------
    for name, w in {"noise_2x1s": mel_cases["noise_2x1s"], "tones_1x2s": wave("tones", 1, 32000, 5)}.items():
        x = torch.from_numpy(w)
        with torch.no_grad():
            feats = ref(x).numpy()
        names = ref.register_hooks_for_layers(["all"])
        embs = ref.extract_embeddings(x, aggregation="none")
        agg = ref.extract_embeddings(x, aggregation="mean").numpy()
        ref.deregister_all_hooks()
        ora = OEF.forward(W, OM.mel_spectrogram(w))
        rep = {"features": diff(ora["features"], feats), "hooks": {}}
        out = {"wav": w, "features": feats, "agg_mean": agg, "layer_names": np.array(names)}
        for n, e in zip(names, embs):
            out["hook__" + n] = e.numpy().astype(np.float16) if e.numel() > 100000 else e.numpy()
            rep["hooks"][n] = diff(ora["hooks"][n], e.numpy())
        REPORT["cases"]["fwd_" + name] = rep
        np.savez_compressed(os.path.join(HERE, f"effnet_fwd_{name}.npz"), **out)

    # ---- classifier mode -------------------------------------------------------------------------------------------
    Wc = make_effnet_weights(seed=3, num_classes=10, bn_stats=stats)
    refc = build_reference(num_classes=10)
    refc.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in Wc.items()}, strict=True)
    refc.eval()
    w = mel_cases["tones_2x1s"]
    with torch.no_grad():
        logits = refc(torch.from_numpy(w)).numpy()
    ora = OEF.forward(Wc, OM.mel_spectrogram(w), want_logits=True)
    REPORT["cases"]["logits_tones_2x1s"] = diff(ora["logits"], logits)
    np.savez_compressed(os.path.join(HERE, "effnet_logits.npz"), wav=w, logits=logits)

    json.dump(REPORT, open(os.path.join(HERE, "REPORT_effnet.json"), "w"), indent=1)
    print(json.dumps(REPORT, indent=1)[:6000])


if __name__ == "__main__":
    main()
