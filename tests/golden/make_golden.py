#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (earthspecies/avex,
imported from /root/reference through tools/ref_shim.py) in the build container.

    python tests/golden/make_golden.py            # writes *.npz + REPORT.json next to this script

The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed as fixtures;
tests/ compare the numpy oracle (CPU) and the CUDA path (GPU) against them.  The script also prints how
far the oracle is from the reference on every case (recorded in REPORT.json).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from tests.golden import cases  # noqa: E402  (before the shim: the reference has its own `tests` package)

import ref_shim  # noqa: E402

avex = ref_shim.install()

import torch  # noqa: E402

from oracle import beats_encoder as OE  # noqa: E402
from oracle import kaldi_fbank as OF  # noqa: E402
from oracle import relpos as OR  # noqa: E402
from oracle.weights import make_beats_weights, make_predictor_weights  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
REPORT: dict = {"reference": "earthspecies/avex v1.2.0 @ /root/reference", "torch": torch.__version__, "cases": {}}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def diff(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return {
        "max_abs": float(np.abs(a - b).max()),
        "rel_l2": float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)),
        "frac_out_1e-4": float((np.abs(a - b) > 1e-4 + 1e-4 * np.abs(b)).mean()),
    }


# ---------------------------------------------------------------------------------------------
# 1. fbank: reference _BatchedFbank (pinned to torchaudio kaldi by the reference's own tests)
# ---------------------------------------------------------------------------------------------
def gen_fbank():
    from avex.models.beats.beats import BEATs, BEATsConfig, _BatchedFbank
    import torchaudio.compliance.kaldi as ta_kaldi

    out = {}
    fb128 = _BatchedFbank(num_mel_bins=128)
    np.savez_compressed(
        os.path.join(HERE, "fbank_tables.npz"),
        window=fb128.window.numpy(),
        mel_fb=fb128.mel_fb.numpy(),
    )
    REPORT["cases"]["fbank_tables"] = {
        "window_vs_oracle": diff(OF.povey_window(), fb128.window.numpy()),
        "mel_vs_oracle": diff(OF.mel_filterbank(), fb128.mel_fb.numpy()),
    }
    beats = BEATs(BEATsConfig())
    for name, (wav, n_mels) in cases.fbank_cases().items():
        w = torch.from_numpy(wav)
        mod = _BatchedFbank(num_mel_bins=n_mels)
        ref = mod(w * 2**15).numpy()
        # the reference's own equivalence (test_batched_fbank.py:29-50): torchaudio kaldi, per sample
        kal = np.stack(
            [
                ta_kaldi.fbank(
                    (x * 2**15).unsqueeze(0), num_mel_bins=n_mels, sample_frequency=16000, frame_length=25, frame_shift=10
                ).numpy()
                for x in w
            ]
        )
        orc = OF.fbank(wav * np.float32(32768.0), n_mels=n_mels)
        f64 = OF.fbank(wav.astype(np.float64) * 32768.0, n_mels=n_mels, dtype=np.float64)
        entry = {
            "input_sha": sha(wav),
            "shape": list(ref.shape),
            "ref_vs_torchaudio_kaldi": diff(ref, kal),
            "oracle_vs_ref": diff(orc, ref),
            "ref_vs_f64": diff(ref, f64),
            "oracle_vs_f64": diff(orc, f64),
        }
        if n_mels == 128:
            pre = beats.preprocess(w).numpy()  # includes (x-mean)/(2 std), beats.py:304-323
            entry["preprocess_oracle_vs_ref"] = diff(OF.beats_preprocess(wav), pre)
            out[name + "__pre"] = pre.astype(np.float32)
        out[name] = ref.astype(np.float32)
        REPORT["cases"]["fbank/" + name] = entry
        print("fbank", name, entry["oracle_vs_ref"], "ref_vs_f64", entry["ref_vs_f64"]["max_abs"])
    np.savez_compressed(os.path.join(HERE, "fbank.npz"), **out)

    # EAT variant (eat/audio_processor.py) -- hanning window, clip DC removal, pad to 1024, const / per-utt norm
    from avex.models.eat.audio_processor import EATAudioProcessor

    eat_out = {}
    wav = cases.eat_case()
    for tag, (m, s) in {"const": (-4.268, 4.569), "perutt": (0.0, 1.0)}.items():
        proc = EATAudioProcessor(sample_rate=16000, target_length=1024, n_mels=128, norm_mean=m, norm_std=s)
        ref = proc(torch.from_numpy(wav)).numpy()
        orc = OF.eat_preprocess(wav, 1024, m, s)
        eat_out[tag] = ref.astype(np.float32)
        REPORT["cases"]["eat_fbank/" + tag] = {"input_sha": sha(wav), "oracle_vs_ref": diff(orc, ref)}
        print("eat", tag, diff(orc, ref))
    np.savez_compressed(os.path.join(HERE, "eat_fbank.npz"), **eat_out)


# ---------------------------------------------------------------------------------------------
# 2. relative-position buckets
# ---------------------------------------------------------------------------------------------
def gen_relpos():
    from avex.models.beats.backbone import _MultiheadAttention

    att = _MultiheadAttention(768, 12, has_relative_attention_bias=True, num_buckets=320, max_distance=800, gru_rel_pos=True)
    rel = torch.arange(-3100, 3101, dtype=torch.long)
    b = att._relative_positions_bucket(rel[None, :], bidirectional=True)[0].numpy().astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "relpos_buckets.npz"), rel=rel.numpy().astype(np.int32), bucket=b)
    mine = OR.relative_position_bucket(rel.numpy())
    REPORT["cases"]["relpos"] = {"n": int(rel.numel()), "oracle_mismatches": int((mine != b).sum()), "distinct": int(len(np.unique(b)))}
    print("relpos mismatches", REPORT["cases"]["relpos"])


# ---------------------------------------------------------------------------------------------
# 3. BEATs encoder through the reference's own plugin API (register_model -> load_model)
# ---------------------------------------------------------------------------------------------
def build_ref_beats(layers: int, W: dict):
    spec = avex.get_model_spec("esp_aves2_sl_beats_all").model_copy(deep=True)
    spec.init_config = dict(spec.init_config, encoder_layers=layers)
    name = f"golden_beats_L{layers}"
    avex.register_model(name, spec)
    model = avex.load_model(name, device="cpu", return_features_only=True).eval()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in W.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [k for k in missing if not (k.startswith("backbone.fbank.") or k.startswith("backbone.predictor."))]
    assert not missing and not unexpected, (missing, unexpected)
    return model


def gen_beats():
    for cname, case in cases.beats_cases().items():
        dims = OE.BeatsDims(layers=case["layers"])
        W = make_beats_weights(dims, seed=case["wseed"], init=case.get("init", "perturbed"))
        model = build_ref_beats(case["layers"], W)
        wav = case["wav"]
        mask = case.get("mask")
        names = model.register_hooks_for_layers(["all"])
        assert len(names) == case["layers"] + 1
        x = torch.from_numpy(wav)
        m = torch.from_numpy(mask) if mask is not None else None
        with torch.no_grad():
            feats = model(x, m).numpy()
        hooks = model.extract_embeddings(x, padding_mask=m, aggregation="none")
        hooks = [h.numpy() for h in hooks]  # batch-first after the wrapper's transpose heuristic
        pooled_hooks = model.extract_embeddings(x, padding_mask=m, aggregation="mean").numpy()
        orc = OE.beats_forward(W, wav, mask, dims)
        entry = {"input_sha": sha(wav), "layers": case["layers"], "tokens": int(feats.shape[1])}
        entry["final_oracle_vs_ref"] = diff(orc["x"], feats)
        entry["hook0_oracle_vs_ref"] = diff(orc["hook0"], hooks[0])
        entry["fc2_last_oracle_vs_ref"] = diff(orc["fc2"][-1], hooks[-1])
        save = {"final": feats.astype(np.float32), "pooled_hooks_mean": pooled_hooks.astype(np.float32)}
        for li in case["keep_hooks"]:
            save[f"hook{li}"] = hooks[li].astype(np.float32)
        if mask is not None:
            # token mask as the reference computes it (beats.py:283-302, twice)
            bk = model.backbone
            fbm = bk.forward_padding_mask(torch.zeros(x.shape[0], OF.frame_count(x.shape[1]), 1), m)
            tkm = bk.forward_padding_mask(torch.zeros(x.shape[0], feats.shape[1], 1), fbm)
            save["key_pad"] = tkm.numpy()
            entry["key_pad_oracle_matches"] = bool((orc["key_pad"] == tkm.numpy()).all())
        # classifier-mode output (masked mean-pool + Linear), beats_model.py:266-277
        REPORT["cases"]["beats/" + cname] = entry
        print("beats", cname, entry)
        np.savez_compressed(os.path.join(HERE, f"beats_{cname}.npz"), **save)


def gen_beats_long():
    """One unmasked 60 s clip (config #5 shape) through the reference; token axis sub-sampled in the fixture."""
    case = cases.beats_long_case()
    dims = OE.BeatsDims(layers=case["layers"])
    W = make_beats_weights(dims, seed=case["wseed"])
    model = build_ref_beats(case["layers"], W)
    x = torch.from_numpy(case["wav"])
    st = case["stride"]
    model.register_hooks_for_layers(["all"])
    with torch.no_grad():
        feats = model(x).numpy()
    hooks = [h.numpy() for h in model.extract_embeddings(x, aggregation="none")]
    pooled_hooks = model.extract_embeddings(x, aggregation="mean").numpy()
    assert feats.shape == (1, 2992, 768), feats.shape
    save = {"final": feats[:, ::st].astype(np.float32), "final_pooled": feats.mean(axis=1).astype(np.float32),
            "pooled_hooks_mean": pooled_hooks.astype(np.float32)}
    for li in case["keep_hooks"]:
        save[f"hook{li}"] = hooks[li][:, ::st].astype(np.float32)
    orc = OE.beats_forward(W, case["wav"], None, dims)
    entry = {"input_sha": sha(case["wav"]), "layers": case["layers"], "tokens": int(feats.shape[1]), "stride": st,
             "final_oracle_vs_ref": diff(orc["x"], feats), "fc2_last_oracle_vs_ref": diff(orc["fc2"][-1], hooks[-1])}
    REPORT["cases"]["beats/L2_1x60s"] = entry
    print("beats long", entry)
    np.savez_compressed(os.path.join(HERE, "beats_L2_1x60s.npz"), **save)


def gen_predictor():
    """`BEATs.extract_features(feature_only=False)` with the predictor head (beats.py:369-380), with and without a padding mask."""
    case = cases.predictor_case()
    dims = OE.BeatsDims(layers=case["layers"])
    W = make_beats_weights(dims, seed=case["wseed"])
    P = make_predictor_weights(case["pseed"])
    model = build_ref_beats(case["layers"], {**W, **P})
    assert model.backbone.predictor is not None
    assert torch.equal(model.backbone.predictor.weight, torch.from_numpy(P["backbone.predictor.weight"]))
    x, m = torch.from_numpy(case["wav"]), torch.from_numpy(case["mask"])
    with torch.no_grad():
        lg_mask, km = model.backbone.extract_features(x, m, feature_only=False)
        lg_none, _ = model.backbone.extract_features(x, None, feature_only=False)
    assert lg_mask.shape == (2, 527) and km.any()
    orc = OE.beats_forward(W, case["wav"], case["mask"], dims)
    lg = orc["x"] @ P["backbone.predictor.weight"].T + P["backbone.predictor.bias"]
    lg[orc["key_pad"]] = 0
    mine = lg.sum(1) / (~orc["key_pad"]).sum(1)[:, None]
    entry = {"input_sha": sha(case["wav"]), "oracle_vs_ref_masked": diff(mine, lg_mask.numpy())}
    REPORT["cases"]["beats/predictor"] = entry
    print("predictor", entry)
    np.savez_compressed(os.path.join(HERE, "beats_predictor.npz"), logits_mask=lg_mask.numpy().astype(np.float32),
                        logits_nomask=lg_none.numpy().astype(np.float32))


if __name__ == "__main__":
    which = sys.argv[1:] or ["fbank", "relpos", "beats", "beats_long", "predictor"]
    rp = os.path.join(HERE, "REPORT.json")
    if os.path.exists(rp):
        try:
            REPORT["cases"].update(json.load(open(rp)).get("cases", {}))
        except Exception:
            pass
    if "fbank" in which:
        gen_fbank()
    if "relpos" in which:
        gen_relpos()
    if "beats" in which:
        gen_beats()
    if "beats_long" in which:
        gen_beats_long()
    if "predictor" in which:
        gen_predictor()
    with open(rp, "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)
    print("wrote", rp)
