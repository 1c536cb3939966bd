"""Oracle: Kaldi-style log-mel filterbank (numpy).  Test infrastructure -- see oracle/__init__.py.

Restates avex/models/beats/beats.py:39-163 (`_BatchedFbank`), :304-323 (`BEATs.preprocess`) and the
EAT variant avex/models/eat/audio_processor.py:72-143.  `_BatchedFbank` is itself pinned by the
reference's tests to torchaudio.compliance.kaldi.fbank (tests/unittests/test_batched_fbank.py:29-150).
"""
from __future__ import annotations

import math

import numpy as np

FLOAT32_EPS = float(np.finfo(np.float32).eps)  # beats.py:36


def frame_count(num_samples: int, win: int = 400, hop: int = 160) -> int:
    """snip_edges framing, beats.py:136 (`unfold`): F = 1 + (T - win)//hop, 0 if T < win."""
    if num_samples < win:
        return 0
    return 1 + (num_samples - win) // hop


def povey_window(win_length: int = 400, dtype=np.float32) -> np.ndarray:
    """hann(win, periodic=False) ** 0.85, beats.py:75."""
    n = np.arange(win_length, dtype=np.float64)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / (win_length - 1))
    return (hann**0.85).astype(dtype)


def hanning_window(win_length: int = 400, dtype=np.float32) -> np.ndarray:
    """torchaudio kaldi `window_type="hanning"` = hann(win, periodic=False); eat/audio_processor.py:110-119."""
    n = np.arange(win_length, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * n / (win_length - 1))).astype(dtype)


def mel_filterbank(
    n_fft: int = 512,
    n_mels: int = 128,
    sample_rate: float = 16000.0,
    low_freq: float = 20.0,
    high_freq: float = 0.0,
    dtype=np.float32,
) -> np.ndarray:
    """Triangular mel filterbank [n_fft//2+1, n_mels], beats.py:82-118 (kaldi get_mel_banks).

    The reference evaluates this in float32 torch ops; here the *same formula* is evaluated in the
    requested dtype so that a float64 variant is available for the vs-f64 gate.
    """
    if high_freq <= 0.0:
        high_freq = sample_rate / 2.0 + high_freq  # beats.py:71-72
    num_fft_bins = n_fft // 2
    fft_bin_width = sample_rate / n_fft
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    mel_delta = (mel_high - mel_low) / (n_mels + 1)

    f = np.dtype(dtype).type
    bin_idx = np.arange(n_mels, dtype=np.int64)[:, None]
    # torch: python-float + int64 tensor * python-float  -> float32 arithmetic
    left = f(mel_low) + bin_idx.astype(dtype) * f(mel_delta)
    center = f(mel_low) + (bin_idx.astype(dtype) + f(1.0)) * f(mel_delta)
    right = f(mel_low) + (bin_idx.astype(dtype) + f(2.0)) * f(mel_delta)
    freqs = f(fft_bin_width) * np.arange(num_fft_bins, dtype=np.int64).astype(dtype)
    mel_freqs = (f(1127.0) * np.log(f(1.0) + freqs / f(700.0)))[None, :].astype(dtype)
    up = (mel_freqs - left) / (center - left)
    down = (right - mel_freqs) / (right - center)
    fb = np.maximum(f(0.0), np.minimum(up, down)).astype(dtype)
    fb = np.concatenate([fb, np.zeros((n_mels, 1), dtype=dtype)], axis=1)  # Nyquist column, beats.py:117
    return np.ascontiguousarray(fb.T)


def fbank(
    wav: np.ndarray,
    *,
    n_mels: int = 128,
    window: str = "povey",
    preemph: float = 0.97,
    dtype=np.float32,
    win: int = 400,
    hop: int = 160,
    sample_rate: float = 16000.0,
) -> np.ndarray:
    """`_BatchedFbank.forward`, beats.py:120-163.  wav [B,T] (already scaled) -> [B,F,n_mels].

    dtype=np.float64 gives the high-precision evaluation of the same formulas used by the
    "no worse than the reference vs f64" gate (hard parts).
    """
    wav = np.asarray(wav, dtype=dtype)
    if wav.ndim == 1:
        wav = wav[None]
    B, T = wav.shape
    F = frame_count(T, win, hop)
    if F == 0:
        return np.zeros((B, 0, n_mels), dtype=dtype)
    n_fft = 1
    while n_fft < win:
        n_fft *= 2  # beats.py:65-68
    idx = (np.arange(F)[:, None] * hop + np.arange(win)[None, :]).astype(np.int64)
    frames = wav[:, idx]  # [B,F,win]                                        beats.py:136
    frames = frames - frames.mean(axis=-1, keepdims=True, dtype=dtype)  # beats.py:140
    shifted = np.concatenate([frames[..., :1], frames[..., :-1]], axis=-1)  # replicate pad, beats.py:143
    frames = frames - np.dtype(dtype).type(preemph) * shifted  # beats.py:144
    w = povey_window(win, dtype) if window == "povey" else hanning_window(win, dtype)
    frames = frames * w  # beats.py:147
    spec = np.fft.rfft(frames, n=n_fft, axis=-1)  # zero-pad to n_fft, beats.py:150-154
    power = (spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2).astype(dtype)  # beats.py:155
    mel = power @ mel_filterbank(n_fft, n_mels, sample_rate, dtype=dtype)  # beats.py:159
    return np.log(np.maximum(mel, np.dtype(dtype).type(FLOAT32_EPS))).astype(dtype)  # beats.py:163


def beats_preprocess(wav: np.ndarray, mean: float = 15.41663, std: float = 6.55582, dtype=np.float32) -> np.ndarray:
    """`BEATs.preprocess`, beats.py:304-323: fbank(wav * 2**15) then (x - mean) / (2 std)."""
    fb = fbank(np.asarray(wav, dtype=dtype) * np.dtype(dtype).type(32768.0), dtype=dtype)
    f = np.dtype(dtype).type
    return ((fb - f(mean)) / f(2.0 * std)).astype(dtype)


def eat_preprocess(
    wav: np.ndarray,
    target_frames: int = 1024,
    norm_mean: float = -4.268,
    norm_std: float = 4.569,
    dtype=np.float32,
) -> np.ndarray:
    """`EATAudioProcessor.__call__`, eat/audio_processor.py:72-143.

    per-clip DC removal (:103-104) -> kaldi fbank, hanning window, no 2**15 scaling (:110-119) ->
    zero-pad / truncate to `target_frames` (:121-126) -> (x-mean)/(2 std) with constants, or
    per-utterance mean / unbiased std when (mean,std)==(0,1) (:128-138).
    """
    wav = np.asarray(wav, dtype=dtype)
    if wav.ndim == 1:
        wav = wav[None]
    wav = wav - wav.mean(axis=-1, keepdims=True, dtype=dtype)
    fb = fbank(wav, window="hanning", dtype=dtype)
    B, F, M = fb.shape
    if F < target_frames:
        fb = np.concatenate([fb, np.zeros((B, target_frames - F, M), dtype=dtype)], axis=1)
    else:
        fb = fb[:, :target_frames]
    f = np.dtype(dtype).type
    if norm_mean == 0.0 and norm_std == 1.0:
        out = np.empty_like(fb)
        for b in range(B):
            mu = fb[b].mean(dtype=dtype)
            sd = fb[b].std(ddof=1, dtype=dtype)
            if sd == 0:
                sd = f(1.0)
            out[b] = (fb[b] - mu) / (f(2.0) * sd)
        return out
    return ((fb - f(norm_mean)) / f(2.0 * norm_std)).astype(dtype)
