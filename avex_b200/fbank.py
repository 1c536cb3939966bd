"""Host side of the fused log-mel filterbank: mirrors `_BatchedFbank` (avex/models/beats/beats.py:39-163).

Same constructor arguments, same registered buffers (`window`, `mel_fb`, so `state_dict()` keys match the
reference's `backbone.fbank.*`), same call convention (`forward(waveforms [B,T]) -> [B,F,n_mels]`).  The tables
are built on the host with the reference's fp32 formulas; all per-sample arithmetic runs in the CUDA kernel
behind `avexk_fbank_forward`.  There is no torch fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib

FLOAT32_EPS = torch.finfo(torch.float32).eps


def kaldi_mel_matrix(n_fft: int, n_mels: int, sample_rate: float, low_freq: float, high_freq: float) -> torch.Tensor:
    """Triangular mel weights [n_fft//2+1, n_mels] in the mel domain (1127 ln(1 + f/700)); beats.py:82-118."""
    half = n_fft // 2
    lo = 1127.0 * math.log(1.0 + low_freq / 700.0)
    hi = 1127.0 * math.log(1.0 + high_freq / 700.0)
    step = (hi - lo) / (n_mels + 1)
    m = torch.arange(n_mels).unsqueeze(1)
    left, center, right = lo + m * step, lo + (m + 1.0) * step, lo + (m + 2.0) * step
    mel_of_bin = (1127.0 * (1.0 + (sample_rate / n_fft) * torch.arange(half) / 700.0).log()).unsqueeze(0)
    rising = (mel_of_bin - left) / (center - left)
    falling = (right - mel_of_bin) / (right - center)
    tri = torch.max(torch.zeros(1), torch.min(rising, falling))
    tri = torch.nn.functional.pad(tri, (0, 1), value=0.0)  # Nyquist bin carries no weight
    return tri.T.contiguous()


class KaldiFbank(nn.Module):
    """Batched Kaldi fbank on the GPU (16 kHz, 25/10 ms, n_fft 512, 128 mel bins)."""

    def __init__(
        self,
        num_mel_bins: int = 128,
        sample_frequency: float = 16000.0,
        frame_length_ms: float = 25.0,
        frame_shift_ms: float = 10.0,
        preemphasis_coefficient: float = 0.97,
        low_freq: float = 20.0,
        high_freq: float = 0.0,
        window_type: str = "povey",
    ) -> None:
        super().__init__()
        self.win_length = int(sample_frequency * frame_length_ms / 1000.0)
        self.hop_length = int(sample_frequency * frame_shift_ms / 1000.0)
        n_fft = 1
        while n_fft < self.win_length:
            n_fft *= 2
        self.n_fft = n_fft
        self.num_mel_bins = num_mel_bins
        self.preemphasis_coefficient = preemphasis_coefficient
        if (self.win_length, self.hop_length, n_fft, num_mel_bins) != (400, 160, 512, 128) or preemphasis_coefficient != 0.97:
            raise ValueError(
                "avex_b200 fbank kernel is specialised to the reference geometry "
                "(16 kHz, 25 ms / 10 ms, 128 mel bins, pre-emphasis 0.97)"
            )
        if high_freq <= 0.0:
            high_freq = sample_frequency / 2.0 + high_freq
        window = torch.hann_window(self.win_length, periodic=False)
        if window_type == "povey":
            window = window.pow(0.85)  # beats.py:75
        elif window_type != "hanning":
            raise ValueError(f"unsupported window_type {window_type!r}")
        self.register_buffer("window", window)
        self.register_buffer("mel_fb", kaldi_mel_matrix(n_fft, num_mel_bins, sample_frequency, low_freq, high_freq))
        self._handle = None
        self._handle_device = None

    # -- C-ABI handle ------------------------------------------------------------------------------------
    def handle(self, device: torch.device) -> C.c_void_p:
        if self._handle is None or self._handle_device != device:
            self.close()
            lib = _lib.load()
            win = self.window.detach().to("cpu", torch.float32).contiguous()
            mel = self.mel_fb.detach().to("cpu", torch.float32).contiguous()
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(lib.avexk_fbank_create(win.data_ptr(), mel.data_ptr(), C.byref(h)), "avexk_fbank_create")
            self._handle, self._handle_device = h, device
        return self._handle

    def close(self) -> None:
        if self._handle is not None:
            _lib.load().avexk_fbank_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # the native handle is per process / device: copies and pickles carry the buffers only and re-create it lazily
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_handle"], st["_handle_device"] = None, None
        return st

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__getstate__().items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def num_frames(self, num_samples: int) -> int:
        return 0 if num_samples < self.win_length else 1 + (num_samples - self.win_length) // self.hop_length

    # -- compute -----------------------------------------------------------------------------------------
    def run(
        self,
        waveforms: torch.Tensor,
        *,
        prescale: float = 1.0,
        norm_mean: float = 0.0,
        norm_std2: float = 1.0,
        out_frames: int = 0,
        per_utterance: bool = False,
        out_dtype: torch.dtype = torch.float32,
    ) -> torch.Tensor:
        """out = (log(max(mel(prescale * wav), eps)) - norm_mean) / norm_std2, shape [B, frames, 128]."""
        if waveforms.dim() != 2:
            raise ValueError(f"expected [B, T] waveforms, got {tuple(waveforms.shape)}")
        if not waveforms.is_cuda:
            raise _lib.AvexkError("avex_b200 fbank runs on CUDA tensors only (no CPU fallback)")
        x = waveforms
        if x.dtype != torch.float32:
            x = x.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        B, T = x.shape
        if T < self.win_length:
            raise RuntimeError(f"waveform too short for one frame: {T} < {self.win_length}")  # unfold raises too
        F = out_frames if out_frames > 0 else self.num_frames(T)
        out = torch.empty((B, F, self.num_mel_bins), device=x.device, dtype=out_dtype)
        stats = torch.empty((B, 2), device=x.device, dtype=torch.float64) if per_utterance else None
        lib = _lib.load()
        with torch.cuda.device(x.device):
            rc = lib.avexk_fbank_forward(
                self.handle(x.device), x.data_ptr(), B, T, x.stride(0) if B > 1 else max(T, x.stride(0)),
                float(prescale), float(norm_mean), float(1.0 / norm_std2), int(out_frames), int(per_utterance),
                stats.data_ptr() if stats is not None else None, out.data_ptr(),
                int(out_dtype == torch.bfloat16), torch.cuda.current_stream(x.device).cuda_stream,
            )  # fmt: skip
        _lib.check(rc, "avexk_fbank_forward")
        return out

    def patch_operand(self, waveforms: torch.Tensor, *, prescale: float = 1.0, norm_mean: float = 0.0, norm_std2: float = 1.0) -> torch.Tensor:
        """The normalised fbank as the bf16 [hi|lo|hi] operand of the 16x16 patch-embedding GEMM, [B * N, 768] with
        N = 8 * (frames // 16), written directly by the fbank kernel (what `avexk_beats_forward` uses internally)."""
        if waveforms.dim() != 2 or not waveforms.is_cuda:
            raise _lib.AvexkError("patch_operand expects [B, T] CUDA waveforms (no CPU fallback)")
        x = waveforms if waveforms.dtype == torch.float32 else waveforms.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        B, T = x.shape
        n_tok = 8 * (self.num_frames(T) // 16)
        out = torch.empty((B * n_tok, 768), device=x.device, dtype=torch.bfloat16)
        with torch.cuda.device(x.device):
            rc = _lib.load().avexk_fbank_patch_operand(
                self.handle(x.device), x.data_ptr(), B, T, x.stride(0) if B > 1 else max(T, x.stride(0)), float(prescale),
                float(norm_mean), float(1.0 / norm_std2), out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream,
            )  # fmt: skip
        _lib.check(rc, "avexk_fbank_patch_operand")
        return out

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """`_BatchedFbank.forward` (beats.py:120-163): waveforms already scaled by 2**15 -> raw log-mel."""
        return self.run(waveforms)
