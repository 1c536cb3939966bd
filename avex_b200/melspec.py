"""Host side of the STFT mel-spectrogram kernel: the `AudioProcessor` mel path of the EfficientNet models
(avex/data/audio_utils.py:76-172 with api/configs/official_models/esp_aves2_effnetb0_all.yml).

The window and the mel filterbank are built on the host with the reference's own ops (`torch.hann_window`,
`torchaudio.functional.melscale_fbanks` -- exactly what `torchaudio.transforms.MelScale` holds); all per-sample
arithmetic runs in the CUDA kernel behind `avexk_melspec_forward`.  There is no torch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

N_FFT, HOP, N_MELS, SAMPLE_RATE = 800, 160, 128, 16000


def mel_tables() -> tuple[torch.Tensor, torch.Tensor]:
    import torchaudio

    window = torch.hann_window(N_FFT)  # audio_utils.py:163-165 (periodic)
    fb = torchaudio.functional.melscale_fbanks(N_FFT // 2 + 1, 0.0, float(SAMPLE_RATE // 2), N_MELS, SAMPLE_RATE)  # MelScale defaults
    return window.contiguous(), fb.contiguous()


class MelSpectrogram:
    """[B, T] waveform -> [B, 128, 1 + T // 160] log-mel image (optionally min-max normalised per clip)."""

    def __init__(self) -> None:
        self.window, self.mel_fb = mel_tables()
        self._handle = None
        self._handle_device = None

    def handle(self, device: torch.device) -> C.c_void_p:
        if self._handle is None or self._handle_device != device:
            self.close()
            h = C.c_void_p()
            with torch.cuda.device(device):
                _lib.check(_lib.load().avexk_melspec_create(self.window.data_ptr(), self.mel_fb.data_ptr(), C.byref(h)), "avexk_melspec_create")
            self._handle, self._handle_device = h, device
        return self._handle

    def close(self) -> None:
        if self._handle is not None:
            _lib.load().avexk_melspec_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_handle"], st["_handle_device"] = None, None
        return st

    @staticmethod
    def num_frames(num_samples: int) -> int:
        return 1 + num_samples // HOP

    def run(self, waveforms: torch.Tensor, normalize: bool = True, return_minmax: bool = False):
        if waveforms.dim() == 1:
            waveforms = waveforms.unsqueeze(0)
        if waveforms.dim() != 2:
            raise ValueError(f"expected [B, T] waveforms, got {tuple(waveforms.shape)}")
        if not waveforms.is_cuda:
            raise _lib.AvexkError("avex_b200 mel spectrogram runs on CUDA tensors only (no CPU fallback)")
        x = waveforms if waveforms.dtype == torch.float32 else waveforms.float()
        if x.stride(1) != 1:
            x = x.contiguous()
        B, T = x.shape
        if T <= N_FFT // 2:
            raise RuntimeError(f"reflect padding needs more than {N_FFT // 2} samples, got {T}")  # torch.stft raises too
        out = torch.empty((B, N_MELS, self.num_frames(T)), device=x.device, dtype=torch.float32)
        minmax = torch.empty((B, 2), device=x.device, dtype=torch.int32)
        with torch.cuda.device(x.device):
            rc = _lib.load().avexk_melspec_forward(
                self.handle(x.device), x.data_ptr(), B, T, x.stride(0) if B > 1 else max(T, x.stride(0)), int(normalize),
                out.data_ptr(), minmax.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream,
            )  # fmt: skip
        _lib.check(rc, "avexk_melspec_forward")
        return (out, minmax) if return_minmax else out

    __call__ = run
