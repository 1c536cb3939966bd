"""`AudioProcessor` for the representations the hot path uses (avex/data/audio_utils.py:76-179).

"raw" returns the waveform unchanged (audio_utils.py:130-131) -- this is what every BEATs config uses.
The mel-spectrogram representation (EfficientNet path) is served by the STFT-mel kernel.
"""
from __future__ import annotations

import torch


class AudioProcessor:
    def __init__(self, cfg) -> None:
        self.cfg = cfg
        self.sr = cfg.sample_rate
        self.representation = cfg.representation
        self.n_fft = cfg.n_fft
        self.hop_length = cfg.hop_length or cfg.n_fft // 4
        self.win_length = cfg.win_length or cfg.n_fft
        self.n_mels = cfg.n_mels
        self.normalize = cfg.normalize
        self.target_length_seconds = cfg.target_length_seconds
        self._mel = None

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if waveform.dim() == 1:
            waveform = waveform.unsqueeze(0)  # audio_utils.py:126-128: [T] -> [1, T]
        if self.representation == "raw":
            return waveform
        if self.representation == "mel_spectrogram" and (self.n_fft, self.hop_length, self.win_length, self.n_mels) == (800, 160, 800, 128):
            if self._mel is None:
                from ..melspec import MelSpectrogram

                self._mel = MelSpectrogram()
            return self._mel.run(waveform, normalize=bool(self.normalize))  # [B, 128, frames], audio_utils.py:137-155
        raise NotImplementedError(
            f"avex_b200: representation {self.representation!r} is not on the BEATs hot path "
            "(the EfficientNet mel front end has its own kernel entry point)"
        )
