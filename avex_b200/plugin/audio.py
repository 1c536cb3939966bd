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

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if self.representation == "raw":
            return waveform
        raise NotImplementedError(
            f"avex_b200: representation {self.representation!r} is not on the BEATs hot path "
            "(the EfficientNet mel front end has its own kernel entry point)"
        )
