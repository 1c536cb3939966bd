"""`AudioConfig` / `ModelSpec` with the fields the hot path reads (avex/configs.py:47-168, :170-378)."""
from __future__ import annotations

import copy
from dataclasses import dataclass, field, fields
from typing import Any, Optional


@dataclass
class AudioConfig:
    sample_rate: int = 16000
    n_fft: int = 2048
    hop_length: Optional[int] = None
    win_length: Optional[int] = None
    window: str = "hann"
    n_mels: int = 128
    representation: str = "mel_spectrogram"  # "spectrogram" | "mel_spectrogram" | "raw"
    normalize: bool = True
    target_length_seconds: int = 10
    window_selection: str = "random"
    center: bool = True

    def __post_init__(self):
        if self.representation not in ("spectrogram", "mel_spectrogram", "raw"):
            raise ValueError(f"representation must be 'spectrogram', 'mel_spectrogram' or 'raw', got {self.representation!r}")


@dataclass
class ModelSpec:
    name: str = ""
    pretrained: bool = False
    device: str = "cuda"
    audio_config: Optional[Any] = None
    use_naturelm: Optional[bool] = None
    fine_tuned: Optional[bool] = None
    init_config: Optional[dict] = None
    efficientnet_variant: Optional[str] = None
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        # avex/configs.py:355-372: only "cpu" or "cuda" validate; per-rank placement is torch.cuda.set_device
        if self.device not in ("cpu", "cuda"):
            raise ValueError(f"device must be 'cpu' or 'cuda', got {self.device!r}")
        if isinstance(self.audio_config, dict):
            self.audio_config = AudioConfig(**self.audio_config)

    def model_copy(self, deep: bool = False, update: Optional[dict] = None) -> "ModelSpec":
        new = copy.deepcopy(self) if deep else copy.copy(self)
        for k, v in (update or {}).items():
            setattr(new, k, v)
        return new

    @classmethod
    def from_dict(cls, d: dict) -> "ModelSpec":
        known = {f.name for f in fields(cls)} - {"extra"}
        kw = {k: v for k, v in d.items() if k in known}
        return cls(**kw, extra={k: v for k, v in d.items() if k not in known})
