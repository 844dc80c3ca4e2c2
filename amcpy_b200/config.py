"""Configuration objects with the field names and defaults the extraction stage reads from the
reference's `amcpy.config` (/root/reference/src/amcpy/config.py): Paths (:15-57), SignalConfig
(:60-110), FeatureConfig (:113-148), TrainingConfig (:151-176), Config (:179-186).  Frozen
dataclasses, no file/env input - programmatic override only, like the reference."""

from __future__ import annotations

import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import ClassVar

_SUBDIRS = {
    "mat_data": "mat-data",
    "calculated_features": "calculated-features",
    "arm_data": "arm-data",
    "trained_ann": "ann",
    "figures": "figures",
    "feature_figures": "figures/features",
}
_MODS = ("BPSK", "QPSK", "8PSK", "16QAM", "64QAM")
_MAT_VARS = ("signal_bpsk", "signal_qpsk", "signal_8psk", "signal_qam16", "signal_qam64", "signal_noise")
_LATEX = (r"\gamma_{max}", r"\sigma_{ap}", r"\sigma_{dp}", r"\sigma_{aa}", r"\sigma_{af}", "X", "X_2",
          r"\mu_{42}^{a}", r"\mu_{42}^{f}", "C_{20}", "C_{21}", "C_{40}", "C_{41}", "C_{42}", "C_{60}",
          "C_{61}", "C_{62}", "C_{63}")


@dataclass(frozen=True)
class Paths:
    root: Path = field(default_factory=lambda: Path(os.getcwd()))
    mat_data: Path = field(init=False)
    calculated_features: Path = field(init=False)
    arm_data: Path = field(init=False)
    trained_ann: Path = field(init=False)
    figures: Path = field(init=False)
    feature_figures: Path = field(init=False)
    mat_filename: str = "all_modulations.mat"

    def __post_init__(self) -> None:
        for attr, sub in _SUBDIRS.items():
            object.__setattr__(self, attr, Path(self.root) / sub)

    def ensure_dirs(self) -> None:
        for attr in _SUBDIRS:
            getattr(self, attr).mkdir(parents=True, exist_ok=True)


@dataclass(frozen=True)
class SignalConfig:
    modulations: tuple = _MODS
    modulations_with_noise: tuple = _MODS + ("WGN",)
    labels: tuple = tuple(range(6))
    # index -> SNR in dB as text: -10 .. 20 step 2 (config.py:75-94)
    snr_values: dict = field(default_factory=lambda: {i: str(-10 + 2 * i) for i in range(16)})
    frame_size: int = 2048
    num_frames: int = 1000
    num_threads: int = 8  # kept for signature compatibility; the GPU path has no thread pool
    mat_info: dict = field(default_factory=lambda: dict(zip(_MODS + ("WGN",), _MAT_VARS)))


@dataclass(frozen=True)
class FeatureConfig:
    names: ClassVar[dict] = {i + 1: f"${s}$" for i, s in enumerate(_LATEX)}
    all_features: tuple = tuple(range(1, 19))
    used: tuple = (2, 4, 6, 8, 12, 14)

    @property
    def used_names(self) -> list:
        return [self.names[f] for f in self.used]

    @property
    def num_used(self) -> int:
        return len(self.used)


@dataclass(frozen=True)
class TrainingConfig:
    training_snr: tuple = (10, 11, 12, 13, 14, 15)
    all_snr: tuple = tuple(range(16))
    plotting_snr: tuple = tuple(range(16))
    test_size: float = 0.2
    random_state: int = 42
    activation: str = "relu"
    batch_size: int = 128
    dropout: float = 0.4
    epochs: int = 21
    learning_rate: float = 0.001418378071933655
    optimizer: str = "rmsprop"
    layer_size_hl1: int = 26
    layer_size_hl2: int = 29
    layer_size_hl3: int = 30

    @property
    def feature_files(self) -> list:
        return [f"{m}_features" for m in SignalConfig().modulations_with_noise]


@dataclass(frozen=True)
class Config:
    paths: Paths = field(default_factory=Paths)
    signals: SignalConfig = field(default_factory=SignalConfig)
    features: FeatureConfig = field(default_factory=FeatureConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)
