"""Public types of the reference API (src/types.rs)."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


class LabelFormat(enum.Enum):            # types.rs:60-68
    Text = "text"
    Csv = "csv"
    Json = "json"


class ModelType(enum.Enum):              # types.rs:3-57
    BirdNetV24 = 0
    BirdNetV30 = 1
    PerchV2 = 2

    def sample_rate(self) -> int:
        return 48_000 if self is ModelType.BirdNetV24 else 32_000

    def segment_duration(self) -> float:
        return 3.0 if self is ModelType.BirdNetV24 else 5.0

    def sample_count(self) -> int:
        return 144_000 if self is ModelType.BirdNetV24 else 160_000

    def has_embeddings(self) -> bool:
        return self is not ModelType.BirdNetV24

    def expected_label_format(self) -> LabelFormat:
        return LabelFormat.Text if self is ModelType.BirdNetV24 else LabelFormat.Csv


@dataclass
class ModelConfig:                       # types.rs:72-85
    model_type: ModelType
    sample_rate: int
    segment_duration: float
    sample_count: int
    num_species: int
    embedding_dim: Optional[int]


@dataclass(slots=True)
class Prediction:                        # types.rs:89-96
    species: str
    confidence: float
    index: int


@dataclass(slots=True)
class PredictionResult:                  # types.rs:100-109
    model_type: ModelType
    predictions: List[Prediction]
    embeddings: Optional[np.ndarray]
    raw_scores: np.ndarray


@dataclass(slots=True)
class LocationScore:                     # types.rs:113-120
    species: str
    score: float
    index: int


class ExecutionProviderInfo(enum.Enum):  # types.rs:124-185 (only the providers that exist here)
    Cpu = "CPU"
    Cuda = "CUDA"
    B200 = "B200"                        # this engine: hand-written sm_100a kernels

    def as_str(self) -> str:
        return self.value

    def category(self) -> str:
        return "CPU" if self is ExecutionProviderInfo.Cpu else "GPU"

    def __str__(self) -> str:
        return self.value


def available_execution_providers() -> List[ExecutionProviderInfo]:
    """execution_providers.rs:35-58 probes ORT EPs; here: the B200 engine when a device exists.
    CPU is listed first as in the reference (tests/execution_provider_test.rs:31-47) but is
    informational only: this package has no CPU compute path."""
    from . import _ffi
    out = [ExecutionProviderInfo.Cpu]
    if _ffi.lib.bn_device_count() > 0:
        out.append(ExecutionProviderInfo.B200)
    return out
