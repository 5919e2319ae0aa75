"""TEST INFRASTRUCTURE ONLY.  CPU restatement of RangeFilter::predict (src/rangefilter.rs:435-502) on the
build-authored meta model (birdnet_b200/modelgen/meta_model.py): week = calculate_week (77-81), input
[lat, lon, week], one forward pass, keep `score >= threshold && i < labels.len()` (482-496), sort by score
descending with total_cmp (499).  The reference ships no meta-model file and no test holds an output of it
("parity unpinned" for the MLP arithmetic, like the classifier graphs); the selection / ordering semantics are the
reference's."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def calculate_week(month: int, day: int) -> float:                       # rangefilter.rs:77-81
    return float((month - 1) * 4 + (day - 1) // 7 + 1)            # integer arithmetic, then `as f32`


def forward(weights: Dict[str, np.ndarray], lat: float, lon: float, week: float) -> np.ndarray:
    x = np.array([lat, lon, week], dtype=np.float32) * weights["in_scale"]
    n = sum(1 for k in weights if k.endswith(".weight"))
    for i in range(n):
        x = (weights[f"fc{i}.weight"].astype(np.float64) @ x.astype(np.float64)).astype(np.float32) + weights[f"fc{i}.bias"]
        if i < n - 1:
            x = np.maximum(x, np.float32(0))
    return (np.float32(1) / (np.float32(1) + np.exp(-x.astype(np.float32)))).astype(np.float32)


def predict(weights, labels: List[str], threshold: float, lat, lon, month, day) -> List[Tuple[str, float, int]]:
    s = forward(weights, lat, lon, calculate_week(month, day))
    out = [(labels[i], float(s[i]), i) for i in range(len(s)) if s[i] >= np.float32(threshold) and i < len(labels)]
    out.sort(key=lambda t: -t[1])
    return out
