"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

CPU restatement of the reference CLI's ingest for SURVEY.md section 8(f) row 1:

* ``pcm16_to_f32``  - read_wav's sample conversion, src/bin/birdnet-analyze.rs:21 and 684-687:
  ``f32::from(v) / 32768.0`` (exact in FP32: the divisor is a power of two).
* ``chunk_audio``   - src/bin/birdnet-analyze.rs:707-743: overlap_samples = ``(overlap_secs * sample_rate as f32)
  as usize`` (f32 product, truncation), step = saturating_sub, empty result when step == 0, one segment per
  ``pos`` in ``0, step, 2*step, ... < len``, the last ones zero-padded to ``segment_samples``; start time =
  ``pos as f32 / sample_rate as f32``.

The reference holds no tests for these two functions (its #[cfg(test)] block covers only argument
parsing), so this restatement is pinned by the hand-derived known answers in tests/test_ingest_oracle.py.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

I16_NORMALIZATION_FACTOR = np.float32(32768.0)          # birdnet-analyze.rs:21


def pcm16_to_f32(pcm: np.ndarray) -> np.ndarray:
    return pcm.astype(np.float32) / I16_NORMALIZATION_FACTOR


def chunk_audio(samples: np.ndarray, segment_samples: int, overlap_secs: float,
                sample_rate: int) -> List[Tuple[float, np.ndarray]]:
    overlap_samples = max(int(np.float32(overlap_secs) * np.float32(sample_rate)), 0)
    step = max(segment_samples - overlap_samples, 0)
    if step == 0:
        return []
    out = []
    pos = 0
    n = len(samples)
    while pos < n:
        end = min(pos + segment_samples, n)
        seg = np.zeros(segment_samples, dtype=np.float32)
        seg[:end - pos] = samples[pos:end]
        out.append((float(np.float32(pos) / np.float32(sample_rate)), seg))
        pos += step
    return out
