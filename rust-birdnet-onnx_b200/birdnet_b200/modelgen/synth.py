"""Seeded synthetic audio (SURVEY.md section 8d "Synthetic audio distribution").

The reference's own synthetic inputs are silence (tests/integration_test.rs:52-54,
src/testutil.rs:51-53), sines at amplitude 0.5 (integration_test.rs:57-67; 440 Hz and 1 kHz are
the frequencies it uses) and amplitude 1.0 (testutil.rs:57-67), and an LCG
(testutil.rs:110-121).  The other kinds (chirp, tone mixtures, amplitude-scaled variants)
exercise the min/max normaliser of the v2.4 front-end.
"""
from __future__ import annotations

import numpy as np

LCG_A = np.uint64(1_103_515_245)
LCG_C = np.uint64(12345)


def lcg_bits(count: int, seed: int) -> np.ndarray:
    """16-bit outputs of the reference LCG (testutil.rs:113-118), vectorised.

    state_n = a^n * s0 + c * sum_{i<n} a^i   (mod 2^64)
    """
    if count == 0:
        return np.zeros(0, dtype=np.uint64)
    with np.errstate(over="ignore"):
        a_pow = np.multiply.accumulate(np.full(count, LCG_A, dtype=np.uint64))        # a^1..a^n
        geo = np.concatenate([np.ones(1, dtype=np.uint64), a_pow[:-1]])               # a^0..a^(n-1)
        geo_sum = np.add.accumulate(geo)                                              # sum_{i<n} a^i
        state = a_pow * np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + LCG_C * geo_sum
    return (state >> np.uint64(16)) & np.uint64(0xFFFF)


def random_logits(count: int, seed: int) -> np.ndarray:
    """testutil.rs:110-121: bits.mul_add(10/65535, -5) in f32."""
    bits = lcg_bits(count, seed).astype(np.float32)
    # f32::mul_add is a fused multiply-add: compute in f64 (exact product) then round once
    return (bits.astype(np.float64) * np.float64(np.float32(10.0) / np.float32(65535.0)) - 5.0).astype(np.float32)


def mock_embeddings(dim: int, seed: int) -> np.ndarray:
    """testutil.rs:137-147: bits / 65535 in f32."""
    return (lcg_bits(dim, seed).astype(np.float32) / np.float32(65535.0)).astype(np.float32)


def silence(n: int) -> np.ndarray:
    return np.zeros(n, dtype=np.float32)


def sine(n: int, sample_rate: int, freq: float, amp: float = 1.0) -> np.ndarray:
    """f32 arithmetic like the reference: t = i / sr; sin(2*pi*f*t) * amp."""
    t = np.arange(n, dtype=np.float32) / np.float32(sample_rate)
    x = np.sin((np.float32(2.0) * np.float32(np.pi) * np.float32(freq)) * t, dtype=np.float32)
    return (x * np.float32(amp)).astype(np.float32)


def lcg_noise(n: int, seed: int) -> np.ndarray:
    bits = lcg_bits(n, seed).astype(np.float32)
    return (bits * np.float32(2.0 / 65535.0) - np.float32(1.0)).astype(np.float32)


def chirp(n: int, sample_rate: int, f0: float = 100.0, f1: float = 15000.0,
          amp: float = 0.8) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sample_rate
    dur = n / sample_rate
    phase = 2.0 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return (amp * np.sin(phase)).astype(np.float32)


def noise_tones(n: int, sample_rate: int, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    x = 0.1 * lcg_noise(n, seed + 7).astype(np.float64)
    t = np.arange(n, dtype=np.float64) / sample_rate
    for _ in range(3):
        f = rng.uniform(200.0, min(12000.0, 0.45 * sample_rate))
        a = rng.uniform(0.1, 0.5)
        x += a * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    return x.astype(np.float32)


_KINDS = 10


def segment(index: int, n: int, sample_rate: int, seed: int = 0) -> np.ndarray:
    """Deterministic segment #index of the round-robin test distribution."""
    k = index % _KINDS
    s = seed * 1_000_003 + index
    if k == 0:
        return lcg_noise(n, s)
    if k == 1:
        return sine(n, sample_rate, 440.0, 0.5)
    if k == 2:
        return noise_tones(n, sample_rate, s)
    if k == 3:
        return sine(n, sample_rate, 1000.0, 1.0)
    if k == 4:
        return chirp(n, sample_rate, 100.0, min(15000.0, 0.45 * sample_rate))
    if k == 5:
        return (lcg_noise(n, s) * np.float32(1e-3)).astype(np.float32)
    if k == 6:
        return sine(n, sample_rate, 1000.0, 0.5)
    if k == 7:
        return (noise_tones(n, sample_rate, s) * np.float32(1e-3) + np.float32(0.25)).astype(np.float32)
    if k == 8:
        return sine(n, sample_rate, 440.0, 1.0)
    return silence(n)


def batch(start: int, count: int, n: int, sample_rate: int, seed: int = 0) -> np.ndarray:
    out = np.empty((count, n), dtype=np.float32)
    for i in range(count):
        out[i] = segment(start + i, n, sample_rate, seed)
    return out
