"""birdnet-analyze clone (birdnet_b200/cli.py): the reference CLI's helper arithmetic on CPU, the full run on a GPU."""
import io
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

from birdnet_b200 import cli

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_wav(path, pcm: np.ndarray, rate: int, channels=1, bits=16, tag=1):
    data = pcm.astype("<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, channels, rate, rate * channels * bits // 8, channels * bits // 8, bits))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def test_format_helpers():                      # birdnet-analyze.rs:745-779
    assert cli.format_time(0.0) == "00:00.0" and cli.format_time(3.0) == "00:03.0" and cli.format_time(75.5) == "01:15.5"
    assert cli.format_time(3599.96) in ("59:60.0", "59:59.9", "59:60.0")     # same f32-free arithmetic as `{secs_part:04.1}`
    assert cli.format_duration(45.0) == "45s" and cli.format_duration(203.9) == "3m 23s" and cli.format_duration(4530.0) == "1h 15m 30s"


def test_chunk_audio_matches_the_ingest_oracle():
    from oracle import ingest_oracle as io_
    x = (np.arange(10_000, dtype=np.float32) / 10_000.0)
    for seg, ov, sr in [(3000, 0.0, 1000), (3000, 1.0, 1000), (3000, 2.5, 1000), (144000, 0.0, 48000)]:
        got = cli.chunk_audio(x, seg, ov, sr)
        ref = io_.chunk_audio(x, seg, ov, sr)
        assert len(got) == len(ref)
        for (t0, a), (t1, b) in zip(got, ref):
            assert t0 == t1 and np.array_equal(a, b)
    assert cli.chunk_audio(x, 3000, 3.0, 1000) == []           # step 0 -> no segments (saturating_sub)


def test_read_wav_checks(tmp_path):
    import birdnet_b200 as bb
    p = str(tmp_path / "a.wav")
    pcm = (np.sin(np.arange(4800) * 0.1) * 12000).astype(np.int16)
    _write_wav(p, pcm, 48000)
    got, rate = cli.read_wav_pcm16(p)
    assert rate == 48000 and np.array_equal(got, pcm)
    _write_wav(p, np.stack([pcm, pcm], 1).reshape(-1), 48000, channels=2)
    with pytest.raises(bb.AudioFormat) as e:
        cli.read_wav_pcm16(p)
    assert str(e.value) == "unsupported audio format: WAV must be mono (1 channel), got 2 channels"      # error.rs AudioFormat
    _write_wav(p, pcm, 48000, bits=8)
    with pytest.raises(bb.AudioFormat) as e:
        cli.read_wav_pcm16(p)
    assert "WAV must be 16-bit, got 8-bit" in str(e.value)
    _write_wav(p, pcm[:0], 48000)
    with pytest.raises(bb.AudioFormat) as e:
        cli.read_wav_pcm16(p)
    assert "WAV file has no samples" in str(e.value)
    with pytest.raises(bb.AudioRead):
        cli.read_wav_pcm16(str(tmp_path / "missing.wav"))


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path, v24_model_path, v24_spec):
    from birdnet_b200.modelgen import synth
    from birdnet_b200.modelgen.make_models import synthetic_labels
    audio = synth.batch(0, 10, 144000, 48000).reshape(-1)
    pcm = (np.clip(audio, -1, 1) * 32767).astype(np.int16)[: 144000 * 9 + 50000]      # ragged tail -> zero padded segment
    wav, labels = str(tmp_path / "rec.wav"), str(tmp_path / "labels.txt")
    _write_wav(wav, pcm, 48000)
    with open(labels, "w") as f:
        f.write("\n".join(synthetic_labels(v24_spec.num_species)) + "\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "rust-birdnet-onnx_b200"))
    outs = []
    for extra in ([], ["--pcm-ingest"]):
        r = subprocess.run([sys.executable, "-m", "birdnet_b200.cli", wav, "--model", v24_model_path, "--labels", labels,
                            "-b", "4", "-k", "3", "--min-confidence", "0.05", "-t", "30"] + extra,
                           capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr
        lines = r.stdout.splitlines()
        assert lines[0] == "Using execution provider: B200" and lines[1].startswith("Batch size: 4")
        assert lines[2] == f"Analyzing: {wav} (28s, 48000 Hz)" and lines[3] == "Model: BirdNET v2.4 (3.0s segments, 0.0s overlap)"
        assert lines[-1].startswith("10 segments of 28s audio analyzed in ") and "segments/s" in lines[-1]
        outs.append([ln for ln in lines[5:-2]])
    assert outs[0] == outs[1] and len(outs[0]) >= 1               # host chunking == on-device ingest, line for line
    assert all(ln[2] == ":" and "%" in ln for ln in outs[0])
    r = subprocess.run([sys.executable, "-m", "birdnet_b200.cli", wav, "--model", v24_model_path, "--labels", labels, "-o", "3.0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 1 and "overlap (3.0s) must be less than segment duration (3.0s)" in r.stderr
