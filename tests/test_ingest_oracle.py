"""Known answers for the ingest restatement (oracle/ingest_oracle.py), derived by hand from
src/bin/birdnet-analyze.rs:21, 684-687 (conversion) and 707-743 (chunk_audio)."""
import numpy as np

from oracle import ingest_oracle as io


def test_conversion_is_exact_and_bounded():                      # birdnet-analyze.rs:684-687
    pcm = np.array([-32768, -1, 0, 1, 16384, 32767], dtype=np.int16)
    f = io.pcm16_to_f32(pcm)
    assert f.dtype == np.float32
    assert f.tolist() == [-1.0, -1 / 32768, 0.0, 1 / 32768, 0.5, 32767 / 32768]


def test_no_overlap_exact_multiple():                            # 3 full segments, no padding
    x = np.arange(12, dtype=np.float32)
    segs = io.chunk_audio(x, 4, 0.0, 2)
    assert [t for t, _ in segs] == [0.0, 2.0, 4.0]
    assert [s.tolist() for _, s in segs] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]]


def test_last_segment_is_zero_padded():                          # birdnet-analyze.rs:731-733
    x = np.arange(1, 11, dtype=np.float32)                       # 10 samples, segment 4
    segs = io.chunk_audio(x, 4, 0.0, 1)
    assert len(segs) == 3
    assert segs[2][1].tolist() == [9, 10, 0, 0]


def test_overlap_step_and_tail_segments():
    # overlap 1.0 s at 2 Hz -> 2 samples, step 2: positions 0,2,4,6,8 (< 10); the last two are padded
    x = np.arange(1, 11, dtype=np.float32)
    segs = io.chunk_audio(x, 4, 1.0, 2)
    assert [t for t, _ in segs] == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert segs[3][1].tolist() == [7, 8, 9, 10]
    assert segs[4][1].tolist() == [9, 10, 0, 0]


def test_overlap_truncates_like_as_usize():                      # (1.4 * 2.0) as usize == 2
    x = np.arange(8, dtype=np.float32)
    assert len(io.chunk_audio(x, 4, 1.4, 2)) == 4                # step 2 -> positions 0,2,4,6


def test_step_zero_and_oversized_overlap_return_nothing():       # saturating_sub, 721-724
    x = np.arange(8, dtype=np.float32)
    assert io.chunk_audio(x, 4, 2.0, 2) == []
    assert io.chunk_audio(x, 4, 100.0, 2) == []


def test_empty_input_gives_no_segments():
    assert io.chunk_audio(np.zeros(0, dtype=np.float32), 4, 0.0, 2) == []


def test_v24_cli_defaults_segment_count():
    # 10 s of 48 kHz audio, 3 s segments, overlap 0: ceil(480000 / 144000) = 4 segments, the last padded
    x = np.ones(480000, dtype=np.float32)
    segs = io.chunk_audio(x, 144000, 0.0, 48000)
    assert len(segs) == 4 and segs[3][0] == 9.0
    assert segs[3][1][:48000].all() and not segs[3][1][48000:].any()
