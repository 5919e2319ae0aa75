"""`birdnet-analyze` for the B200 engine: the reference CLI (src/bin/birdnet-analyze.rs) restated over this package.

    python -m birdnet_b200.cli recording.wav --model m.onnx --labels labels.txt [-o 0.0] [-k 3]
           [--min-confidence 0.1] [--model-type v24|v30|perch] [-b 32] [-t 1] [-v] [--provider b200]
           [--pcm-ingest]

Same arguments, defaults, validation messages, output lines and end-of-run report as the reference
(`Args` 45-98, main loop 500-647, `read_wav` 650-704, `chunk_audio` 707-743, `format_time` /
`format_duration` 745-779).  Differences, all forced by the engine:
  * providers: `cpu` is refused (there is no CPU compute path); `cuda` / `b200` select this engine;
  * `--pcm-ingest` (not in the reference) hands the 16-bit PCM to the device as it is: conversion and
    chunking run on the GPU (`Classifier.predict_pcm16_stream`);
  * Ctrl-C: the first one cancels through the CancellationToken (the batch in flight returns `Cancelled`),
    the second one exits with status 1 - as in the reference (541-556).
"""
from __future__ import annotations

import argparse
import datetime
import signal
import struct
import sys
import time
from typing import List, Optional, Tuple

import numpy as np

I16_NORMALIZATION_FACTOR = np.float32(32768.0)     # birdnet-analyze.rs:21
DEFAULT_GPU_BATCH_SIZE = 32                        # birdnet-analyze.rs:42
PROVIDERS = ["cpu", "cuda", "b200"]


def timestamp() -> str:
    return datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.%fZ")


def format_time(secs: float) -> str:
    """MM:SS.d (birdnet-analyze.rs:751-756)."""
    total = int(secs)
    mins = total // 60
    return f"{mins:02d}:{secs - mins * 60:04.1f}"


def format_duration(secs: float) -> str:
    """birdnet-analyze.rs:759-774."""
    total = int(secs)
    h, m, s = total // 3600, (total % 3600) // 60, total % 60
    if h > 0:
        return f"{h}h {m}m {s}s"
    if m > 0:
        return f"{m}m {s}s"
    return f"{s}s"


def read_wav_pcm16(path: str) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE reader with the reference's checks (650-684): mono, 16-bit, integer PCM.  Returns (int16 samples, rate)."""
    from .errors import AudioFormat, AudioRead
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise AudioRead(path, str(e))
    if len(data) < 12 or data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise AudioRead(path, "no RIFF tag found")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            if len(body) < 16:
                raise AudioRead(path, "invalid fmt chunk")
            fmt = struct.unpack_from("<HHIIHH", body, 0)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise AudioRead(path, "missing fmt or data chunk")
    tag, channels, rate, _, _, bits = fmt
    if tag == 0xFFFE and len(data) > 0:              # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
        tag = 1
    if channels != 1:
        raise AudioFormat(f"WAV must be mono (1 channel), got {channels} channels")
    if bits != 16:
        raise AudioFormat(f"WAV must be 16-bit, got {bits}-bit")
    if tag != 1:
        raise AudioFormat("WAV must be integer format, not float")
    samples = np.frombuffer(pcm[: len(pcm) // 2 * 2], dtype="<i2")
    if samples.size == 0:
        raise AudioFormat("WAV file has no samples")
    return samples, int(rate)


def chunk_audio(samples: np.ndarray, segment_samples: int, overlap_secs: float, sample_rate: int) -> List[Tuple[float, np.ndarray]]:
    """birdnet-analyze.rs:707-743: (start time in f32 arithmetic, zero-padded segment)."""
    overlap = int(np.float32(overlap_secs) * np.float32(sample_rate))
    step = max(segment_samples - max(overlap, 0), 0)
    if step == 0:
        return []
    out, pos = [], 0
    while pos < len(samples):
        seg = samples[pos:pos + segment_samples]
        if len(seg) < segment_samples:
            seg = np.concatenate([seg, np.zeros(segment_samples - len(seg), dtype=np.float32)])
        out.append((float(np.float32(pos) / np.float32(sample_rate)), seg))
        pos += step
    return out


def parse_model_type(arg: Optional[str]):
    from .errors import ModelDetection
    from .types import ModelType
    if arg is None:
        return None
    table = {"v24": ModelType.BirdNetV24, "v30": ModelType.BirdNetV30, "perch": ModelType.PerchV2}
    if arg not in table:
        raise ModelDetection(f"unknown model type '{arg}', expected: v24, v30, perch")
    return table[arg]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="birdnet-analyze", description="Analyze WAV files for bird species")
    ap.add_argument("audio_file", nargs="?")
    ap.add_argument("-m", "--model")
    ap.add_argument("-l", "--labels")
    ap.add_argument("-o", "--overlap", type=float, default=0.0)
    ap.add_argument("-k", "--top-k", type=int, default=3)
    ap.add_argument("--min-confidence", type=float, default=0.1)
    ap.add_argument("--model-type")
    ap.add_argument("--list-providers", action="store_true")
    ap.add_argument("--provider", default="b200")
    ap.add_argument("-b", "--batch-size", type=int)
    ap.add_argument("-t", "--timeout", type=int, default=1)
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--pcm-ingest", action="store_true", help="upload 16-bit PCM once; convert and chunk on the GPU")
    args = ap.parse_args(argv)

    import birdnet_b200 as bb
    from .errors import Cancelled, Error, ModelDetection, AudioFormat
    if args.list_providers:
        print("Available execution providers:")
        avail = bb.available_execution_providers()
        for p in (bb.ExecutionProviderInfo.Cpu, bb.ExecutionProviderInfo.B200):
            mark = "x" if p in avail and p is not bb.ExecutionProviderInfo.Cpu else " "
            note = "listed for parity with the reference; no compute path in this engine" if p is bb.ExecutionProviderInfo.Cpu \
                else "hand-written sm_100a kernels"
            print(f"  [{mark}] {p.as_str():<10} {note}")
        return 0
    try:
        if not args.audio_file or not args.model or not args.labels:
            ap.error("the following arguments are required: audio_file, --model, --labels")
        prov = args.provider.lower()
        if prov not in PROVIDERS:
            raise ModelDetection(f"unknown provider '{args.provider}'. Valid providers: {', '.join(PROVIDERS)}")
        if prov == "cpu":
            raise bb.RuntimeInit("this engine has no CPU execution provider; use --provider b200")
        batch_size = args.batch_size or DEFAULT_GPU_BATCH_SIZE
        t0 = time.perf_counter()
        builder = (bb.Classifier.builder().model_path(args.model).labels_path(args.labels)
                   .top_k(args.top_k).min_confidence(args.min_confidence))
        mt = parse_model_type(args.model_type)
        if mt is not None:
            builder = builder.model_type(mt)
        clf = builder.build()
        if args.verbose:
            print(f"{timestamp()} [DEBUG] Classifier built in {time.perf_counter() - t0:.3f}s", file=sys.stderr)
        cfg = clf.config()
        pcm, rate = read_wav_pcm16(args.audio_file)
        duration = float(np.float32(len(pcm)) / np.float32(rate))
        if rate != cfg.sample_rate:
            raise AudioFormat(f"model expects {cfg.sample_rate} Hz audio, WAV is {rate} Hz")
        if args.overlap >= cfg.segment_duration:
            raise ModelDetection(f"overlap ({args.overlap:.1f}s) must be less than segment duration ({cfg.segment_duration:.1f}s)")
        ctx = None
        try:
            ctx = clf.create_batch_context(batch_size, allow_perch=args.pcm_ingest)
            if args.verbose:
                print(f"{timestamp()} [DEBUG] Created batch context (max_batch_size={batch_size}, "
                      f"input_buffer={ctx.input_buffer_bytes() / 1048576.0:.1f}MB pre-allocated)", file=sys.stderr)
        except Error as e:                     # PerchV2: reference falls back to predict_batch (464-496)
            if args.verbose:
                print(f"{timestamp()} [DEBUG] Batch context not available: {e}, using standard batch inference", file=sys.stderr)
        names = {bb.ModelType.BirdNetV24: "BirdNET v2.4", bb.ModelType.BirdNetV30: "BirdNET v3.0", bb.ModelType.PerchV2: "Perch v2"}
        print("Using execution provider: B200")
        print(f"Batch size: {batch_size} (batch context enabled)" if ctx is not None else f"Batch size: {batch_size}")
        print(f"Analyzing: {args.audio_file} ({format_duration(duration)}, {rate} Hz)")
        print(f"Model: {names[cfg.model_type]} ({cfg.segment_duration:.1f}s segments, {args.overlap:.1f}s overlap)")
        print()

        token = bb.CancellationToken()
        state = {"cancelled": False}

        def on_sigint(signum, frame):          # birdnet-analyze.rs:547-556
            if state["cancelled"]:
                print("\nForce exiting...", file=sys.stderr)
                sys.exit(1)
            state["cancelled"] = True
            print("\nCancelling... (press Ctrl+C again to force exit)", file=sys.stderr)
            token.cancel()
        try:
            signal.signal(signal.SIGINT, on_sigint)
        except ValueError:
            pass                               # not the main thread: no handler, like `.ok()` on a second registration
        opts = (bb.InferenceOptions.with_only_timeout(float(args.timeout)) if args.timeout > 0 else bb.InferenceOptions.new())
        opts = opts.with_cancellation_token(token)

        def show(offset: float, result) -> None:
            if result.predictions:
                print(f"{format_time(offset)}  " + ", ".join(f"{p.species} ({p.confidence * 100.0:.1f}%)" for p in result.predictions))

        start = time.perf_counter()
        if args.pcm_ingest and ctx is not None:
            S = cfg.sample_count
            overlap = int(np.float32(args.overlap) * np.float32(rate))
            step = max(S - overlap, 0)
            n_seg = 0 if step == 0 else (len(pcm) + step - 1) // step
            try:
                for off, res in clf.predict_pcm16_stream(ctx, pcm, args.overlap, opts):
                    show(off, res)
            except Cancelled:
                print("Processing cancelled by user", file=sys.stderr)
        else:
            samples = pcm.astype(np.float32) / I16_NORMALIZATION_FACTOR
            segments = chunk_audio(samples, cfg.sample_count, args.overlap, rate)
            n_seg = len(segments)
            total_batches = -(-n_seg // batch_size)
            for bi in range(total_batches):
                if state["cancelled"]:
                    print("Processing cancelled by user", file=sys.stderr)
                    break
                chunk = segments[bi * batch_size:(bi + 1) * batch_size]
                if args.verbose:
                    print(f"{timestamp()} [DEBUG] Processing batch {bi + 1}/{total_batches} ({len(chunk)} segments)...", file=sys.stderr)
                tb = time.perf_counter()
                segs = [s for _, s in chunk]
                try:
                    results = clf.predict_batch_with_context(ctx, segs, opts) if ctx is not None else clf.predict_batch(segs, opts)
                except Cancelled:
                    print("Processing cancelled by user", file=sys.stderr)
                    break
                if args.verbose:
                    print(f"{timestamp()} [DEBUG] Batch {bi + 1} completed in {time.perf_counter() - tb:.4f}s", file=sys.stderr)
                for (off, _), res in zip(chunk, results):
                    show(off, res)
        elapsed = max(time.perf_counter() - start, 1e-9)
        print()
        print(f"{n_seg} segments of {format_duration(duration)} audio analyzed in {elapsed:.1f}s "
              f"({n_seg / elapsed:.1f} segments/s, {duration / elapsed:.1f}x realtime)")
        return 0
    except Error as e:
        print(f"Error: {e}", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
