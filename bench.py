#!/usr/bin/env python
"""Benchmark of the hot path: BirdNET v2.4 segments/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA engine)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the
                                                             # reference's path on the host cores

A step = one pass of the whole hot path over one batch of 256 synthetic 48 kHz segments
(BASELINE.json configs[1]: create_batch_context(256) + predict_batch_with_context).
`--config 1|3|4|5` runs the other BASELINE.json configurations with the same line schema
(1: predict_batch B=32; 3: v3.0 B=512; 4: Perch v2 B=256; 5: 28,800 segments + range filter,
strong scaling over the ranks).

  value : device-timed throughput with the batch already resident in HBM (front-end + CNN + fused
          top-k epilogue + D2H of the results), K steps enqueued back to back, CUDA events on the
          engine's own stream, max over ranks.
  e2e   : the same metric through the public API (Classifier.predict_batch_with_context) with
          HOST segment slices: host gather into pinned staging, H2D, kernels, D2H and the
          host-side Prediction objects are all inside the timed region.
  roofline : the dominant kernel of the step, timed live with CUDA events (profiling mode of the
          engine) against MEASURED_PEAKS.json.
  cpu_baseline : oracle port (torch CPU FP32 stand-in for ORT CPU) on a
          bounded sample, rank 0 at N=1 only.

Multi-GPU: one process per GPU (torchrun); segments shard across ranks with no data-path
collective; NCCL is used only for the barrier and the max-over-ranks time (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = "segments/s"

# BASELINE.json configs[0..4] -> --config 1..5.  The driver's default (no flag) is --config 2, the configuration the
# headline metric is quoted on; the others are run by hand (profiles/r02_bench_cfg*.json) with the same line schema.
CONFIGS = {
    1: dict(family="birdnet_v24", batch=32, api="predict_batch", metric="birdnet_v2.4_segments_per_sec",
            workload="BirdNET v2.4-like (random-init seed 0, build-authored graph) predict_batch batch=32, top_k=5 "
                     "min_conf=0.1 (BASELINE.json configs[0])"),
    2: dict(family="birdnet_v24", batch=256, api="ctx", metric="birdnet_v2.4_segments_per_sec",
            workload="BirdNET v2.4-like (random-init seed 0, build-authored graph) batch=256 via "
                     "BatchInferenceContext: front-end + CNN + fused top-k epilogue (BASELINE.json configs[1])"),
    3: dict(family="birdnet_v30", batch=512, api="ctx", metric="birdnet_v3.0_segments_per_sec",
            workload="BirdNET v3.0-like (32 kHz, 160,000 samples, log-mel, 1024-d embedding + 11,560 logits) batch=512 via "
                     "BatchInferenceContext (BASELINE.json configs[2])"),
    4: dict(family="perch_v2", batch=256, api="predict_batch", metric="perch_v2_segments_per_sec",
            workload="Perch-v2-like (32 kHz, 160,000 samples, log-mel 500x128, 1536-d embedding + 14,795 logits) batch=256 "
                     "via predict_batch - the reference's context path refuses Perch (BASELINE.json configs[3])"),
    5: dict(family="birdnet_v24", batch=256, api="ctx", metric="birdnet_v2.4_segments_per_sec", total_segments=28800,
            workload="24 h of 48 kHz audio = 28,800 BirdNET v2.4 segments, sharded over the GPUs in batches of 256, fused "
                     "range filter (rerank on) + top_k=5 min_conf=0.1 (BASELINE.json configs[4]); strong scaling"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), sm_hz=float(d.get("sm_max_mhz", 1965.0)) * 1e6,
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _ncu_traffic():
    """{stage: DRAM bytes per launch} from the committed ncu --set full capture (profiles/*_ncu_traffic.json,
    written by tools/ncu_summarize.py); the most recent file by name wins."""
    d = os.path.join(ROOT, "profiles")
    try:
        names = sorted(n for n in os.listdir(d) if n.endswith("_ncu_traffic.json"))
        if not names:
            return {}, None
        with open(os.path.join(d, names[-1])) as f:
            return json.load(f).get("traffic", {}), names[-1]
    except (OSError, ValueError):
        return {}, None


def _bind_near_gpu(gpu_index: int):
    """Keep this rank's threads (and therefore the page-locked buffers they allocate and fill) on the NUMA node the GPU
    hangs off: on a two-socket box half of the ranks would otherwise stage their input through the other socket.
    Returns the node, or None when the box has one node / the information is missing; never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        with open(f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        cpus = set()
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 4 or len(cpus) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores
# ----------------------------------------------------------------------------------------------
def _cpu_oracle_rate(n_target_s: float, batch: int, max_segments: int, family: str = "birdnet_v24"):
    import torch
    from birdnet_b200.modelgen import get_spec, synth      # modelgen never maps the product .so (lazy package root)
    from birdnet_b200.modelgen.make_models import ensure_model
    from oracle.model_oracle import ModelOracle, load_initializers
    from oracle import postprocess_oracle as po
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = get_spec(family)
    orc = ModelOracle(spec, load_initializers(ensure_model(family)))
    fe = spec.frontend
    audio = synth.batch(0, batch, fe.sample_count, fe.sample_rate)

    def step():
        logits, _ = orc.logits_and_embeddings(audio)
        po.top_k_batch(logits, 5, 0.1)
    step()                                   # warm-up (thread pools, FFT plans)
    t0 = time.perf_counter()
    step()
    one = time.perf_counter() - t0
    n_steps = int(max(1, min(max_segments // batch, n_target_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        step()
    dt = time.perf_counter() - t0
    return (n_steps * batch) / dt, cores, n_steps * batch, dt, step


def run_reference(args):
    """CPU arm: the oracle port of the reference's path (front-end + CNN in torch CPU FP32 + the C restatement of
    postprocess.rs) on all host cores.  A step is a bounded sample of the workload - one batch of 32 segments, the
    reference's own CPU-runnable batch (BASELINE.json configs[0]); the metric is a per-segment rate, so the sample size
    does not change its meaning (the GPU arm's step is 256 segments: stated in `config`).  Timed per step, at least
    10 s in total, median over steps (round-1 runs of 20 steps spread 309-496 segments/s)."""
    rank, local, world = _dist()
    if rank != 0:
        return 0
    cfg = CONFIGS[args.config]
    batch = 32
    rate, cores, nseg, dt, step = _cpu_oracle_rate(2.0, batch, 64, cfg["family"])   # builds the oracle + a short calibration
    for _ in range(max(args.warmup, 0)):
        step()
    per_step = []
    t_all = time.perf_counter()
    while len(per_step) < args.steps or (time.perf_counter() - t_all) < 10.0:
        t0 = time.perf_counter()
        step()
        per_step.append(time.perf_counter() - t0)
        if len(per_step) >= 400:
            break
    med = float(np.median(per_step))
    value = batch / med
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(per_step), "warmup": args.warmup, "ms_per_step": 1e3 * med,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"] + "; this arm: bounded sample, batch 32 per step on the host cores "
                               "(per-segment rate; the GPU arm's step is %d segments)" % cfg["batch"],
                   "global_batch": batch, "top_k": 5, "min_confidence": 0.1,
                   "timing": "median of %d per-step times over %.1f s (min %.0f, max %.0f segments/s)"
                             % (len(per_step), sum(per_step), batch / max(per_step), batch / min(per_step))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(per_step)} steps x {batch} segments, torch {__import__('torch').__version__} CPU FP32 "
                                   "oracle port (ORT CPU cannot run in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------
# this build
# ----------------------------------------------------------------------------------------------
def _kernel_class(name: str) -> str:
    """Stage name -> the kernel (class) that runs it; `kernel_classes` sums every launch of one class."""
    if name == "normalize":
        return "front-end: min/max + normalise (k_minmax_partial, k_normalize*)"
    if name.startswith("spectrogram"):
        return "front-end: normalise + both spectrogram GEMMs (k_spec_v24; k_tc_conv with BN_DISABLE_FE_FUSED=1)"
    if name == "logmel":
        return "front-end: log-mel (k_logmel)"
    if name.startswith("stem"):
        return "stem conv (k_stem_planes)"
    if ".mbconv" in name:
        return "fused MBConv block (k_mbconv)"
    if ".dw." in name:
        return "depthwise + squeeze-excite (k_dw_se)"
    if ".fused." in name:
        return "3x3 conv (k_tc_conv, halo / gather producers)"
    if ".expand." in name or ".project" in name or name.startswith("head_conv"):
        return "1x1 conv (k_tc_conv, TMA producer)"
    if name.startswith("gap"):
        return "global pool (k_gap_planes)"
    if name == "topk_epilogue":
        return "top-k / sigmoid / range mask (k_topk_small)"
    if name == "d2h":
        return "result fetch (D2H)"
    return "dense head (k_tc_conv)"


def _stage_roofline(spec, stage_times, batch, peaks):
    """Per-stage algorithmic work -> achieved rate; returns (dominant stage dict, all stages)."""
    rows = {r["name"]: r for r in spec.layer_table()}
    fe = spec.frontend
    out = []
    total = sum(ms for _, ms in stage_times) or 1.0
    for name, ms in stage_times:
        if ms <= 0:
            continue
        base = name[:-5] if name.endswith(".conv") else name
        if "_" in base and base.rsplit("_", 1)[1].isdigit() and base.rsplit("_", 1)[0] in rows:
            base = base.rsplit("_", 1)[0]
        d = {"stage": name, "ms": ms, "share": ms / total}
        if name.endswith(".mbconv"):
            # fused MBConv block: algorithmic bytes = block input read once + block output written once (+ weights);
            # the expanded tensors never leave the SM.  FLOPs = expand + depthwise + project.
            blk = name[:-len(".mbconv")]
            parts = [rows[k] for k in (blk + ".expand", blk + ".dw", blk + ".project") if k in rows]
            if len(parts) == 3:
                # Four ceilings for a fused block; the one that takes longest at peak is its roofline:
                #   HBM     block input read once + block output written once (+ weights)
                #   tensor  expand + projection FLOPs at the sustained dense rate
                #   FP32    depthwise MACs on the CUDA cores: 128 FMA lanes x SMs x clock
                #   MUFU    two SiLUs per expanded element (ex2 + rcp each): 16 special-function lanes x SMs x clock
                byts = 4.0 * ((parts[0]["in_elems"] + parts[2]["out_elems"]) * batch + sum(r["w_elems"] for r in parts))
                tflops = 2.0 * (parts[0]["macs"] + parts[2]["macs"]) * batch
                fma = float(parts[1]["macs"]) * batch
                mufu = 2.0 * (parts[0]["out_elems"] + parts[1]["out_elems"]) * batch
                clk = peaks.get("sm_hz", 1.965e9)
                t = {"hbm": byts / (peaks["hbm"] * 1e9), "tensor": tflops / (peaks["tf_sust"] * 1e12),
                     "fp32-fma": fma / (148 * 128 * clk), "mufu": mufu / (148 * 16 * clk)}
                b = max(t, key=t.get)
                if b == "hbm":
                    d.update(bound="hbm", achieved=byts / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s")
                elif b == "tensor":
                    d.update(bound="tensor", achieved=tflops / (ms * 1e-3) / 1e12, peak=peaks["tf_sust"], unit="TFLOP/s")
                elif b == "fp32-fma":
                    d.update(bound="fp32-fma", achieved=fma / (ms * 1e-3) / 1e12, peak=148 * 128 * clk / 1e12, unit="TFMA/s")
                else:
                    d.update(bound="mufu", achieved=mufu / (ms * 1e-3) / 1e12, peak=148 * 16 * clk / 1e12, unit="Tops/s")
                d.update(alg_per_segment=4 * (parts[0]["in_elems"] + parts[2]["out_elems"]),
                         ceilings_ms={k: round(v * 1e3, 5) for k, v in t.items()})
            else:
                continue
        elif base in rows:
            # a layer is bound by whichever is slower at peak: its FLOPs on the tensor pipe or its
            # activation + weight bytes through HBM (FP32-equivalent storage: 4 B per element)
            r = rows[base]
            flops = 2.0 * r["macs"] * batch
            byts = 4.0 * ((r["in_elems"] + r["out_elems"]) * batch + r["w_elems"])
            t_tensor = flops / (peaks["tf_sust"] * 1e12)
            t_hbm = byts / (peaks["hbm"] * 1e9)
            if r["kind"] != "dw" and t_tensor >= t_hbm:
                d.update(bound="tensor", achieved=flops / (ms * 1e-3) / 1e12, peak=peaks["tf_sust"], unit="TFLOP/s",
                         alg_per_segment=2 * r["macs"])
            else:
                d.update(bound="hbm", achieved=byts / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                         alg_per_segment=4 * (r["in_elems"] + r["out_elems"]),
                         tensor_tflops=flops / (ms * 1e-3) / 1e12)
        elif name.startswith("spectrogram"):
            idx = int(name[-1]) if name[-1].isdigit() else None
            specs = fe.specs if idx is None else [fe.specs[idx]]
            # algorithmic bytes: audio read once (shared by both branches: counted half each) + spectrogram written once
            per = sum(fe.sample_count * 4 / len(fe.specs) + sp.n_mels * sp.n_frames(fe.sample_count) * 4 for sp in specs)
            fl = sum(2.0 * sp.n_frames(fe.sample_count) * sp.n_fft * sp.n_mels for sp in specs)
            # bound = whichever is slower at peak: the contraction on the tensor pipe or the audio + spectrogram bytes through
            # HBM.  `achieved` counts the ALGORITHMIC flops (2 per MAC); the FP32-equivalent policy issues three fp16
            # products per MAC (hi*hi, hi*lo, lo*hi), which `frac_counting_3_products` shows.
            t_hbm = per * batch / (peaks["hbm"] * 1e9)
            t_tensor = fl * batch / (peaks["tf_sust"] * 1e12)
            ceil = {"hbm": t_hbm * 1e3, "tensor": t_tensor * 1e3, "tensor_3_products": 3.0 * t_tensor * 1e3}
            if t_tensor >= t_hbm:
                ach = fl * batch / (ms * 1e-3) / 1e12
                d.update(bound="tensor", achieved=ach, peak=peaks["tf_sust"], unit="TFLOP/s", alg_per_segment=fl,
                         products_per_mac=3, frac_counting_3_products=3.0 * ach / peaks["tf_sust"],
                         hbm_gbs=per * batch / (ms * 1e-3) / 1e9, alg_bytes_per_segment=per, ceilings_ms=ceil)
            else:
                d.update(bound="hbm", achieved=per * batch / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                         alg_per_segment=per, tensor_equiv_tflops=fl * batch / (ms * 1e-3) / 1e12, ceilings_ms=ceil)
        elif name == "logmel":
            sp = fe.specs[0]
            per = fe.sample_count * 4 + fe.n_frames() * sp.n_mels * 4      # audio read once + spectrogram written once
            d.update(bound="hbm", achieved=per * batch / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s", alg_per_segment=per)
        elif name == "normalize":
            # min/max pass: the audio is read once (the normalised copy is no longer materialised)
            byts = fe.sample_count * 4 * batch
            d.update(bound="hbm", achieved=byts / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                     alg_per_segment=fe.sample_count * 4)
        elif name == "topk_epilogue":
            byts = (spec.num_species * 4 + 5 * 8) * batch
            d.update(bound="hbm", achieved=byts / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s",
                     alg_per_segment=spec.num_species * 4 + 40)
        else:
            continue
        d["frac"] = d["achieved"] / d["peak"]
        out.append(d)
    dom = max(out, key=lambda x: x["ms"]) if out else None
    return dom, out


def _pinned_copy_rates(torch, device, in_bytes, out_bytes, reps=10, barrier=None):
    """Live staging roofline: pinned cudaMemcpyAsync H2D / D2H rates (GB/s) of batch-sized buffers on this GPU, CUDA
    events on the copy stream.  With several ranks every rank copies at the same time (barrier first), so the figure is
    this GPU's share of the box's host-to-device fabric, not the rate of a GPU copying alone (tools/pcie_bench.cu:
    55.5 GB/s alone, 23-35 GB/s each when eight GPUs copy together)."""
    h_in = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=device)
    d_out = torch.empty(out_bytes, dtype=torch.uint8, device=device)
    st = torch.cuda.Stream(device=device)
    rates = []
    with torch.cuda.stream(st):
        for src, dst, n in ((h_in, d_in, in_bytes), (d_out, h_out, out_bytes)):
            for _ in range(2):
                dst.copy_(src, non_blocking=True)
            st.synchronize()
            if barrier is not None:
                barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            e1.record(st)
            e1.synchronize()
            rates.append(reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return rates[0], rates[1]


def _range_mask(n: int, seed: int = 5):
    """Dense tri-state + scores from a seeded synthetic location-score vector: scores = the
    reference's mock_embeddings LCG, 30 % of the species absent from the meta model's list, threshold 0.01 for the
    list (RangeFilter::predict) and 0.03 for the filter so the drop arm fires."""
    from birdnet_b200.modelgen import synth
    score = synth.mock_embeddings(n, seed).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(seed))
    present = (rng.random(n) < 0.7) & (score >= np.float32(0.01))
    state = np.where(~present, 0, np.where(score >= np.float32(0.03), 1, 2)).astype(np.uint8)
    return state, score


def run_ours(args):
    import torch
    import birdnet_b200 as bb
    from birdnet_b200.modelgen import get_spec, synth
    from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels

    cfg = CONFIGS[args.config]
    family = cfg["family"]
    rank, local, world = _dist()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this build has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_node = _bind_near_gpu(local) if world > 1 else None      # several ranks per box: each stays beside its GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = _peaks()
    spec = get_spec(family)
    fe = spec.frontend
    S, SR = fe.sample_count, fe.sample_rate
    path = ensure_model(family)
    # host gather threads per call: the ranks of one box share its cores
    pack_threads = max(2, ((os.cpu_count() or 2) // max(world, 1)) // 2)
    clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species))
           .top_k(5).min_confidence(0.1).device_id(local).pack_threads(pack_threads).build())
    B = args.batch or cfg["batch"]
    K, W = args.steps, max(args.warmup, 3)
    perch = family == "perch_v2"
    strong = "total_segments" in cfg
    if strong:
        state, score = _range_mask(spec.num_species)
        clf.set_range_filter_dense(state, score, True)
        # strong scaling: the 28,800 segments are split over the ranks in whole batches (multi_gpu.shard_range)
        from birdnet_b200.multi_gpu import shard_range
        lo, hi = shard_range(cfg["total_segments"], rank, world, B)
        my_batches = [(x, min(hi, x + B)) for x in range(lo, hi, B)]
        K = len(my_batches)
    # each rank owns its own shard of the synthetic stream (weak scaling: B segments per rank per step; strong
    # scaling: the recording is the 256 distinct synthetic segments repeated, the rank's batches rotate through them)
    audio = synth.batch(rank * B, B, S, SR)
    segs = list(audio)
    new_ctx = (lambda: clf.create_batch_context(B, allow_perch=True)) if perch else (lambda: clf.create_batch_context(B))
    ctx = new_ctx()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident input, K steps back to back, alternating over the engine's compute lanes ----
    d_audio = torch.from_numpy(audio).cuda(local)
    lanes = [ctx]
    while len(lanes) < clf.compute_lanes():
        lanes.append(new_ctx())
    streams = [torch.cuda.ExternalStream(c.stream_ptr(), device=torch.device("cuda", local)) for c in lanes]
    for i in range(W * len(lanes)):
        lanes[i % len(lanes)].enqueue_device(d_audio.data_ptr(), B, True)
    for c in lanes:
        c.wait()
    launches_per_step = ctx.last_launch_count()
    sampler = ClockSampler(local)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev_end = [torch.cuda.Event(enable_timing=True) for _ in lanes]
    barrier()
    sampler.start()
    ev0.record(streams[0])
    n_dev_segments = 0
    for i in range(K):
        nb = (my_batches[i][1] - my_batches[i][0]) if strong else B      # the last batch of a shard may be ragged
        lanes[i % len(lanes)].enqueue_device(d_audio.data_ptr(), nb, True)
        n_dev_segments += nb
    for e, s_ in zip(ev_end, streams):
        e.record(s_)
    for c in lanes:
        c.wait()
    barrier()
    dev_ms = max(ev0.elapsed_time(e) for e in ev_end)
    # keep the GPU busy a little longer so the 100 ms sampler sees load even for short runs
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        ctx.enqueue_device(d_audio.data_ptr(), B, False)
        ctx.wait()
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([dev_ms, float(n_dev_segments)], device=f"cuda:{local}", dtype=torch.float64)
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_ms, n_all = float(tm[0].item()), float(t[1].item())
    else:
        n_all = float(n_dev_segments)
    value = n_all / (dev_ms * 1e-3)

    # ---- e2e: public API with host slices, `depth` contexts driven by `depth` host threads ------
    # headline: the caller keeps its segments in page-locked host memory (bb.pinned_array -> bn_host_alloc), so
    # the engine DMAs them in place; `e2e_pageable`: ordinary numpy memory, gathered into the engine's pinned slab
    use_ctx = cfg["api"] == "ctx"
    depth = max(1, args.pipeline_depth)
    if not use_ctx:
        depth = min(depth, 4)                               # bn_engine_run: pool of 4 internal contexts
    ctxs = ([ctx] + [new_ctx() for _ in range(depth - 1)]) if use_ctx else []
    pinned = bb.pinned_array(audio.shape)
    pinned[:] = audio
    segs_pinned = list(pinned)
    E2E_REPS = 5                                         # median of >= 5 runs
    if strong:
        n_e2e = K
    else:
        n_e2e = depth * max(16, -(-K // depth))          # whole rounds: every thread drives the same number of batches
    # every call hands back ~1,800 small result objects; with torch's million-object heap loaded, the full garbage
    # collections they trigger each stop all driver threads for ~25 ms (seen in the BN_TRACE_RUN timeline) - park the
    # start-up heap in the permanent generation, as a long-running service would
    import gc
    gc.collect()
    gc.freeze()

    run_log = []

    def call(t, seg_list):
        if use_ctx:
            return clf.predict_batch_with_context(ctxs[t], seg_list)
        return clf.predict_batch(seg_list)

    def run_e2e(seg_list):
        for t in range(depth):
            for _ in range(2):
                call(t, seg_list)
        sink = [0] * depth
        done = [0] * depth

        def e2e_worker(t):
            for i in range(t, n_e2e, depth):
                sl = seg_list
                if strong and (my_batches[i][1] - my_batches[i][0]) != B:
                    sl = seg_list[: my_batches[i][1] - my_batches[i][0]]
                res = call(t, sl)
                sink[t] += len(res[0].predictions) + len(res)
                done[t] += len(res)
        rates = []
        for _ in range(E2E_REPS):
            for t in range(depth):
                done[t] = 0
            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=e2e_worker, args=(t,)) for t in range(depth)]
            [x.start() for x in th]
            [x.join() for x in th]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            nseg = float(sum(done))
            if dist is not None:
                tm = torch.tensor([dt], device=f"cuda:{local}", dtype=torch.float64)
                ts = torch.tensor([nseg], device=f"cuda:{local}", dtype=torch.float64)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dist.all_reduce(ts, op=dist.ReduceOp.SUM)
                dt, nseg = float(tm.item()), float(ts.item())
            rates.append(nseg / dt)
        run_log.append([round(r) for r in rates])
        return float(np.median(rates))

    e2e_value = run_e2e(segs_pinned)
    e2e_pageable = run_e2e(segs)
    E = spec.embedding_dim or 0
    h2d = B * S * 4
    d2h = B * spec.num_species * 4 + B * E * 4 + B * 5 * 8 + B * 4

    # single-call latency of the API at this batch (one thread, nothing else in flight)
    lat = []
    for _ in range(7):
        t0 = time.perf_counter()
        call(0, segs_pinned)
        lat.append(1e3 * (time.perf_counter() - t0))
    call_latency_ms = float(np.median(lat))

    # ---- extra: the same metric through the on-device ingest path: the recording is
    # handed over as 16-bit PCM and read_wav's conversion + chunk_audio run on the GPU (bn_ctx_run_pcm16) ----
    ingest = None
    if not args.no_ingest and use_ctx and not strong:
        n_b = 8                                                   # batches per recording
        pcm = bb.pinned_array((n_b * B * S,), np.int16)           # the recording sits in page-locked host memory
        pcm[:] = np.tile((np.clip(audio.reshape(-1), -1.0, 1.0) * 32767.0).astype(np.int16), n_b)
        for c in ctxs:
            clf.predict_pcm16_stream(c, pcm[: B * S])
        got = [0] * depth

        def pcm_worker(t):
            got[t] = len(clf.predict_pcm16_stream(ctxs[t], pcm, 0.0))
        pcm_rates = []
        for _ in range(E2E_REPS):
            barrier()
            t0 = time.perf_counter()
            th = [threading.Thread(target=pcm_worker, args=(t,)) for t in range(depth)]
            [x.start() for x in th]
            [x.join() for x in th]
            torch.cuda.synchronize()
            pcm_s = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([pcm_s], device=f"cuda:{local}")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                pcm_s = float(t.item())
            pcm_rates.append(world * sum(got) / pcm_s)
        ingest = {"value": float(np.median(pcm_rates)), "unit": UNIT, "h2d_bytes_per_step": B * S * 2,
                  "segments": world * sum(got), "api": "Classifier.predict_pcm16_stream -> bn_ctx_run_pcm16 "
                  "(16-bit PCM in, conversion + chunking on the device; not the reference's f32-slice API)"}

    # ---- staging roofline: every rank copies at the same time -> this GPU's share of the box's host-to-device fabric ----
    h2d_gbs, d2h_gbs = _pinned_copy_rates(torch, torch.device("cuda", local), h2d, max(d2h, 1 << 20), barrier=barrier)
    h2d_box = h2d_gbs
    if dist is not None:
        t = torch.tensor([h2d_gbs], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d_box = float(t.item())

    # ---- roofline of the dominant kernel, timed live with CUDA events on the engine's stream ----
    roof, stages, classes = None, [], []
    if rank == 0:
        ctx.set_profiling(True)
        acc = {}
        reps = 5
        for _ in range(reps):
            ctx.run_device(d_audio.data_ptr(), B, True)
            for name, ms in ctx.stage_times():
                acc.setdefault(name, []).append(ms)
        ctx.set_profiling(False)
        st = [(n, float(np.mean(v))) for n, v in acc.items()]
        dom, stages = _stage_roofline(spec, st, B, peaks)
        tot_ms = sum(ms for _, ms in st) or 1.0
        by_class = {}
        for n, ms in st:
            c = by_class.setdefault(_kernel_class(n), {"ms": 0.0, "stages": 0})
            c["ms"] += ms
            c["stages"] += 1
        classes = [{"class": k, "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4), "stages": v["stages"]}
                   for k, v in sorted(by_class.items(), key=lambda kv: -kv[1]["ms"])]
        # staging over PCIe: what the end-to-end API moved per second against the pinned-copy rate measured live
        moved = e2e_value * (h2d + d2h) / B / 1e9                          # GB/s the end-to-end run moved, all GPUs
        stages.append({"stage": "staging (H2D of the segments + D2H of the results, whole job)", "bound": "pcie",
                       "achieved": moved, "peak": h2d_box, "unit": "GB/s", "frac": moved / h2d_box,
                       "alg_per_segment": (h2d + d2h) / B, "ms": 1e3 * (h2d / (h2d_gbs * 1e9) + d2h / (d2h_gbs * 1e9)),
                       "share": None, "peak_source": "pinned cudaMemcpyAsync H2D measured in this run with all %d ranks copying at once: "
                                                     "%.1f GB/s aggregate (rank 0: %.1f GB/s, D2H %.1f GB/s)" % (world, h2d_box, h2d_gbs, d2h_gbs)})
        if dom:
            traffic, traffic_src = _ncu_traffic()
            roof = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
                    "frac": dom["frac"], "traffic": traffic.get(dom["stage"]), "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": (dom["alg_per_segment"] * B if dom["bound"] == "hbm" else None),
                    "kernel": dom["stage"], "kernel_ms": dom["ms"],
                    "share_of_step": dom["share"], "peak_source": peaks["src"] + (" sustained" if dom["bound"] == "tensor" else ""),
                    "algorithmic_per_segment": dom["alg_per_segment"],
                    "algorithmic_flops_per_launch": (dom["alg_per_segment"] * B if dom["bound"] == "tensor" else None),
                    "dominant_class": classes[0] if classes else None}
            for k in ("products_per_mac", "frac_counting_3_products", "ceilings_ms"):
                if k in dom:
                    roof[k] = dom[k]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, nseg, dt, _ = _cpu_oracle_rate(12.0, 8, 4096, family)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nseg} segments in batches of 8 ({dt:.1f} s), torch CPU FP32 oracle port; ORT CPU cannot run in this image"}

    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "bench_config": args.config,
                       "global_batch": B * world, "segment_samples": S, "top_k": 5, "min_confidence": 0.1,
                       "total_segments": cfg.get("total_segments"),
                       "l2_policy": "inputs larger than L2 (%d MB batch > 126 MB L2)" % (B * S * 4 // 1000000) if B * S * 4 > 126e6
                                    else "batch of %d MB fits L2; 2 compute lanes alternate two contexts' buffers" % (B * S * 4 // 1000000),
                       "compute_lanes": len(lanes), "numa_node": numa_node,
                       "parallelism": f"{world} independent per-GPU shards, no collective",
                       "precision_policy": "FP32-equivalent (see DESIGN.md)",
                       "api": "predict_batch_with_context" if use_ctx else "predict_batch",
                       "call_latency_ms": call_latency_ms,
                       "e2e_pipeline_depth": depth, "e2e_runs": "median of 5 runs of %d batches" % n_e2e, "e2e_run_values": run_log, "host_gc": "gc.freeze() after start-up", "host_cores": os.cpu_count(), "host_pack_threads_per_call": pack_threads},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "inputs": "%d host slices per step in page-locked host memory (bn_host_alloc), copied to the GPU in place" % B},
            "e2e_pageable": {"value": e2e_pageable, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "inputs": "%d pageable host slices per step (the true drop-in case), gathered into the engine's pinned slab first" % B},
            "ingest_pcm16": ingest,
            "gpu_launches": int(launches_per_step * K),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_classes": classes,
            "stages": [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in s.items()} for s in
                       sorted(stages, key=lambda x: -(x["ms"] or 0))[:14]],
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[n-1]; default 2 = the headline")
    ap.add_argument("--batch", type=int, default=0, help="override the configuration's batch size")
    ap.add_argument("--pipeline-depth", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ingest", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
