#!/usr/bin/env python
"""GPU check of the fused v2.4 front-end: spectrogram of the same batch through the fused kernel and through the
frame-matrix path (BN_DISABLE_FE_FUSED=1), both against the FP64 oracle (the two measures of tests/test_gpu_parity.py).

    python tools/fe_check.py [--batch 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=10)
    args = ap.parse_args()
    import torch
    import birdnet_b200 as bb
    from birdnet_b200.modelgen import get_spec, synth
    from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
    from oracle.model_oracle import ModelOracle, load_initializers
    spec = get_spec("birdnet_v24")
    path = ensure_model("birdnet_v24")
    B = args.batch
    audio = synth.batch(0, B, 144000, 48000)
    orc = ModelOracle(spec, load_initializers(path))
    o64 = type(orc)(orc.spec, {k: v.numpy() for k, v in orc.w.items()}, dtype=torch.float64)
    ref = o64.forward(audio, keep=["spec"])["spec"]
    ref32 = orc.forward(audio, keep=["spec"])["spec"].astype(np.float64)
    e2 = 2.0 * float(orc.w["fe.spec0.exponent"])
    outs = {}
    for mode in ("fused", "planes"):
        if mode == "planes":
            os.environ["BN_DISABLE_FE_FUSED"] = "1"
        clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build())
        os.environ.pop("BN_DISABLE_FE_FUSED", None)
        ctx = clf.create_batch_context(B)
        clf.predict_batch_with_context(ctx, list(audio))
        outs[mode] = ctx.read_tensor("spec", B).reshape(B, 96, 511, 2).transpose(0, 3, 1, 2).astype(np.float64)
    outs["fp32-oracle"] = ref32
    for mode, sp in outs.items():
        for br in range(2):
            rel, lin = [], []
            for i in range(B):
                a, b = sp[i, br], ref[i, br]
                peak = float(ref[i].max())
                if peak <= 0:
                    rel.append(0.0); lin.append(0.0); continue
                big = b > 0.05 * peak
                rel.append(float((np.abs(a - b)[big] / b[big]).max()) if big.any() else 0.0)
                lin.append(float(np.abs(np.maximum(a, 0) ** (1 / e2) - b ** (1 / e2)).max() / peak ** (1 / e2)))
            print(f"{mode:12s} branch {br}  rel(>5% peak) " + " ".join(f"{v:.1e}" for v in rel))
            print(f"{mode:12s} branch {br}  lin/fullscale " + " ".join(f"{v:.1e}" for v in lin))
    d = np.abs(outs["fused"] - outs["planes"])
    print("fused vs planes: max", d.max(), "at", np.unravel_index(d.argmax(), d.shape))


if __name__ == "__main__":
    main()
