"""Parity + timing diagnostic for the 32 kHz families (v3.0-like, Perch-v2-like); run on a GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-birdnet-onnx_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels
from oracle.model_oracle import ModelOracle, load_initializers
from oracle import postprocess_oracle as po

fams = sys.argv[1:] or ["birdnet_v30", "perch_v2"]
for fam in fams:
    path = ensure_model(fam)
    spec = get_spec(fam)
    t0 = time.time()
    clf = bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species)).top_k(5).min_confidence(0.1).build()
    print(fam, "build s", round(time.time() - t0, 2), clf.config())
    B = 10
    audio = synth.batch(0, B, 160000, 32000)
    res = clf.predict_batch(list(audio))
    orc = ModelOracle(spec, load_initializers(path))
    ref_logits, ref_emb = orc.logits_and_embeddings(audio)
    got = np.stack([r.raw_scores for r in res])
    emb = np.stack([r.embeddings for r in res])
    print(" logits max|d|", np.abs(got - ref_logits).max(), "ref max", np.abs(ref_logits).max(),
          "per seg", np.round(np.abs(got - ref_logits).max(axis=1), 5))
    print(" emb max|d|", np.abs(emb - ref_emb).max(), "ref max", np.abs(ref_emb).max())
    for i, r in enumerate(res):
        o = po.top_k_predictions(ref_logits[i], 5, 0.1)
        ok = [p.index for p in r.predictions] == [j for j, _ in o]
        dc = max([abs(p.confidence - c) for p, (_, c) in zip(r.predictions, o)] + [0])
        print("  ", i, ok, round(dc, 6), [(p.index, round(p.confidence, 4)) for p in r.predictions][:3])
    if fam != "perch_v2":
        ctx = clf.create_batch_context(B)
        res2 = clf.predict_batch_with_context(ctx, list(audio))
        print(" ctx == engine_run:", all(np.array_equal(a.raw_scores, b.raw_scores) for a, b in zip(res, res2)))
        ref = orc.forward(audio, keep=["spec"])
        sp = ctx.read_tensor("spec", B)
        rs = ref["spec"].reshape(B, -1)
        print(" spec max|d|", np.abs(sp - rs).max(), "ref max", np.abs(rs).max(), "mean|d|", np.abs(sp - rs).mean())
        Bb = int(os.environ.get("B", "256"))
        big = synth.batch(0, Bb, 160000, 32000)
        d = torch.from_numpy(big).cuda()
        ctx = clf.create_batch_context(Bb)
        for _ in range(2):
            ctx.run_device(d.data_ptr(), Bb, True)
        ctx.set_profiling(True)
        ctx.run_device(d.data_ptr(), Bb, True)
        st = ctx.stage_times()
        tot = sum(ms for _, ms in st)
        for name, ms in sorted(st, key=lambda x: -x[1])[:12]:
            print(f"   {name:28s} {ms:8.3f} ms {100*ms/tot:5.1f}%")
        ctx.set_profiling(False)
        t = time.time()
        for _ in range(10):
            ctx.enqueue_device(d.data_ptr(), Bb, True)
        ctx.wait()
        dt = (time.time() - t) / 10
        print(f" device-resident batch {Bb}: {dt*1e3:.2f} ms -> {Bb/dt:.0f} seg/s (stage sum {tot:.2f} ms)")
    else:
        try:
            clf.create_batch_context(4)
            print(" ERROR: Perch context should be rejected")
        except bb.Inference as ex:
            print(" ctx rejected:", ex)
        Bb = int(os.environ.get("B", "256"))
        big = list(synth.batch(0, Bb, 160000, 32000))
        for _ in range(2):
            clf.predict_batch(big)
        t = time.time()
        for _ in range(3):
            clf.predict_batch(big)
        dt = (time.time() - t) / 3
        print(f" predict_batch({Bb}) host e2e: {dt*1e3:.2f} ms -> {Bb/dt:.0f} seg/s")
