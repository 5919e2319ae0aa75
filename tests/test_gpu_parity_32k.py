"""Parity tests for the 32 kHz families (BASELINE configs 3 and 4):
log-mel front-end + CNN + embeddings through the C ABI against the oracle.  -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import birdnet_b200 as bb
from birdnet_b200.modelgen import get_spec, synth
from birdnet_b200.modelgen.make_models import ensure_model, synthetic_labels

LOGIT_TOL = 5e-3     # max-abs on raw logits (FP32-equivalent arithmetic on both sides)
EMB_TOL = 1e-3       # north star: embeddings within max-abs 1e-3
CONF_TOL = 1e-3      # north star: confidences within max-abs 1e-3
SEP_TOL = 2 * LOGIT_TOL
GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _build(fam):
    spec = get_spec(fam)
    path = ensure_model(fam)
    clf = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(spec.num_species))
           .top_k(5).min_confidence(0.1).build())
    return spec, path, clf


def _oracle(spec, path, dtype=None):
    import torch
    from oracle.model_oracle import ModelOracle, load_initializers
    return ModelOracle(spec, load_initializers(path), dtype=dtype or torch.float32)


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x.astype(np.float64)))


def _check(results, ref_logits, ref_emb, model_type, k=5, mc=0.1):
    """North-star tolerances on EVERY segment of every synthetic kind, no per-segment allowance: embeddings and
    confidences within max-abs 1e-3, raw logits within 5e-3 (+3e-4 relative: |logit| reaches 28 where sigmoid
    saturates), top-k lists identical wherever the oracle's scores are separated by more than the tolerance.
    (The log floor of the build-authored log-mel front-ends is 1e-2, which keeps ln() well-conditioned: the FP32
    and FP64 oracles agree to 7e-5 in the logits on all ten kinds - graphspec.FrontEnd.log_floor.)"""
    from oracle import postprocess_oracle as po
    got = np.stack([r.raw_scores for r in results])
    emb = np.stack([r.embeddings for r in results])
    for i in range(len(results)):
        tol = LOGIT_TOL + 3e-4 * np.abs(ref_logits[i]).max()
        assert np.abs(got[i] - ref_logits[i]).max() < tol, (i, np.abs(got[i] - ref_logits[i]).max(), tol)
        assert np.abs(emb[i] - ref_emb[i]).max() < EMB_TOL, (i, np.abs(emb[i] - ref_emb[i]).max())
        assert np.abs(_sigmoid(got[i]) - _sigmoid(ref_logits[i])).max() < CONF_TOL, i
    worst = 0.0
    for i, r in enumerate(results):
        assert r.model_type is model_type
        ref = po.top_k_predictions(ref_logits[i], k, mc)
        srt = np.sort(ref_logits[i])[::-1]
        separated = np.all(np.abs(np.diff(srt[:k + 1])) > SEP_TOL) and \
            np.all(np.abs(srt[:k + 1] - np.log(mc / (1 - mc))) > SEP_TOL)
        if separated:
            # the reference leaves the order of EQUAL confidences unspecified (postprocess.rs:207;
            # sigmoid saturates to exactly 1.0f above ~17): compare order only across distinct values
            assert sorted(p.index for p in r.predictions) == sorted(j for j, _ in ref), i
            got_c = [p.confidence for p in r.predictions]
            assert got_c == sorted(got_c, reverse=True), i
            if len({c for _, c in ref}) == len(ref):
                assert [p.index for p in r.predictions] == [j for j, _ in ref], i
        by_idx = dict(ref)
        for p in r.predictions:
            if p.index in by_idx:
                worst = max(worst, abs(p.confidence - by_idx[p.index]))
    assert worst < CONF_TOL


@pytest.fixture(scope="module")
def v30():
    return _build("birdnet_v30")


@pytest.fixture(scope="module")
def perch():
    return _build("perch_v2")


def test_v30_matches_oracle_and_golden(v30):
    spec, path, clf = v30
    cfg = clf.config()
    assert (cfg.sample_rate, cfg.sample_count, cfg.embedding_dim, cfg.num_species) == (32000, 160000, 1024, 11560)
    audio = synth.batch(0, 10, 160000, 32000)             # one segment of every synthetic kind
    orc = _oracle(spec, path)
    ref_logits, ref_emb = orc.logits_and_embeddings(audio)
    res = clf.predict_batch(list(audio))                   # classifier.rs:676-727
    _check(res, ref_logits, ref_emb, bb.ModelType.BirdNetV30)
    ctx = clf.create_batch_context(10)                     # batch_context.rs:252-262 (output_0 / output_1)
    res2 = clf.predict_batch_with_context(ctx, list(audio))
    for a, b in zip(res, res2):
        assert np.array_equal(a.raw_scores, b.raw_scores) and np.array_equal(a.embeddings, b.embeddings)
    # log-mel front-end alone: [frames][mels], ln() of FP32 FFT magnitudes on both sides
    import torch
    ref64_spec = _oracle(spec, path, torch.float64).forward(audio, keep=["spec"])["spec"].reshape(10, -1)
    spec_gpu = ctx.read_tensor("spec", 10)
    for i in range(10):                                    # against the FP64 oracle, every segment, values in [-0.46, 0.61]
        d = np.abs(spec_gpu[i] - ref64_spec[i])
        assert d.max() < 1e-3 and d.mean() < 2e-5, (i, d.max(), d.mean())
    g = np.load(os.path.join(GOLD_DIR, "v30_seed0.npz"))
    got = np.stack([r.raw_scores for r in res])
    emb = np.stack([r.embeddings for r in res])
    assert np.abs(got[:, ::32] - g["logits_every_32"]).max() < 2 * LOGIT_TOL         # all ten segments
    assert np.abs(emb[:, ::8] - g["emb_every_8"]).max() < EMB_TOL
    one = clf.predict(audio[4])                            # batch-size invariance, bit for bit
    assert np.array_equal(one.raw_scores, got[4]) and np.array_equal(one.embeddings, emb[4])


def test_perch_matches_oracle_and_golden(perch):
    spec, path, clf = perch
    cfg = clf.config()
    assert (cfg.sample_count, cfg.embedding_dim, cfg.num_species) == (160000, 1536, 14795)   # detection.rs:214-232
    audio = synth.batch(0, 10, 160000, 32000)
    ref_logits, ref_emb = _oracle(spec, path).logits_and_embeddings(audio)
    res = clf.predict_batch(list(audio))
    _check(res, ref_logits, ref_emb, bb.ModelType.PerchV2)
    g = np.load(os.path.join(GOLD_DIR, "perch_seed0.npz"))
    got = np.stack([r.raw_scores for r in res])
    assert np.abs(got[:, ::32] - g["logits_every_32"]).max() < 2 * LOGIT_TOL          # all ten segments
    with pytest.raises(bb.Inference) as e:                 # batch_context.rs:107-114
        clf.create_batch_context(4)
    assert str(e.value) == ("inference failed: BatchInferenceContext does not yet support PerchV2 models. "
                            "Use predict_batch() instead.")
    # opt-in context for Perch (bn_ctx_create_ex + BN_CTX_ALLOW_PERCH), same staged path, same bits;
    # outputs 1 and 2 (which the reference ignores, classifier.rs:929-934) stay on the device unless asked for
    ctx = clf.create_batch_context(10, allow_perch=True)
    res2 = clf.predict_batch_with_context(ctx, list(audio))
    for a, b in zip(res, res2):
        assert np.array_equal(a.raw_scores, b.raw_scores) and np.array_equal(a.embeddings, b.embeddings)
        assert [(p.index, p.confidence) for p in a.predictions] == [(p.index, p.confidence) for p in b.predictions]
    orc = _oracle(spec, path)
    keep = orc.forward(audio)
    spat = ctx.read_tensor("spatial_embedding", 10)
    assert spat.shape == (10, 16 * 4 * 1536)
    ref_spat = np.asarray(keep["spatial_embedding"]).reshape(10, -1)
    # un-pooled activations of the last conv: |x| reaches ~40, same 1e-3 absolute bound + 1e-4 relative
    assert np.abs(spat - ref_spat).max() < 1e-3 + 1e-4 * float(np.abs(ref_spat).max()), np.abs(spat - ref_spat).max()
    assert ctx.read_tensor("spectrogram", 10).shape == (10, 500 * 128)
    with pytest.raises(bb.InputSize) as e:
        clf.predict(np.zeros(144000, dtype=np.float32))
    assert str(e.value) == "input size mismatch: expected 160000 samples, got 144000"


def test_v30_full_batch_512_properties(v30):
    """BASELINE config 3 at full size: batch 512 with the 1024-dim embedding output."""
    _, _, clf = v30
    B = 512
    audio = synth.batch(0, B, 160000, 32000)
    ctx = clf.create_batch_context(B)
    assert ctx.input_buffer_bytes() == 327_680_000
    res = clf.predict_batch_with_context(ctx, list(audio))
    logits = np.stack([r.raw_scores for r in res])
    emb = np.stack([r.embeddings for r in res])
    assert logits.shape == (B, 11560) and emb.shape == (B, 1024)
    assert np.isfinite(logits).all() and np.isfinite(emb).all()
    perm = np.random.Generator(np.random.PCG64(1)).permutation(B)
    res_p = clf.predict_batch_with_context(ctx, [audio[j] for j in perm])
    assert np.array_equal(np.stack([r.raw_scores for r in res_p]), logits[perm])
    assert np.array_equal(np.stack([r.embeddings for r in res_p]), emb[perm])
    small = clf.predict_batch_with_context(ctx, list(audio[300:305]))
    assert np.array_equal(np.stack([r.raw_scores for r in small]), logits[300:305])
    assert np.array_equal(logits[9], logits[19])            # identical inputs (silence)
    for r in res:
        c = [p.confidence for p in r.predictions]
        assert len(c) <= 5 and c == sorted(c, reverse=True) and all(x >= 0.1 for x in c)
        top = np.argsort(-r.raw_scores, kind="stable")[:len(c)]
        assert [p.index for p in r.predictions] == top.tolist()


def test_perch_full_batch_256_properties(perch):
    """BASELINE config 4 at full size via predict_batch (the context path rejects Perch)."""
    _, _, clf = perch
    B = 256
    audio = synth.batch(0, B, 160000, 32000)
    res = clf.predict_batch(list(audio))
    logits = np.stack([r.raw_scores for r in res])
    emb = np.stack([r.embeddings for r in res])
    assert logits.shape == (B, 14795) and emb.shape == (B, 1536)
    assert np.isfinite(logits).all() and np.isfinite(emb).all()
    small = clf.predict_batch(list(audio[40:47]))
    assert np.array_equal(np.stack([r.raw_scores for r in small]), logits[40:47])
    assert np.array_equal(np.stack([r.embeddings for r in small]), emb[40:47])
    assert np.array_equal(logits[9], logits[19])
    for r in res:
        c = [p.confidence for p in r.predictions]
        top = np.argsort(-r.raw_scores, kind="stable")[:len(c)]
        assert [p.index for p in r.predictions] == top.tolist()
