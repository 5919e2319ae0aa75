"""Parity tests proper: the CUDA path (through the C ABI) against the oracle.  -m gpu."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.modelgen import synth
from birdnet_b200.modelgen.make_models import synthetic_labels

LOGIT_TOL = 5e-3     # max-abs on raw logits (FP32-equivalent arithmetic on both sides)
CONF_TOL = 1e-3      # north star: confidences within max-abs 1e-3
SEP_TOL = 2 * LOGIT_TOL   # top-k must be identical where the oracle's scores are this far apart

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v24_seed0.npz")


@pytest.fixture(scope="module")
def clf(v24_model_path, v24_spec):
    return (bb.Classifier.builder().model_path(v24_model_path).labels(synthetic_labels(v24_spec.num_species))
            .top_k(5).min_confidence(0.1).build())


@pytest.fixture(scope="module")
def oracle(v24_model_path, v24_spec):
    from oracle.model_oracle import ModelOracle, load_initializers
    return ModelOracle(v24_spec, load_initializers(v24_model_path))


# ------------------------------------------------------------------ epilogue alone: bit-exact
def _gpu_topk(logits, k, min_conf=None, state=None, score=None, rerank=False):
    a = np.ascontiguousarray(logits, dtype=np.float32)
    if a.ndim == 1:
        a = a[None]
    rows, n = a.shape
    ke = min(k, n)
    out = np.zeros((rows, max(ke, 1), 2), dtype=np.uint32)
    cnt = np.zeros(rows, dtype=np.uint32)
    st = _ffi.lib.bn_topk_apply(None, a.ctypes.data_as(C.POINTER(C.c_float)), rows, n, k,
                                0 if min_conf is None else 1, 0.0 if min_conf is None else min_conf,
                                state.ctypes.data_as(C.POINTER(C.c_uint8)) if state is not None else None,
                                score.ctypes.data_as(C.POINTER(C.c_float)) if score is not None else None,
                                1 if rerank else 0, out.ctypes.data_as(C.POINTER(_ffi.Pred)),
                                cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert st == 0, _ffi.last_error()
    return out[:, :ke, 0], out[:, :ke, 1].copy().view(np.float32), cnt


def test_epilogue_reference_known_answers():
    from oracle import postprocess_oracle as po
    cases = [([0.1, 0.5, 0.9, 0.3, 0.7], 3, None), ([-5.0, 0.0, 5.0], 10, 0.4), ([0.1, 0.2], 100, None),
             ([0.1, 0.9, 0.5], 3, None), ([-10.0, -5.0, -1.0, -20.0], 2, None), ([-10.0, 0.0, 10.0], 10, 0.0),
             ([-10.0, 0.0, 10.0], 10, 1.0), ([0.1, 0.2, 0.3], 2 ** 64 - 1, None), ([0.1, 0.2, 0.3, 0.4], 4, None)]
    for lg, k, mc in cases:
        idx, conf, cnt = _gpu_topk(lg, k, mc)
        ref = po.top_k_predictions(lg, k, mc)
        assert cnt[0] == len(ref), (lg, k, mc)
        assert idx[0, :cnt[0]].tolist() == [i for i, _ in ref]
        assert np.allclose(conf[0, :cnt[0]], [c for _, c in ref], atol=2e-7, rtol=0)
    _, _, cnt = _gpu_topk([0.1, 0.2, 0.3], 0)
    assert cnt[0] == 0                                             # k = 0 -> empty (postprocess.rs:46)
    idx, conf, cnt = _gpu_topk([0.5] * 4, 2)                       # ties: two results, equal confidence
    assert cnt[0] == 2 and conf[0, 0] == conf[0, 1] and idx[0].tolist() == [0, 1]   # documented rule: lower index first
    idx, conf, cnt = _gpu_topk([1.0, float("nan"), 2.0, 0.5], 3)   # +NaN is the largest in total order
    assert cnt[0] == 3 and idx[0, 0] == 1 and np.isnan(conf[0, 0]) and idx[0, 1:].tolist() == [2, 0]
    idx, conf, cnt = _gpu_topk([1.0, float("nan"), 2.0, 0.5], 3, 0.0)   # NaN >= min is false -> dropped
    assert cnt[0] == 2 and idx[0, :2].tolist() == [2, 0]
    idx, conf, cnt = _gpu_topk([float("inf"), float("-inf"), 0.0], 3)
    assert conf[0].tolist() == [1.0, 0.5, 0.0] and idx[0].tolist() == [0, 2, 1]
    idx, _, _ = _gpu_topk([-0.0, 0.0], 2)
    assert idx[0].tolist() == [1, 0]                               # total_cmp: -0 < +0


@pytest.mark.parametrize("n,k,mc", [(6522, 5, 0.1), (6522, 10, None), (6522, 100, 0.3), (14795, 5, None),
                                    (1000, 1000, None), (6522, 6522, 0.5), (37, 5, None), (1, 3, None)])
def test_epilogue_bit_exact_on_lcg_logits(n, k, mc):
    from oracle import postprocess_oracle as po
    rows = 24
    lg = np.stack([po.random_logits(n, 1000 + s) for s in range(rows)])
    # LCG logits take only 65,536 distinct values: break ties so the comparison is pinned
    lg = (lg + np.arange(n, dtype=np.float32)[None, :] * np.float32(1e-7)).astype(np.float32)
    idx, conf, cnt = _gpu_topk(lg, k, mc)
    ridx, rconf, rcnt = po.top_k_batch(lg, k, mc)
    assert np.array_equal(cnt, rcnt)
    for r in range(rows):
        c = cnt[r]
        # oracle order == confidence-desc; equal confidences (saturated sigmoid) may permute
        assert sorted(idx[r, :c].tolist()) == sorted(ridx[r, :c].tolist())
        assert np.abs(conf[r, :c] - rconf[r, :c]).max(initial=0) <= 2e-7
        strict = np.nonzero(np.diff(rconf[r, :c]) < 0)[0]
        if len(strict) == c - 1:
            assert idx[r, :c].tolist() == ridx[r, :c].tolist()


@pytest.mark.parametrize("n,k", [(14795, 14795), (14795, 2 ** 64 - 1), (40000, 50), (70001, 70001)])
def test_epilogue_has_no_size_limit(n, k):
    """postprocess.rs:50 clamps k to n and accepts any n; sorts that do not fit shared memory run from a
    global workspace (k > ~11.7k on Perch used to be cudaErrorInvalidValue)."""
    from oracle import postprocess_oracle as po
    rows = 3
    lg = np.stack([po.random_logits(n, 4000 + s) for s in range(rows)])
    lg = (lg + np.arange(n, dtype=np.float32)[None, :] * np.float32(1e-7)).astype(np.float32)
    idx, conf, cnt = _gpu_topk(lg, k, 0.25)
    ridx, rconf, rcnt = po.top_k_batch(lg, k, 0.25)
    assert np.array_equal(cnt, rcnt)
    for r in range(rows):
        c = cnt[r]
        assert np.abs(conf[r, :c] - rconf[r, :c]).max(initial=0) <= 2e-7
        strict = np.nonzero(np.diff(rconf[r, :c]) < 0)[0]
        if len(strict) == c - 1:
            assert idx[r, :c].tolist() == ridx[r, :c].tolist()
        else:
            assert sorted(idx[r, :c].tolist()) == sorted(ridx[r, :c].tolist())


def test_epilogue_range_mask_and_rerank():
    from oracle import postprocess_oracle as po
    n, rows, k = 6522, 16, 10
    lg = np.stack([po.random_logits(n, 77 + s) for s in range(rows)])
    lg = (lg + np.arange(n, dtype=np.float32)[None, :] * np.float32(1e-7)).astype(np.float32)
    loc = po.mock_embeddings(n, 5)                       # location scores in [0,1]
    rng = np.random.Generator(np.random.PCG64(3))
    present = rng.random(n) < 0.7                        # 30% of species absent from the meta model
    thr = np.float32(0.35)
    state = np.where(~present, 0, np.where(loc >= thr, 1, 2)).astype(np.uint8)
    labels = [f"sp{i}" for i in range(n)]
    loc_list = [(labels[i], float(loc[i])) for i in range(n) if present[i]]
    for rerank in (False, True):
        idx, conf, cnt = _gpu_topk(lg, k, 0.2, state, loc, rerank)
        for r in range(rows):
            base = po.top_k_predictions(lg[r], k, 0.2)
            ref = po.filter_predictions([(labels[i], c, i) for i, c in base], loc_list, float(thr), rerank)
            assert cnt[r] == len(ref)
            assert idx[r, :cnt[r]].tolist() == [i for _, _, i in ref]
            assert np.abs(conf[r, :cnt[r]] - np.array([c for _, c, _ in ref], dtype=np.float32)).max(initial=0) <= 2e-7


def test_range_filter_drop_in_known_answers():
    """RangeFilter::filter_predictions on Prediction lists (rangefilter.rs:703-916), on device."""
    P, L = bb.Prediction, bb.LocationScore
    rf = bb.RangeFilter.from_labels([], threshold=0.03)
    preds = [P("Species A", 0.8, 0), P("Species B", 0.3, 1), P("Species C", 0.05, 2)]
    loc = [L("Species A", 0.9, 0), L("Species B", 0.02, 1), L("Species C", 0.5, 2)]
    assert [p.species for p in rf.filter_predictions(preds, loc, False)] == ["Species A", "Species C"]
    preds = [P("Species A", 0.9, 0), P("Species B", 0.8, 1), P("Species C", 0.7, 2)]
    loc = [L("Species A", 0.5, 0), L("Species B", 0.9, 1), L("Species C", 0.6, 2)]
    f = rf.filter_predictions(preds, loc, True)
    assert [p.species for p in f] == ["Species B", "Species A", "Species C"]
    assert all(abs(p.confidence - w) < 1e-3 for p, w in zip(f, (0.72, 0.45, 0.42)))
    preds = [P("Species A", 0.8, 0), P("Species B", 0.7, 1), P("Species D", 0.9, 3)]
    loc = [L("Species A", 0.9, 0), L("Species C", 0.8, 2)]
    f = rf.filter_predictions(preds, loc, False)
    assert [(p.species, p.index) for p in f] == [("Species A", 0), ("Species B", 1), ("Species D", 3)]
    assert [round(p.confidence, 6) for p in f] == [0.8, 0.7, 0.9]
    rf = bb.RangeFilter.from_labels([], threshold=0.1)
    r = rf.filter_batch_predictions([[P("Species A", 0.8, 0)], [P("Species B", 0.6, 1)]],
                                    [L("Species A", 0.9, 0), L("Species B", 0.05, 1)], False)
    assert [len(x) for x in r] == [1, 0]
    assert rf.filter_predictions([], loc, True) == []


# ------------------------------------------------------------------ front-end + CNN vs oracle
def _check_against_oracle(results, ref_logits, k=5, mc=0.1):
    from oracle import postprocess_oracle as po
    got = np.stack([r.raw_scores for r in results])
    assert np.abs(got - ref_logits).max() < LOGIT_TOL
    worst = 0.0
    for i, r in enumerate(results):
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
        else:
            assert {p.index for p in r.predictions} <= set(np.argsort(-ref_logits[i])[:k + 2].tolist())
        by_idx = dict(ref)
        for p in r.predictions:
            if p.index in by_idx:
                worst = max(worst, abs(p.confidence - by_idx[p.index]))
        assert all(p.species == f"Avis synthetica{p.index}_Synthetic Bird {p.index}" for p in r.predictions)
        assert r.embeddings is None and r.model_type is bb.ModelType.BirdNetV24
    assert worst < CONF_TOL
    return worst


def test_stages_against_oracle(clf, oracle):
    import torch
    B = 10                                               # one segment of every synthetic kind
    audio = synth.batch(0, B, 144000, 48000)
    os.environ["BN_KEEP_NORMALIZED"] = "1"               # keep the FP32 normalised audio for this context
    ctx = clf.create_batch_context(B)
    del os.environ["BN_KEEP_NORMALIZED"]
    res = clf.predict_batch_with_context(ctx, list(audio))
    ref = oracle.forward(audio, keep=["spec"])
    norm_ref = oracle.frontend(torch.from_numpy(audio))["normalized"].numpy()
    assert np.array_equal(ctx.read_normalized(B), norm_ref)          # bit-exact normaliser
    spec = ctx.read_tensor("spec", B).reshape(B, 96, 511, 2).transpose(0, 3, 1, 2).astype(np.float64)
    # Spectrogram against the FP64 oracle.  y = |v|^(2e), e = 0.226: the compression is ill-conditioned only at
    # v -> 0, so (1) bins that are not tiny (> 5 % of the segment's peak; y reaches ~16) must agree to 6e-4
    # RELATIVE, (2) every bin must agree in the linear domain |v| = y^(1/2e) to 3e-5 of the segment's full scale
    # (the FP32 oracle itself sits at 9e-5 / 4.5e-7 on these two measures; the hi/lo fp16 operands have an ABSOLUTE
    # floor of 2^-24 per element, the K = 2048 contraction drops the lo*lo terms and its 134 tensor-core accumulation steps
    # truncate: measured 1.3e-5 of full scale and, on the 5 %-of-peak bins, 1.7e-4 relative with the frame-matrix kernels /
    # 2.7e-4 with the fused front-end kernel, whose K order is column pair by column pair), (3) a loose absolute cap
    # everywhere (worst case: the constant segment, whose spectrum is pure cancellation residue, 2.2e-2 after compression).
    ref64 = type(oracle)(oracle.spec, {k: v.numpy() for k, v in oracle.w.items()}, dtype=torch.float64).forward(audio, keep=["spec"])["spec"]
    e2 = 2.0 * float(oracle.w["fe.spec0.exponent"])
    for i in range(B):
        a, b = spec[i], ref64[i]
        peak = float(b.max())
        if peak <= 0:
            assert np.abs(a).max() == 0                 # silence stays exactly zero
            continue
        big = b > 0.05 * peak
        assert (np.abs(a - b)[big] / b[big]).max() < 6e-4, (i, (np.abs(a - b)[big] / b[big]).max())
        lin = np.abs(np.maximum(a, 0) ** (1 / e2) - b ** (1 / e2)).max() / peak ** (1 / e2)
        assert lin < 3e-5, (i, lin)
        assert np.abs(a - b).max() < 3e-2 and np.abs(a - b).mean() < 2e-3, (i, np.abs(a - b).max(), np.abs(a - b).mean())
    _check_against_oracle(res, ref["output"])
    assert ctx.last_launch_count() > 0


def test_predict_and_predict_batch_match_oracle_and_golden(clf, oracle):
    g = np.load(GOLDEN)
    audio = synth.batch(0, 20, 144000, 48000)
    res = clf.predict_batch(list(audio))                 # cfg1 path (classifier.rs:676-727)
    ref_logits, _ = oracle.logits_and_embeddings(audio)
    _check_against_oracle(res, ref_logits)
    got = np.stack([r.raw_scores for r in res])
    assert np.abs(got[:, ::32] - g["logits_every_32"]).max() < LOGIT_TOL
    for i, r in enumerate(res):
        c = int(g["top5_count"][i])
        assert len(r.predictions) == c or abs(len(r.predictions) - c) <= 1   # threshold-adjacent entries
    one = clf.predict(audio[3])                          # classifier.rs:610-643
    assert np.array_equal(one.raw_scores, got[3])        # batch-size invariance, bit for bit
    assert [p.index for p in one.predictions] == [p.index for p in res[3].predictions]


def test_full_batch_256_properties(clf):
    """BASELINE config 2 at full size: size-independent properties instead of the slow oracle."""
    B = 256
    audio = synth.batch(0, B, 144000, 48000)
    ctx = clf.create_batch_context(B)
    assert ctx.max_batch_size() == 256 and ctx.input_buffer_bytes() == 147_456_000
    res = clf.predict_batch_with_context(ctx, list(audio))
    assert len(res) == B
    logits = np.stack([r.raw_scores for r in res])
    assert np.isfinite(logits).all()
    # (1) permutation equivariance and batch-composition invariance, bit for bit
    perm = np.random.Generator(np.random.PCG64(0)).permutation(B)
    res_p = clf.predict_batch_with_context(ctx, [audio[j] for j in perm])
    assert np.array_equal(np.stack([r.raw_scores for r in res_p]), logits[perm])
    small = clf.predict_batch_with_context(ctx, list(audio[100:107]))        # ragged batch < max
    assert np.array_equal(np.stack([r.raw_scores for r in small]), logits[100:107])
    # (2) the normaliser makes amplitude irrelevant: 440 Hz at 0.5 (k=1) and at 1.0 (k=8)
    assert np.abs(logits[1] - logits[8]).max() < 2e-2
    # (3) identical inputs -> identical outputs (all silence segments)
    assert np.array_equal(logits[9], logits[19])
    # (4) post-processing contract on every segment (integration_test.rs:437-480 properties)
    for r in res:
        c = [p.confidence for p in r.predictions]
        assert len(c) <= 5 and c == sorted(c, reverse=True) and all(x >= 0.1 for x in c)
        top = np.argsort(-r.raw_scores, kind="stable")[:len(c)]
        assert [p.index for p in r.predictions] == top.tolist()


# ------------------------------------------------------------------ API behaviour (error order, options, threads)
def test_validation_order_and_errors(clf):
    good = np.zeros(144000, dtype=np.float32)
    assert clf.predict_batch([]) == []                                  # classifier.rs:681-683
    ctx = clf.create_batch_context(4)
    assert clf.predict_batch_with_context(ctx, []) == []                # classifier.rs:832-834
    with pytest.raises(bb.InputSize) as e:
        clf.predict(np.zeros(1000, dtype=np.float32))
    assert str(e.value) == "input size mismatch: expected 144000 samples, got 1000"
    with pytest.raises(bb.BatchInputSize) as e:
        clf.predict_batch([good, good, np.zeros(50000, dtype=np.float32)])
    assert str(e.value) == "batch input size mismatch: segment 2 has 50000 samples, expected 144000"
    with pytest.raises(bb.Inference) as e:                              # max check first: batch_context.rs:191-196
        clf.predict_batch_with_context(ctx, [np.zeros(5, dtype=np.float32)] * 5)
    assert str(e.value) == "inference failed: batch size 5 exceeds context max 4"
    with pytest.raises(bb.BatchInputSize) as e:
        clf.predict_batch_with_context(ctx, [good, np.zeros(7, dtype=np.float32)])
    assert e.value.index == 1 and e.value.got == 7
    assert len(clf.predict_batch_with_context(ctx, [good, good])) == 2  # context still usable


def test_label_count_mismatch(v24_model_path):
    with pytest.raises(bb.LabelCount) as e:
        bb.Classifier.builder().model_path(v24_model_path).labels(["a", "b"]).build()
    assert str(e.value) == "label count mismatch: model expects 6522, got 2"


def test_timeout_and_cancellation(clf):
    audio = list(synth.batch(0, 64, 144000, 48000))
    ctx = clf.create_batch_context(64)
    ok = clf.predict_batch_with_context(ctx, audio)
    with pytest.raises(bb.Timeout) as e:                                # 1 us cannot be met
        clf.predict_batch_with_context(ctx, audio, bb.InferenceOptions.with_only_timeout(1e-6))
    assert str(e.value) == "inference timed out after 1µs"
    again = clf.predict_batch_with_context(ctx, audio, bb.InferenceOptions.with_only_timeout(60.0))
    assert np.array_equal(again[5].raw_scores, ok[5].raw_scores)        # drained and reusable
    tok = bb.CancellationToken()
    tok.cancel()
    with pytest.raises(bb.Cancelled):
        clf.predict_batch_with_context(ctx, audio, bb.InferenceOptions.new().with_cancellation_token(tok))
    tok2 = bb.CancellationToken()
    r = clf.predict_batch(audio[:2], bb.InferenceOptions.new().with_cancellation_token(tok2).with_timeout(30.0))
    assert len(r) == 2
    # cancel from another thread while a long job is in flight
    tok3 = bb.CancellationToken()
    big = audio * 4
    ctx2 = clf.create_batch_context(256)
    t = threading.Timer(0.002, tok3.cancel)
    t.start()
    try:
        clf.predict_batch_with_context(ctx2, big, bb.InferenceOptions.new().with_cancellation_token(tok3))
    except bb.Cancelled:
        pass                                                            # either outcome is legal: Ok wins if it finished
    t.join()
    assert len(clf.predict_batch_with_context(ctx2, big[:3])) == 3


def test_send_sync_four_threads(clf):                                   # integration_test.rs:495-529
    audio = synth.batch(0, 4, 144000, 48000)
    want = [clf.predict(a).raw_scores for a in audio]
    errs = []

    def work(t):
        try:
            for _ in range(10):
                r = clf.predict(audio[t])
                assert np.array_equal(r.raw_scores, want[t])
        except Exception as ex:       # noqa: BLE001
            errs.append(ex)
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_fused_range_filter_equals_post_filter(clf, v24_spec):
    from oracle import postprocess_oracle as po
    labels = clf.labels()
    loc_all = po.mock_embeddings(v24_spec.num_species, 11)
    # location scores produced with a lower threshold than the filter's (exercises the drop arm)
    loc = [bb.LocationScore(labels[i], float(loc_all[i]), i) for i in range(len(labels)) if loc_all[i] >= 0.2 or i % 3]
    audio = list(synth.batch(0, 10, 144000, 48000))
    plain = clf.predict_batch(audio)
    rf = bb.RangeFilter.from_labels(labels, threshold=0.4)
    for rerank in (False, True):
        want = rf.filter_batch_predictions([r.predictions for r in plain], loc, rerank)
        clf.set_range_filter(loc, 0.4, rerank)
        fused = clf.predict_batch(audio)
        clf.clear_range_filter()
        for w, f, pl in zip(want, fused, plain):
            assert [(p.index, p.species) for p in f.predictions] == [(p.index, p.species) for p in w]
            assert np.allclose([p.confidence for p in f.predictions], [p.confidence for p in w], atol=1e-7)
            ref = po.filter_predictions([(p.species, p.confidence, p.index) for p in pl.predictions],
                                        [(s.species, s.score) for s in loc], 0.4, rerank)
            assert [(p.species, p.index) for p in w] == [(s, i) for s, _, i in ref]


def test_device_pool_matches_single_context(v24_model_path, clf):
    from birdnet_b200.multi_gpu import DevicePool
    n_dev = _ffi.lib.bn_device_count()
    ids = list(range(min(n_dev, 2))) if n_dev > 1 else [0, 0]     # two replicas on one GPU also exercise the gather
    pool = DevicePool(v24_model_path, ids, ctx_batch=8)
    pool.set_postprocess(5, 0.1)
    audio = list(synth.batch(0, 37, 144000, 48000))               # ragged last block
    logits, emb, idx, conf, cnt = pool.run(audio)
    res = clf.predict_batch(audio)
    assert emb is None and np.array_equal(logits, np.stack([r.raw_scores for r in res]))
    for i, r in enumerate(res):
        assert idx[i, :cnt[i]].tolist() == [p.index for p in r.predictions]


@pytest.mark.gpu
def test_pcm16_stream_matches_host_chunking(clf):
    """bn_ctx_run_pcm16 (on-device i16 -> f32 / 32768 + chunk_audio) == the reference CLI's host-side
    read_wav conversion + chunk_audio fed through predict_batch_with_context: bit-exact logits."""
    from oracle import ingest_oracle as io
    rng = np.random.default_rng(7)
    S, sr = 144000, 48000
    n = 5 * S + 12345                                             # ragged tail -> zero padding
    t = np.arange(n, dtype=np.float64) / sr
    pcm = (0.3 * np.sin(2 * np.pi * 1234.0 * t) * 32767 + rng.integers(-2000, 2000, n)).astype(np.int16)
    ctx = clf.create_batch_context(4)                             # forces several calls per recording
    for overlap in (0.0, 1.5):
        got = clf.predict_pcm16_stream(ctx, pcm, overlap)
        ref_segs = io.chunk_audio(io.pcm16_to_f32(pcm), S, overlap, sr)
        assert len(got) == len(ref_segs) and len(got) >= 6
        ref = []
        for i in range(0, len(ref_segs), 4):
            ref += clf.predict_batch_with_context(ctx, [s for _, s in ref_segs[i:i + 4]])
        for (t0, r), (t1, _), rr in zip(got, ref_segs, ref):
            assert t0 == t1
            assert np.array_equal(r.raw_scores, rr.raw_scores)
            assert [(p.index, p.confidence) for p in r.predictions] == [(p.index, p.confidence) for p in rr.predictions]
    assert clf.predict_pcm16_stream(ctx, pcm, 3.0) == []          # step == 0 (birdnet-analyze.rs:721-724)
    with pytest.raises(bb.Error):
        clf.predict_pcm16_stream(ctx, np.zeros(0, dtype=np.int16))


@pytest.mark.gpu
def test_cfg5_day_of_audio_sharded_with_range_filter(v24_model_path, clf, v24_spec):
    """BASELINE config 5 at full size: 24 h of 48 kHz audio = 28,800 segments sharded over the visible GPUs
    (DevicePool: contiguous whole-batch blocks, host gather in caller order, no collective) with the fused range
    filter.  The oracle cannot run 28,800 segments in test time, so the full run is checked through
    size-independent properties: the stream cycles through 320 distinct synthetic segments, every repetition
    must reproduce the single-context result for that segment bit for bit, and the fused filter must equal the
    reference's post-filter (rangefilter.rs:527-579 semantics) on every one of them."""
    from birdnet_b200.multi_gpu import DevicePool
    from birdnet_b200.rangefilter import dense_range_state
    from oracle import postprocess_oracle as po
    n_total, n_distinct = 28_800, 320
    base = synth.batch(0, n_distinct, 144000, 48000)
    labels = clf.labels()
    loc_all = po.mock_embeddings(v24_spec.num_species, 5)
    loc = [bb.LocationScore(labels[i], float(loc_all[i]), i) for i in range(len(labels)) if loc_all[i] >= 0.0 or i % 2]
    thr = 0.01                                                    # rangefilter.rs:165 default
    # single-context ground truth for the 320 distinct segments (itself checked against the oracle elsewhere)
    ctx = clf.create_batch_context(64)
    plain = []
    for i in range(0, n_distinct, 64):
        plain += clf.predict_batch_with_context(ctx, list(base[i:i + 64]))
    rf = bb.RangeFilter.from_labels(labels, threshold=thr)
    want = rf.filter_batch_predictions([r.predictions for r in plain], loc, True)
    n_dev = _ffi.lib.bn_device_count()
    ids = list(range(n_dev)) if n_dev > 1 else [0, 0]
    pool = DevicePool(v24_model_path, ids, ctx_batch=256)
    pool.set_postprocess(5, 0.1)
    state, score = dense_range_state(labels, loc, thr)
    pool.set_range_filter(state, score, True)
    order = (np.arange(n_total) * 7) % n_distinct                 # 7 is coprime to 320: every segment repeats 90 times
    ref_logits = np.stack([r.raw_scores for r in plain])
    done = 0
    for lo in range(0, n_total, 3200):                            # 9 pool calls of 3,200 segments (12.5 batches each: ragged)
        sel = order[lo:lo + 3200]
        logits, emb, idx, conf, cnt = pool.run([base[j] for j in sel])
        assert emb is None and np.array_equal(logits, ref_logits[sel])
        for row, j in enumerate(sel):
            w = want[j]
            assert idx[row, :cnt[row]].tolist() == [p.index for p in w]
            assert np.allclose(conf[row, :cnt[row]], [p.confidence for p in w], atol=1e-7)
        done += len(sel)
    assert done == n_total


@pytest.mark.gpu
def test_pinned_segments_skip_the_gather_and_match(clf):
    """Segments that live in page-locked host memory (bn_host_alloc) are DMA'd in place; results are bit-identical
    to the gather path, for contiguous rows, permuted rows and a mix of pinned and pageable slices."""
    B = 24
    audio = synth.batch(3, B, 144000, 48000)
    ctx = clf.create_batch_context(B)
    ref = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, list(audio))])
    pinned = bb.pinned_array(audio.shape)
    pinned[:] = audio
    got = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, list(pinned))])
    assert np.array_equal(got, ref)
    perm = np.random.Generator(np.random.PCG64(1)).permutation(B)
    got_p = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, [pinned[j] for j in perm])])
    assert np.array_equal(got_p, ref[perm])
    mixed = [pinned[j] if j % 2 else audio[j] for j in range(B)]          # falls back to the gather path
    got_m = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, mixed)])
    assert np.array_equal(got_m, ref)
    assert np.array_equal(np.stack([r.raw_scores for r in clf.predict_batch(list(pinned[:5]))]), ref[:5])


def test_registered_caller_memory_takes_the_in_place_path(clf):
    """bn_host_register: memory the caller already owns, page-locked in place, is copied to the GPU without the gather
    (the literal pageable case of batch_context.rs:199-211 pays a host memcpy per segment); results are bit-identical."""
    B = 12
    audio = np.ascontiguousarray(synth.batch(5, B, 144000, 48000))
    ctx = clf.create_batch_context(B)
    ref = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, list(audio))])
    assert not ctx.last_run_in_place()                                     # plain numpy memory: gathered
    with bb.registered(audio) as a:
        got = np.stack([r.raw_scores for r in clf.predict_batch_with_context(ctx, list(a))])
        assert ctx.last_run_in_place()
        with pytest.raises(bb.Inference):                                   # the range is already registered
            bb.registered(audio).__enter__()
    assert np.array_equal(got, ref)
    clf.predict_batch_with_context(ctx, list(audio))
    assert not ctx.last_run_in_place()                                     # unregistered again
    pinned = bb.pinned_array(audio.shape)
    pinned[:] = audio
    clf.predict_batch_with_context(ctx, list(pinned))
    assert ctx.last_run_in_place()


def test_predict_batch_of_any_size_runs_in_chunks(clf):
    """bn_engine_run: 300 segments = one chunk of 256 + one of 44 on a pooled internal context; the result must be
    what the batch-context path gives for the same segments (classifier.rs:676-727 takes any batch)."""
    audio = synth.batch(0, 10, 144000, 48000)
    segs = [audio[i % 10] for i in range(300)]
    res = clf.predict_batch(segs)
    assert len(res) == 300
    ctx = clf.create_batch_context(16)
    ref = clf.predict_batch_with_context(ctx, [audio[i] for i in range(10)])
    for i, r in enumerate(res):
        q = ref[i % 10]
        assert [p.index for p in r.predictions] == [p.index for p in q.predictions]
        assert np.array_equal(r.raw_scores, q.raw_scores), i
    # many short-lived threads: the internal pool stays bounded (4 contexts), calls queue instead of allocating
    out = [None] * 12

    def work(t):
        out[t] = clf.predict_batch([audio[t % 10]] * 3)
    th = [threading.Thread(target=work, args=(t,)) for t in range(12)]
    [x.start() for x in th]
    [x.join() for x in th]
    for t in range(12):
        assert np.array_equal(out[t][0].raw_scores, ref[t % 10].raw_scores)


def test_activation_overflow_is_reported(v24_spec, v24_model_path, tmp_path):
    """hi/lo fp16 operands carry the fp16 RANGE: a model whose activations leave it yields NaN logits; the run is
    Ok (as a NaN would be in the reference) and bn_ctx_nonfinite_segments says how many segments were hit."""
    from birdnet_b200.modelgen import make_weights, write_model
    w = make_weights(v24_spec)
    for name in ("stem.weight", "s1b0.fused.weight", "s2b0.fused.weight"):
        w[name] = (w[name] * np.float32(2.0e3)).astype(np.float32)       # weights stay far below 65504
    path = os.path.join(str(tmp_path), "hot.onnx")
    write_model(v24_spec, path, w)
    c = (bb.Classifier.builder().model_path(path).labels(synthetic_labels(v24_spec.num_species)).top_k(5).build())
    ctx = c.create_batch_context(4)
    audio = synth.batch(0, 4, 144000, 48000)
    res = c.predict_batch_with_context(ctx, list(audio))
    assert len(res) == 4
    bad = sum(int(not np.isfinite(r.raw_scores).all()) for r in res)
    assert bad >= 1 and ctx.nonfinite_segments() == bad
    ok = (bb.Classifier.builder().model_path(v24_model_path)
          .labels(synthetic_labels(v24_spec.num_species)).build())
    ctx2 = ok.create_batch_context(4)
    ok.predict_batch_with_context(ctx2, list(audio))
    assert ctx2.nonfinite_segments() == 0
