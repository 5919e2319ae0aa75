"""Range-filter meta model: RangeFilter::predict (src/rangefilter.rs:435-502) on the
device, and the dense range mask built there for the fused epilogue."""
import ctypes as C
import os

import numpy as np
import pytest

import birdnet_b200 as bb
from birdnet_b200 import _ffi
from birdnet_b200.modelgen import meta_model as mm
from birdnet_b200.modelgen import synth
from birdnet_b200.modelgen.make_models import synthetic_labels
from oracle import meta_oracle as mo

SCORE_TOL = 2e-6      # f32 dot products of <= 256 terms in a different summation order, then a sigmoid in (0, 1)
PLACES = [(60.17, 24.94, 6, 15), (-33.87, 151.21, 12, 31), (0.0, 0.0, 1, 1), (90.0, -180.0, 2, 8), (-90.0, 180.0, 7, 22)]


GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meta_seed0.npz"))


def _weights(path):
    from oracle.model_oracle import load_initializers
    return load_initializers(path)


# ------------------------------------------------------------------ CPU: generator + oracle + load-time checks
def test_meta_model_file_contract(tmp_path):
    from birdnet_b200.modelgen.onnx_writer import parse_model
    raw = mm.build_meta_model_bytes(50, seed=3)
    assert raw == mm.build_meta_model_bytes(50, seed=3)          # deterministic
    m = parse_model(raw)
    w = m["initializers"]
    assert w["fc0.weight"].shape == (128, 3) and w["fc2.weight"].shape == (50, 256) and w["in_scale"].shape == (3,)
    ref = mm.meta_weights(50, seed=3)
    assert all(np.array_equal(w[k], ref[k]) for k in ref)


def test_meta_oracle_week_and_selection():
    # calculate_week known answers (rangefilter.rs:589-627)
    assert [mo.calculate_week(*d) for d in ((1, 1), (1, 8), (2, 1), (12, 31))] == [1.0, 2.0, 5.0, 49.0]
    w = mm.meta_weights(300, seed=1)
    labels = [f"sp{i}" for i in range(300)]
    s = mo.forward(w, 10.0, 20.0, mo.calculate_week(6, 17))
    assert s.dtype == np.float32 and s.shape == (300,) and (s > 0).all() and (s < 1).all()
    assert (s < 0.01).any() and (s >= 0.01).any()            # the stand-in exercises both sides of the default threshold
    out = mo.predict(w, labels, 0.01, 10.0, 20.0, 6, 17)       # week 23
    assert [i for _, _, i in out] == sorted(np.nonzero(s >= np.float32(0.01))[0], key=lambda i: -s[i])
    assert all(a[1] >= b[1] for a, b in zip(out, out[1:]))
    assert all(lbl == labels[i] for lbl, _, i in out)
    # fewer labels than outputs: the `i < labels.len()` guard (rangefilter.rs:487)
    assert all(i < 100 for _, _, i in mo.predict(w, labels[:100], 0.0, 10.0, 20.0, 6, 17))


def test_meta_create_rejects_bad_models_before_touching_a_device(v24_model_path, tmp_path):
    h = C.c_void_p()
    assert _ffi.lib.bn_meta_create(None, 0, C.byref(h)) == _ffi.BN_ERR_MODEL_PATH_REQUIRED
    assert _ffi.lib.bn_meta_create(b"/nonexistent.onnx", 0, C.byref(h)) == _ffi.BN_ERR_MODEL_LOAD
    # a classifier graph is not a meta model: input is [B, 144000], not [1, 3]
    assert _ffi.lib.bn_meta_create(v24_model_path.encode(), 0, C.byref(h)) == _ffi.BN_ERR_MODEL_DETECTION
    assert not h.value
    # two outputs -> "meta model expects 1 output, got 2" (rangefilter.rs:254-258)
    from birdnet_b200.modelgen.make_models import ensure_model
    st = _ffi.lib.bn_meta_create(ensure_model("birdnet_v30").encode(), 0, C.byref(h))
    assert st == _ffi.BN_ERR_MODEL_DETECTION and "meta model expects 1 output, got 2" in _ffi.last_error()


def test_range_filter_builder_required_fields_and_no_cpu_fallback(has_gpu):
    with pytest.raises(bb.ModelPathRequired):
        bb.RangeFilter.builder().labels(["a"]).build()
    with pytest.raises(bb.LabelsRequired):
        bb.RangeFilter.builder().model_path("x.onnx").build()
    if not has_gpu:
        with pytest.raises(bb.RuntimeInit):                        # never computes on the host
            bb.RangeFilter.builder().model_path(mm.ensure_meta_model(64, 0)).labels(["x"] * 64).build()
    rf = bb.RangeFilter.from_labels(["a"])
    with pytest.raises(bb.RangeFilterInference):
        rf.predict(0.0, 0.0, 1, 1)


def test_meta_oracle_against_committed_golden():
    """tests/golden/meta_seed0.npz (tests/golden/make_golden_meta.py) pins the stand-in model file and the oracle."""
    import hashlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "meta_seed0.npz"))
    path = mm.ensure_meta_model(6522, 0)
    assert hashlib.sha256(open(path, "rb").read()).digest() == bytes(g["model_sha256"])
    w = _weights(path)
    for i, (lat, lon, month, day) in enumerate(g["places"]):
        s = mo.forward(w, np.float32(lat), np.float32(lon), np.float32(mo.calculate_week(int(month), int(day))))
        assert np.abs(s[::8] - g["scores_every_8"][i]).max() <= 1e-6
        assert int(s.argmax()) == int(g["argmax"][i])
        assert abs(int((s >= np.float32(0.01)).sum()) - int(g["n_above_default_threshold"][i])) <= 2


# ------------------------------------------------------------------ GPU: parity through the C ABI
@pytest.fixture(scope="module")
def meta_path():
    return mm.ensure_meta_model(6522, 0)


@pytest.fixture(scope="module")
def rf(meta_path):
    return bb.RangeFilter.builder().model_path(meta_path).labels(synthetic_labels(6522)).threshold(0.01).build()


@pytest.mark.gpu
def test_meta_scores_match_oracle(meta_path):
    w = _weights(meta_path)
    h = C.c_void_p()
    assert _ffi.lib.bn_meta_create(meta_path.encode(), 0, C.byref(h)) == 0, _ffi.last_error()
    try:
        assert _ffi.lib.bn_meta_num_outputs(h) == 6522
        got = np.empty(6522, dtype=np.float32)
        for lat, lon, month, day in PLACES:
            week = mo.calculate_week(month, day)
            assert _ffi.lib.bn_meta_predict(h, lat, lon, week, got.ctypes.data_as(C.POINTER(C.c_float)), 6522) == 0
            ref = mo.forward(w, np.float32(lat), np.float32(lon), np.float32(week))
            assert np.abs(got - ref).max() <= SCORE_TOL
            gi = [tuple(p) for p in GOLD["places"]].index((lat, lon, month, day))
            assert np.abs(got[::8] - GOLD["scores_every_8"][gi]).max() <= SCORE_TOL      # committed fixture
        short = np.empty(10, dtype=np.float32)
        assert _ffi.lib.bn_meta_predict(h, 0.0, 0.0, 1.0, short.ctypes.data_as(C.POINTER(C.c_float)), 10) != 0
    finally:
        _ffi.lib.bn_meta_destroy(h)


@pytest.mark.gpu
def test_predict_matches_oracle_selection_and_order(rf, meta_path):
    w = _weights(meta_path)
    labels = synthetic_labels(6522)
    for lat, lon, month, day in PLACES:
        got = rf.predict(lat, lon, month, day)
        s = mo.forward(w, np.float32(lat), np.float32(lon), np.float32(mo.calculate_week(month, day)))
        thr = np.float32(0.01)
        sure_in = {int(i) for i in np.nonzero(s >= thr + SCORE_TOL)[0]}
        maybe = {int(i) for i in np.nonzero(np.abs(s - thr) < SCORE_TOL)[0]}
        got_idx = {p.index for p in got}
        assert sure_in <= got_idx <= sure_in | maybe
        assert all(p.species == labels[p.index] and abs(p.score - float(s[p.index])) <= SCORE_TOL for p in got)
        assert all(a.score >= b.score for a, b in zip(got, got[1:]))            # sorted by score descending
        assert all(p.score >= thr for p in got)
        assert 0 < len(got) < 6522


@pytest.mark.gpu
def test_predict_validation_and_label_count(rf, meta_path):
    with pytest.raises(bb.InvalidCoordinates):
        rf.predict(91.0, 0.0, 6, 15)
    with pytest.raises(bb.InvalidCoordinates):
        rf.predict(0.0, -181.0, 6, 15)
    with pytest.raises(bb.InvalidDate):
        rf.predict(0.0, 0.0, 13, 1)
    with pytest.raises(bb.InvalidDate):
        rf.predict(0.0, 0.0, 1, 0)
    with pytest.raises(bb.LabelCount) as e:                                        # rangefilter.rs:260-266
        bb.RangeFilter.builder().model_path(meta_path).labels(["a", "b"]).build()
    assert "6522" in str(e.value) and "2" in str(e.value)


@pytest.mark.gpu
def test_device_built_mask_equals_host_filter(rf, v24_model_path, v24_spec):
    """install_on (forward pass + dense tri-state on the device) gives the predictions that the reference's sequence
    predict() -> filter_batch_predictions() gives, both with the predict threshold and with a stricter filter threshold
    (the drop arm, rangefilter.rs:353-372)."""
    labels = synthetic_labels(v24_spec.num_species)
    clf = bb.Classifier.builder().model_path(v24_model_path).labels(labels).top_k(10).build()
    audio = list(synth.batch(0, 12, 144000, 48000))
    plain = clf.predict_batch(audio)
    lat, lon, month, day = PLACES[0]
    loc = rf.predict(lat, lon, month, day)
    for filter_thr, rerank in ((None, False), (None, True), (0.2, False), (0.2, True)):
        thr = rf.threshold() if filter_thr is None else filter_thr
        want = bb.RangeFilter.from_labels(labels, threshold=thr).filter_batch_predictions(
            [r.predictions for r in plain], loc, rerank)
        rf.install_on(clf, lat, lon, month, day, filter_threshold=filter_thr, rerank=rerank)
        fused = clf.predict_batch(audio)
        clf.clear_range_filter()
        assert any(len(w) < len(r.predictions) for w, r in zip(want, plain)) or filter_thr is None
        for w, f in zip(want, fused):
            assert [(p.index, p.species) for p in f.predictions] == [(p.index, p.species) for p in w]
            assert np.allclose([p.confidence for p in f.predictions], [p.confidence for p in w], atol=1e-7)
    with pytest.raises(bb.RangeFilterInference):
        other = bb.Classifier.builder().model_path(v24_model_path).labels([s + "x" for s in labels]).build()
        rf.install_on(other, lat, lon, month, day)
