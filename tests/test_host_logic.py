"""Host-side mirror of the reference API: the reference's own unit tests restated
(src/labels.rs:125-359, src/detection.rs:177-285, src/error.rs:134-250,
src/inference_options.rs:117-199, src/types.rs:188-294, src/rangefilter.rs:589-700,
src/batch_context.rs:342-360, src/classifier.rs builder tests)."""
import ctypes as C

import numpy as np
import pytest

import birdnet_b200 as bb
from birdnet_b200 import _ffi, errors
from birdnet_b200.labels import (load_labels_from_file, parse_csv_labels, parse_json_labels,
                                 parse_labels, parse_text_labels)
from birdnet_b200.rangefilter import dense_range_state
from birdnet_b200.types import LabelFormat, LocationScore, ModelType


# ------------------------------------------------------------------ labels.rs
def test_text_labels():
    assert parse_text_labels("American Robin\nNorthern Cardinal\n\nBlue Jay\n") == \
        ["American Robin", "Northern Cardinal", "Blue Jay"]
    assert parse_text_labels("  American Robin  \n  Northern Cardinal  ") == ["American Robin", "Northern Cardinal"]
    assert parse_text_labels("Species 1\n\nSpecies 2\n\n\nSpecies 3") == ["Species 1", "Species 2", "Species 3"]
    assert parse_text_labels("Pingüino Emperador\n鸟类\nПтица\n🐦") == ["Pingüino Emperador", "鸟类", "Птица", "🐦"]
    assert parse_text_labels("   \n\t\n  \n") == []
    assert parse_text_labels("a\r\nb\r\n") == ["a", "b"]


def test_csv_labels():
    assert parse_csv_labels("American Robin\nNorthern Cardinal\nBlue Jay") == \
        ["American Robin", "Northern Cardinal", "Blue Jay"]
    assert parse_csv_labels("label,scientific_name\nAmerican Robin,Turdus migratorius\n"
                            "Northern Cardinal,Cardinalis cardinalis") == ["American Robin", "Northern Cardinal"]
    assert parse_csv_labels("species\nAmerican Robin\nNorthern Cardinal") == ["American Robin", "Northern Cardinal"]
    assert parse_csv_labels("inat2024_fsd50k\nAmerican Robin\nNorthern Cardinal") == ["American Robin", "Northern Cardinal"]
    assert parse_csv_labels("dataset_fsd50k\nAmerican Robin\nNorthern Cardinal") == ["American Robin", "Northern Cardinal"]
    assert parse_csv_labels("label,scientific\nSpecies 1,Name1,Extra\nSpecies 2,Name2") == ["Species 1", "Species 2"]
    assert parse_csv_labels("label\n\nSpecies 1\n\nSpecies 2") == ["Species 1", "Species 2"]
    assert parse_csv_labels('label\n"Species, with comma"\n"Species with ""quotes"""\nSpecies normal') == \
        ["Species, with comma", 'Species with "quotes"', "Species normal"]


def test_json_labels():
    assert parse_json_labels('["American Robin", "Northern Cardinal", "Blue Jay"]') == \
        ["American Robin", "Northern Cardinal", "Blue Jay"]
    assert parse_json_labels('{"labels": ["American Robin", "Northern Cardinal"]}') == ["American Robin", "Northern Cardinal"]
    assert parse_json_labels('[{"name": "American Robin"}, {"name": "Northern Cardinal"}]') == ["American Robin", "Northern Cardinal"]
    assert parse_json_labels('[{"label": "American Robin"}, {"label": "Northern Cardinal"}]') == ["American Robin", "Northern Cardinal"]
    assert parse_json_labels('[{"species": "American Robin"}, {"species": "Northern Cardinal"}]') == ["American Robin", "Northern Cardinal"]
    assert parse_json_labels("[]") == []
    assert parse_json_labels('["Pingüino", "鸟类", "Птица"]') == ["Pingüino", "鸟类", "Птица"]
    assert parse_json_labels('[{"name": "Species 1"}, {"other": "Species 2"}]') == ["Species 1"]
    for bad in ('{"invalid": "format"}', '{"data": {"labels": ["Species 1"]}}'):
        with pytest.raises(errors.LabelParse):
            parse_json_labels(bad)


def test_parse_by_format_and_missing_file():
    assert len(parse_labels("American Robin\nNorthern Cardinal", LabelFormat.Text)) == 2
    assert len(parse_labels('["American Robin", "Northern Cardinal"]', LabelFormat.Json)) == 2
    with pytest.raises(errors.LabelLoad) as e:
        load_labels_from_file("/nonexistent/path.txt", ModelType.BirdNetV24)
    assert "failed to load labels" in str(e.value)


def test_reference_label_files_if_present():
    """The real label assets define num_species."""
    import os
    p = "/root/reference/data/labels"
    if not os.path.isdir(p):
        pytest.skip("reference tree not mounted (GPU box)")
    txt = [f for f in os.listdir(os.path.join(p, "birdnet_v2.4")) if f.endswith(".txt")][0]
    assert len(load_labels_from_file(os.path.join(p, "birdnet_v2.4", txt), ModelType.BirdNetV24)) == 6522
    assert len(load_labels_from_file(os.path.join(p, "perch_v2", "labels.csv"), ModelType.PerchV2)) == 14795


# ------------------------------------------------------------------ detection.rs (through the C ABI)
def _detect(inp, outs, override=-1):
    info = _ffi.IoInfo()
    ind = (C.c_int64 * len(inp))(*inp)
    flat = [d for o in outs for d in o]
    od = (C.c_int64 * max(len(flat), 1))(*flat)
    ranks = (C.c_int32 * max(len(outs), 1))(*[len(o) for o in outs])
    st = _ffi.lib.bn_detect_model_type(ind, len(inp), od, ranks, len(outs), override, C.byref(info))
    return st, info


def test_detect_v24_v30_perch():
    st, i = _detect([1, 144000], [[1, 6522]])
    assert st == 0 and (i.model_type, i.sample_rate, i.segment_duration, i.sample_count, i.num_species, i.embedding_dim) == \
        (0, 48000, 3.0, 144000, 6522, 0)
    st, i = _detect([1, 160000], [[1, 1024], [1, 1000]])
    assert st == 0 and (i.model_type, i.sample_rate, i.segment_duration, i.num_species, i.embedding_dim) == (1, 32000, 5.0, 1000, 1024)
    st, i = _detect([1, 160000], [[1, 1536], [1, 16, 4, 1536], [1, 500, 128], [1, 14795]])
    assert st == 0 and (i.model_type, i.num_species, i.embedding_dim) == (2, 14795, 1536)
    st, i = _detect([-1, 1, 144000], [[-1, 6522]])          # [batch, 1, samples], dynamic batch
    assert st == 0 and i.sample_count == 144000


def test_detect_override_and_errors():
    st, i = _detect([1, 160000], [[1, 512], [1, 16, 4, 512], [1, 500, 128], [1, 500]], override=2)
    assert st == 0 and (i.model_type, i.embedding_dim, i.num_species) == (2, 512, 500)
    st, _ = _detect([1, 160000], [[1, 1024], [1, 1000]], override=0)
    assert st == _ffi.BN_ERR_MODEL_DETECTION
    assert _ffi.last_error() == "model type BirdNetV24 expects 144000 samples, but model has 160000"
    st, _ = _detect([1, 100000], [[1, 1000]])
    assert st == _ffi.BN_ERR_MODEL_DETECTION
    assert _ffi.last_error() == "unsupported model: 100000 samples, 1 outputs (expected 144000/1, 160000/2, or 160000/4)"
    st, _ = _detect([1, 160000], [[1, 1024]], override=1)
    assert st == _ffi.BN_ERR_MODEL_DETECTION and _ffi.last_error() == "`BirdNET` v3.0 expects 2 outputs, got 1"
    st, _ = _detect([144000], [[1, 6522]])
    assert st == _ffi.BN_ERR_MODEL_DETECTION and _ffi.last_error() == "unexpected input shape: [144000]"
    st, _ = _detect([1, -1], [[1, 6522]])
    assert st == _ffi.BN_ERR_MODEL_DETECTION and _ffi.last_error() == "invalid sample count: -1"
    st, _ = _detect([1, 144000], [[]])
    assert st == _ffi.BN_ERR_MODEL_DETECTION and _ffi.last_error() == "empty output shape"


# ------------------------------------------------------------------ error.rs
def test_error_display_strings():
    assert str(errors.InputSize(144000, 100000)) == "input size mismatch: expected 144000 samples, got 100000"
    assert str(errors.BatchInputSize(3, 144000, 50000)) == \
        "batch input size mismatch: segment 3 has 50000 samples, expected 144000"
    assert str(errors.ModelDetection("unsupported model")) == "model detection failed: unsupported model"
    assert str(errors.LabelCount(6522, 1000)) == "label count mismatch: model expects 6522, got 1000"
    assert str(errors.AudioFormat("WAV must be mono")) == "unsupported audio format: WAV must be mono"
    assert str(errors.AudioRead("/path/to/file.wav", "file not found")) == \
        "failed to read audio file /path/to/file.wav: file not found"
    s = str(errors.InvalidCoordinates(95.0, 200.0, "latitude out of range"))
    assert "latitude: 95" in s and "longitude: 200" in s
    s = str(errors.InvalidDate(13, 32, "month out of range"))
    assert "month: 13" in s and "day: 32" in s and "month out of range" in s
    assert str(errors.RangeFilterInference("model invoke failed")) == "range filter inference failed: model invoke failed"
    assert str(errors.Timeout(30.0)) == "inference timed out after 30s"
    assert str(errors.Timeout(1.5)) == "inference timed out after 1.5s"
    assert str(errors.Timeout(0.1)) == "inference timed out after 100ms"
    assert str(errors.Cancelled()) == "inference was cancelled"
    assert str(errors.ModelPathRequired()) == "model path required"
    assert str(errors.LabelsRequired()) == "labels required (provide path or vec)"
    assert str(errors.Inference("batch size 5 exceeds context max 4")) == "inference failed: batch size 5 exceeds context max 4"


# ------------------------------------------------------------------ inference_options.rs
def test_cancellation_token_and_options():
    t = bb.CancellationToken()
    assert not t.is_cancelled()
    c = t.clone()
    c.cancel()
    assert t.is_cancelled() and c.is_cancelled()
    o = bb.InferenceOptions()
    assert o.timeout is None and o.cancellation_token is None and not o.needs_monitor()
    assert bb.InferenceOptions.with_only_timeout(30.0).needs_monitor()
    o = bb.InferenceOptions.new().with_timeout(5.0).with_cancellation_token(bb.CancellationToken())
    assert o.timeout == 5.0 and o.cancellation_token is not None and o.needs_monitor()
    assert bb.InferenceOptions.new().with_cancellation_token(bb.CancellationToken()).needs_monitor()


# ------------------------------------------------------------------ types.rs
def test_model_type_constants():
    assert [m.sample_rate() for m in ModelType] == [48000, 32000, 32000]
    assert [m.segment_duration() for m in ModelType] == [3.0, 5.0, 5.0]
    assert [m.sample_count() for m in ModelType] == [144000, 160000, 160000]
    assert [m.has_embeddings() for m in ModelType] == [False, True, True]
    assert ModelType.BirdNetV24.expected_label_format() is LabelFormat.Text
    assert ModelType.PerchV2.expected_label_format() is LabelFormat.Csv
    assert bb.available_execution_providers()[0] is bb.ExecutionProviderInfo.Cpu   # tests/execution_provider_test.rs


# ------------------------------------------------------------------ rangefilter.rs calendar / validators
def test_calendar_and_validators():
    assert bb.calculate_week(1, 1) == 1.0 and bb.calculate_week(1, 8) == 2.0
    assert bb.calculate_week(2, 1) == 5.0 and bb.calculate_week(12, 31) == 49.0
    for lat, lon in ((45.0, -122.0), (0.0, 0.0), (-90.0, -180.0), (90.0, 180.0)):
        bb.validate_coordinates(lat, lon)
    with pytest.raises(errors.InvalidCoordinates) as e:
        bb.validate_coordinates(91.0, 0.0)
    assert "latitude must be in range [-90, 90], got 91" in str(e.value)
    with pytest.raises(errors.InvalidCoordinates):
        bb.validate_coordinates(0.0, 181.0)
    for m, d in ((1, 1), (6, 15), (12, 31)):
        bb.validate_date(m, d)
    for m, d in ((0, 1), (13, 1), (1, 0), (1, 32)):
        with pytest.raises(errors.InvalidDate):
            bb.validate_date(m, d)


def test_builders_required_fields():
    with pytest.raises(errors.ModelPathRequired):          # rangefilter.rs:689-693, classifier tests
        bb.RangeFilter.builder().build()
    with pytest.raises(errors.LabelsRequired):
        bb.RangeFilter.builder().model_path("/tmp/model.onnx").build()
    with pytest.raises(errors.ModelPathRequired):
        bb.Classifier.builder().build()
    with pytest.raises(errors.LabelsRequired):
        bb.Classifier.builder().model_path("/tmp/model.onnx").build()
    with pytest.raises((errors.ModelLoad, errors.RuntimeInit)):
        bb.Classifier.builder().model_path("/nonexistent/model.onnx").labels(["a"]).build()


def test_dense_range_state_is_string_keyed():
    labels = ["A", "B", "C", "A"]                         # duplicate label strings share one map entry
    loc = [LocationScore("A", 0.9, 0), LocationScore("B", 0.02, 1), LocationScore("A", 0.5, 0)]
    state, score = dense_range_state(labels, loc, 0.03)
    assert state.tolist() == [1, 2, 0, 1]
    assert np.allclose(score, [0.5, 0.02, 0.0, 0.5])      # last duplicate wins, like HashMap::collect


def test_result_blocks_are_reused_only_when_unreferenced():
    """raw_scores / embeddings rows are views of a per-call block (classifier.rs:897-903 hands out owned Vecs); a block
    goes back into circulation only when the caller has dropped every view of it."""
    import threading
    from birdnet_b200.classifier import Classifier
    c = Classifier.__new__(Classifier)
    c._pool, c._pool_lock = [], threading.Lock()
    addr = lambda a: a.__array_interface__["data"][0]
    a = c._owned_block(4, 10)
    first, row = addr(a), a[2]
    b = c._owned_block(4, 10)
    assert addr(b) != first                       # a is alive
    del a
    assert addr(c._owned_block(4, 10)) != first   # a row of it is alive
    del row
    assert addr(c._owned_block(4, 10)) == first   # nothing refers to it any more
    assert addr(c._owned_block(5, 10)) != first   # other shapes get their own
    for _ in range(40):                           # bounded
        keep = [c._owned_block(2, 2) for _ in range(3)]
    assert len(c._pool) <= Classifier._BLOCK_POOL_MAX
