"""RangeFilter (src/rangefilter.rs): calendar, validators, prediction filter / rerank.

`filter_predictions` runs on the device through bn_range_filter_apply; species are matched by
label *string* like the reference's HashMap (rangefilter.rs:340-343): distinct strings of the
call are numbered densely and the kernel works on those ids.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi
from .errors import (InvalidCoordinates, InvalidDate, LabelsRequired, ModelPathRequired,
                     RangeFilterInference, raise_for_status, _rust_f32)
from .types import LocationScore, Prediction

_lib = _ffi.lib
DEFAULT_THRESHOLD = 0.01                  # rangefilter.rs:165


def calculate_week(month: int, day: int) -> float:        # rangefilter.rs:77-81
    return float((month - 1) * 4 + (day - 1) // 7 + 1)


def validate_coordinates(latitude: float, longitude: float) -> None:   # rangefilter.rs:91-107
    if not (-90.0 <= latitude <= 90.0):
        raise InvalidCoordinates(latitude, longitude,
                                 f"latitude must be in range [-90, 90], got {_rust_f32(latitude)}")
    if not (-180.0 <= longitude <= 180.0):
        raise InvalidCoordinates(latitude, longitude,
                                 f"longitude must be in range [-180, 180], got {_rust_f32(longitude)}")


def validate_date(month: int, day: int) -> None:          # rangefilter.rs:117-133
    if not (1 <= month <= 12):
        raise InvalidDate(month, day, f"month must be in range [1, 12], got {month}")
    if not (1 <= day <= 31):
        raise InvalidDate(month, day, f"day must be in range [1, 31], got {day}")


def dense_range_state(labels: Sequence[str], location_scores: Sequence[LocationScore], threshold: float):
    """Per-class tri-state: 0 = absent from the map (keep unchanged), 1 = score >= threshold
    (keep, x score when reranking), 2 = score < threshold (drop).  rangefilter.rs:346-378."""
    m: Dict[str, float] = {}
    for s in location_scores:
        m[s.species] = s.score           # HashMap collect: the last duplicate wins
    thr = np.float32(threshold)
    state = np.zeros(len(labels), dtype=np.uint8)
    score = np.zeros(len(labels), dtype=np.float32)
    for i, lab in enumerate(labels):
        v = m.get(lab)
        if v is None:
            continue
        v32 = np.float32(v)
        score[i] = v32
        state[i] = 1 if v32 >= thr else 2     # NaN score -> not >= -> drop, like `Some(_)` arm
    return state, score


def _apply_on_device(batch: List[List[Prediction]], location_scores, threshold, rerank, engine=None):
    rows = len(batch)
    stride = max((len(p) for p in batch), default=0)
    if rows == 0:
        return []
    if stride == 0:
        return [[] for _ in batch]
    ids: Dict[str, int] = {}
    names: List[str] = []
    pin = (_ffi.Pred * (rows * stride))()
    cin = (C.c_uint32 * rows)()
    orig_index: Dict[int, Dict[int, int]] = {}
    for r, preds in enumerate(batch):
        cin[r] = len(preds)
        for j, p in enumerate(preds):
            sid = ids.get(p.species)
            if sid is None:
                sid = ids[p.species] = len(names)
                names.append(p.species)
            pin[r * stride + j].index = sid
            pin[r * stride + j].confidence = p.confidence
    state, score = dense_range_state(names, location_scores, threshold)
    pout = (_ffi.Pred * (rows * stride))()
    cout = (C.c_uint32 * rows)()
    # the model index of a species string (first occurrence per row keeps its own index below)
    st = _lib.bn_range_filter_apply(engine, pin, cin, rows, stride,
                                    state.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    score.ctypes.data_as(C.POINTER(C.c_float)), len(names),
                                    1 if rerank else 0, pout, cout)
    raise_for_status(st)
    out = []
    for r, preds in enumerate(batch):
        # species -> original model indices in list order (strings may repeat with other indices)
        pending: Dict[int, List[int]] = {}
        for p in preds:
            pending.setdefault(ids[p.species], []).append(p.index)
        row = []
        for j in range(cout[r]):
            q = pout[r * stride + j]
            lst = pending[q.index]
            row.append(Prediction(names[q.index], float(q.confidence), lst.pop(0) if len(lst) > 1 else lst[0]))
        out.append(row)
    return out


class RangeFilterBuilder:                                   # rangefilter.rs:144-277
    def __init__(self):
        self._model_path: Optional[str] = None
        self._labels: Optional[List[str]] = None
        self._labels_path: Optional[str] = None
        self._threshold = DEFAULT_THRESHOLD
        self._device_id = 0

    def model_path(self, path: str) -> "RangeFilterBuilder":
        self._model_path = str(path)
        return self

    def labels_path(self, path: str) -> "RangeFilterBuilder":
        self._labels_path, self._labels = str(path), None
        return self

    def labels(self, labels: List[str]) -> "RangeFilterBuilder":
        self._labels, self._labels_path = list(labels), None
        return self

    def from_classifier_labels(self, labels: Sequence[str]) -> "RangeFilterBuilder":
        return self.labels(list(labels))

    def threshold(self, threshold: float) -> "RangeFilterBuilder":
        self._threshold = float(threshold)
        return self

    def device_id(self, device_id: int) -> "RangeFilterBuilder":
        self._device_id = device_id
        return self

    def build(self) -> "RangeFilter":
        if self._model_path is None:
            raise ModelPathRequired()
        if self._labels is None and self._labels_path is None:
            raise LabelsRequired()
        from .labels import parse_text_labels
        labels = self._labels
        if labels is None:
            from .errors import LabelLoad
            try:
                with open(self._labels_path, "r", encoding="utf-8") as f:
                    labels = parse_text_labels(f.read())
            except OSError as e:
                raise LabelLoad(self._labels_path, str(e))
        # load the meta model (rangefilter.rs:239-249), then the one-output and label-count checks (251-266)
        handle = C.c_void_p()
        raise_for_status(_ffi.lib.bn_meta_create(self._model_path.encode(), int(self._device_id), C.byref(handle)))
        expected = int(_ffi.lib.bn_meta_num_outputs(handle))
        if len(labels) != expected:
            _ffi.lib.bn_meta_destroy(handle)
            from .errors import LabelCount
            raise LabelCount(expected, len(labels))
        return RangeFilter(self._model_path, labels, self._threshold, self._device_id, handle)


class RangeFilter:                                          # rangefilter.rs:389-579
    def __init__(self, model_path: Optional[str], labels: List[str], threshold: float, device_id: int = 0, handle=None):
        self._model_path = model_path
        self._labels = labels
        self._threshold = threshold
        self._device_id = device_id
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _ffi.lib.bn_meta_destroy(h)
            except Exception:       # interpreter shutdown
                pass

    @staticmethod
    def builder() -> RangeFilterBuilder:
        return RangeFilterBuilder()

    @staticmethod
    def from_labels(labels: Sequence[str], threshold: float = DEFAULT_THRESHOLD) -> "RangeFilter":
        """Filter-only instance (no meta model): enough for filter_predictions*."""
        return RangeFilter(None, list(labels), threshold)

    def threshold(self) -> float:
        return self._threshold

    def predict(self, latitude: float, longitude: float, month: int, day: int) -> List[LocationScore]:
        validate_coordinates(latitude, longitude)          # rangefilter.rs:443
        validate_date(month, day)                          # rangefilter.rs:446
        if not self._h:
            raise RangeFilterInference("this RangeFilter was built without a meta model (from_labels)")
        week = calculate_week(month, day)                  # rangefilter.rs:449
        n = len(self._labels)
        scores = np.empty(n, dtype=np.float32)
        st = _ffi.lib.bn_meta_predict(self._h, float(latitude), float(longitude), week,
                                      scores.ctypes.data_as(C.POINTER(C.c_float)), n)
        if st != 0:
            raise RangeFilterInference(_ffi.last_error())
        thr = np.float32(self._threshold)
        keep = np.nonzero(scores >= thr)[0]                # rangefilter.rs:482-496 (NaN >= thr is false, like Rust)
        # sort_unstable_by(|a, b| b.score.total_cmp(&a.score)) (499); equal scores keep index order here
        order = keep[np.argsort(-scores[keep], kind="stable")]
        return [LocationScore(self._labels[i], float(scores[i]), int(i)) for i in order]

    def install_on(self, classifier, latitude: float, longitude: float, month: int, day: int,
                   filter_threshold: Optional[float] = None, rerank: bool = False) -> None:
        """predict() + Classifier.set_range_filter() without leaving the device: the scores of this location are turned
        into the dense per-class mask on the GPU and installed as the classifier's fused range filter (labels of both
        must be the same list)."""
        validate_coordinates(latitude, longitude)
        validate_date(month, day)
        if not self._h:
            raise RangeFilterInference("this RangeFilter was built without a meta model (from_labels)")
        if list(classifier.labels()) != list(self._labels):
            raise RangeFilterInference("install_on needs identical label lists on the classifier and the range filter")
        ft = self._threshold if filter_threshold is None else filter_threshold
        raise_for_status(_ffi.lib.bn_meta_install_range_filter(
            self._h, classifier._h, float(latitude), float(longitude), calculate_week(month, day),
            float(self._threshold), float(ft), 1 if rerank else 0))

    def filter_predictions(self, predictions: Sequence[Prediction],
                           location_scores: Sequence[LocationScore], rerank: bool) -> List[Prediction]:
        return _apply_on_device([list(predictions)], location_scores, self._threshold, rerank)[0]

    def filter_batch_predictions(self, predictions_batch: Sequence[Sequence[Prediction]],
                                 location_scores: Sequence[LocationScore], rerank: bool) -> List[List[Prediction]]:
        return _apply_on_device([list(p) for p in predictions_batch], location_scores,
                                self._threshold, rerank)
