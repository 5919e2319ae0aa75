"""Pins oracle/postprocess_oracle.c against the reference's own known-answer tests
(src/postprocess.rs:102-331, src/rangefilter.rs:589-916, src/testutil.rs:150-265)."""
import math

import numpy as np
import pytest

from oracle import postprocess_oracle as po


def test_sigmoid():                                    # postprocess.rs:102-106
    assert abs(po.sigmoid(0.0) - 0.5) < 1e-4
    assert po.sigmoid(10.0) > 0.99
    assert po.sigmoid(-10.0) < 0.01


def test_top_k_basic():                                # postprocess.rs:109-121
    p = po.top_k_predictions([0.1, 0.5, 0.9, 0.3, 0.7], 3, None)
    assert len(p) == 3
    assert p[0][1] >= p[1][1] >= p[2][1]
    assert p[0][0] == 2


def test_top_k_with_min_confidence():                  # postprocess.rs:124-135
    p = po.top_k_predictions([-5.0, 0.0, 5.0], 10, 0.4)
    assert len(p) == 2 and all(c >= 0.4 for _, c in p)


def test_top_k_larger_than_input():                    # postprocess.rs:138-145
    assert len(po.top_k_predictions([0.1, 0.2], 100, None)) == 2


def test_top_k_empty_and_zero_k():                     # postprocess.rs:148-159
    assert po.top_k_predictions([], 10, None) == []
    assert po.top_k_predictions([0.1, 0.2, 0.3], 0, None) == []


def test_indices():                                    # postprocess.rs:163-172
    p = po.top_k_predictions([0.1, 0.9, 0.5], 3, None)
    assert p[0][0] == 1


def test_sigmoid_special_values():                     # postprocess.rs:177-203
    assert abs(po.sigmoid(float("inf")) - 1.0) < 1.2e-7
    assert abs(po.sigmoid(float("-inf"))) < 1.2e-7
    assert math.isnan(po.sigmoid(float("nan")))
    assert po.sigmoid(100.0) > 0.9999 and po.sigmoid(-100.0) < 0.0001


def test_all_equal_scores():                           # postprocess.rs:206-216
    p = po.top_k_predictions([0.5] * 4, 2, None)
    assert len(p) == 2 and abs(p[0][1] - p[1][1]) < 1e-4


def test_negative_logits():                            # postprocess.rs:219-230
    p = po.top_k_predictions([-10.0, -5.0, -1.0, -20.0], 2, None)
    assert len(p) == 2 and p[0][1] >= p[1][1] and p[0][0] == 2


def test_nan_values():                                 # postprocess.rs:233-242
    assert len(po.top_k_predictions([1.0, float("nan"), 2.0, 0.5], 3, None)) > 0


def test_min_confidence_bounds():                      # postprocess.rs:245-266
    assert len(po.top_k_predictions([-10.0, 0.0, 10.0], 10, 0.0)) == 3
    assert len(po.top_k_predictions([-10.0, 0.0, 10.0], 10, 1.0)) == 0


def test_top_k_max_usize():                            # postprocess.rs:269-277
    assert len(po.top_k_predictions([0.1, 0.2, 0.3], 2 ** 64 - 1, None)) == 3


def test_missing_labels_indices_cover_all():           # postprocess.rs:280-295
    p = po.top_k_predictions([0.1, 0.2, 0.3, 0.4], 4, None)
    assert sorted(i for i, _ in p) == [0, 1, 2, 3]     # 2 and 3 would print as unknown_2/3


def test_top_k_matches_sort_on_lcg_logits():
    # testutil.rs:110-121 generator; distinct-enough values -> top-k == argsort
    for seed in (1, 42, 12345):
        lg = po.random_logits(6522, seed)
        p = po.top_k_predictions(lg, 10, None)
        order = np.argsort(-lg, kind="stable")
        assert sorted(lg[[i for i, _ in p]].tolist(), reverse=True) == sorted(lg[order[:10]].tolist(), reverse=True)
        confs = [c for _, c in p]
        assert confs == sorted(confs, reverse=True)


def test_random_logits_range_and_determinism():        # testutil.rs:216-243
    a, b = po.random_logits(100, 42), po.random_logits(100, 42)
    assert np.array_equal(a, b) and a.min() >= -5.0 and a.max() <= 5.0
    assert not np.array_equal(a, po.random_logits(100, 43))
    e = po.mock_embeddings(1024, 42)
    assert e.shape == (1024,) and e.min() >= 0.0 and e.max() <= 1.0


def test_calculate_week():                             # rangefilter.rs:589-627
    assert po.calculate_week(1, 1) == 1.0
    assert po.calculate_week(1, 8) == 2.0
    assert po.calculate_week(2, 1) == 5.0
    assert po.calculate_week(12, 31) == 49.0


def test_filter_above_threshold():                     # rangefilter.rs:703-751
    preds = [("Species A", 0.8, 0), ("Species B", 0.3, 1), ("Species C", 0.05, 2)]
    loc = [("Species A", 0.9), ("Species B", 0.02), ("Species C", 0.5)]
    f = po.filter_predictions(preds, loc, 0.03, False)
    assert [s for s, _, _ in f] == ["Species A", "Species C"]


def test_filter_with_rerank():                         # rangefilter.rs:754-813
    preds = [("Species A", 0.9, 0), ("Species B", 0.8, 1), ("Species C", 0.7, 2)]
    loc = [("Species A", 0.5), ("Species B", 0.9), ("Species C", 0.6)]
    f = po.filter_predictions(preds, loc, 0.03, True)
    assert [s for s, _, _ in f] == ["Species B", "Species A", "Species C"]
    for (_, c, _), want in zip(f, (0.72, 0.45, 0.42)):
        assert abs(c - want) < 1e-3


def test_filter_species_not_in_meta_model():           # rangefilter.rs:816-864
    preds = [("Species A", 0.8, 0), ("Species B", 0.7, 1), ("Species D", 0.9, 3)]
    loc = [("Species A", 0.9), ("Species C", 0.8)]
    f = po.filter_predictions(preds, loc, 0.03, False)
    assert [(s, i) for s, _, i in f] == [("Species A", 0), ("Species B", 1), ("Species D", 3)]
    assert [round(c, 6) for _, c, _ in f] == [0.8, 0.7, 0.9]


def test_filter_batch():                               # rangefilter.rs:880-916
    loc = [("Species A", 0.9), ("Species B", 0.05)]
    r = [po.filter_predictions(b, loc, 0.1, False) for b in ([("Species A", 0.8, 0)], [("Species B", 0.6, 1)])]
    assert [len(x) for x in r] == [1, 0]
