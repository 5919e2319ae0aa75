"""Model files are deterministic, parse back to the same weights, load through the C ABI
(no GPU touched), and the torch oracle reproduces the committed golden vectors."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from birdnet_b200 import _ffi
from birdnet_b200.modelgen import build_model_bytes, get_spec, make_weights, parse_model, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v24_seed0.npz")


def test_onnx_round_trip_and_contract(v24_spec):
    w = make_weights(v24_spec)
    m = parse_model(build_model_bytes(v24_spec, w))
    assert m["opset"] == 17
    assert m["inputs"] == [("input", [-1, 144000])]            # batch_context.rs:221-223
    assert m["outputs"] == [("output", [-1, 6522])]            # batch_context.rs:248-250
    for k, v in w.items():
        assert np.array_equal(m["initializers"][k], v), k
    ops = {n["op"] for n in m["nodes"]}
    assert {"STFT", "ReduceMin", "ReduceMax", "Conv", "Gemm", "GlobalAveragePool", "Sigmoid"} <= ops
    assert "BatchNormalization" not in ops                     # folded, SURVEY appendix A.4


def test_model_bytes_are_deterministic(v24_spec):
    a = hashlib.sha256(build_model_bytes(v24_spec)).hexdigest()
    b = hashlib.sha256(build_model_bytes(get_spec("birdnet_v24"))).hexdigest()
    assert a == b
    assert a != hashlib.sha256(build_model_bytes(get_spec("birdnet_v24", seed=1))).hexdigest()


def test_other_families_declare_the_reference_contract():
    v30 = parse_model(build_model_bytes(get_spec("birdnet_v30", num_species=1000)))
    assert v30["inputs"] == [("input", [-1, 160000])]
    assert v30["outputs"] == [("output_0", [-1, 1024]), ("output_1", [-1, 1000])]   # batch_context.rs:252-262
    per = get_spec("perch_v2", num_species=500)
    rows = per.layer_table()
    assert per._shapes["spatial_embedding"] == (1536, 16, 4)                         # detection.rs:214-232
    outs = [tuple(o["shape"][1:]) for o in per.outputs]
    assert outs == [(1536,), (16, 4, 1536), (500, 128), (500,)]


def test_c_abi_parses_the_model_without_a_gpu(v24_model_path):
    info = _ffi.IoInfo()
    assert _ffi.lib.bn_model_inspect(v24_model_path.encode(), -1, C.byref(info)) == 0, _ffi.last_error()
    assert (info.model_type, info.sample_count, info.num_species, info.embedding_dim) == (0, 144000, 6522, 0)
    assert info.input.name == b"input" and info.input.shape() == [-1, 144000]
    assert info.n_outputs == 1 and info.outputs[0].name == b"output" and info.outputs[0].shape() == [-1, 6522]
    # override that contradicts the file -> reference's detection error text
    st = _ffi.lib.bn_model_inspect(v24_model_path.encode(), 1, C.byref(info))
    assert st == _ffi.BN_ERR_MODEL_DETECTION
    assert _ffi.last_error() == "model type BirdNetV30 expects 160000 samples, but model has 144000"


def test_synthetic_audio_is_pinned():
    g = np.load(GOLDEN)
    audio = synth.batch(0, 20, 144000, 48000)
    assert np.array_equal(np.frombuffer(hashlib.sha256(audio.tobytes()).digest(), dtype=np.uint8), g["audio_sha256"])
    assert audio.dtype == np.float32 and np.all(audio[9] == 0.0)          # silence segment
    assert abs(float(np.abs(audio[1]).max()) - 0.5) < 1e-3                # 440 Hz at 0.5 (integration_test.rs:57-67)


def test_model_file_is_pinned(v24_model_path):
    g = np.load(GOLDEN)
    h = hashlib.sha256(open(v24_model_path, "rb").read()).digest()
    assert np.array_equal(np.frombuffer(h, dtype=np.uint8), g["model_sha256"])


def test_oracle_reproduces_golden(v24_spec, v24_model_path):
    from oracle.model_oracle import ModelOracle, load_initializers
    from oracle import postprocess_oracle as po
    g = np.load(GOLDEN)
    audio = synth.batch(0, 6, 144000, 48000)
    out = ModelOracle(v24_spec, load_initializers(v24_model_path)).forward(audio)["output"]
    assert np.abs(out[:, ::32] - g["logits_every_32"][:6]).max() < 2e-3     # FP32 summation-order noise
    idx, conf, counts = po.top_k_batch(out, 5, 0.1)
    assert np.array_equal(counts, g["top5_count"][:6])
    assert np.abs(conf - g["top5_conf"][:6]).max() < 1e-3


def test_layer_table_totals(v24_spec):
    rows = v24_spec.layer_table()
    macs = sum(r["macs"] for r in rows)
    assert 0.55e9 < macs < 0.65e9 and v24_spec.frontend_macs() == 511 * 96 * (2048 + 1024)
    assert abs(sum(r["w_elems"] for r in rows) * 4 / 1e6 - 42.8) < 1.0     # "~50 MB weights" datapoint


# ------------------------------------------------------------------ 32 kHz families (row A9)
@pytest.mark.parametrize("fam,gold,mt,emb,nsp,outs", [
    ("birdnet_v30", "v30_seed0.npz", 1, 1024, 11560, [b"output_0", b"output_1"]),
    ("perch_v2", "perch_seed0.npz", 2, 1536, 14795, None),
])
def test_32k_families_parse_and_reproduce_golden(fam, gold, mt, emb, nsp, outs):
    from birdnet_b200.modelgen.make_models import ensure_model
    from oracle.model_oracle import ModelOracle, load_initializers
    from oracle import postprocess_oracle as po
    path = ensure_model(fam)
    spec = get_spec(fam)
    info = _ffi.IoInfo()
    assert _ffi.lib.bn_model_inspect(path.encode(), -1, C.byref(info)) == 0, _ffi.last_error()
    assert (info.model_type, info.sample_count, info.num_species, info.embedding_dim) == (mt, 160000, nsp, emb)
    assert info.sample_rate == 32000 and info.segment_duration == 5.0          # types.rs:17-31
    if outs is not None:                                                        # batch_context.rs:252-262
        assert [info.outputs[i].name for i in range(info.n_outputs)] == outs
    else:                                                                       # detection.rs:214-232
        assert info.n_outputs == 4
        assert [info.outputs[i].shape()[1:] for i in range(4)] == [[1536], [16, 4, 1536], [500, 128], [14795]]
    g = np.load(os.path.join(os.path.dirname(GOLDEN), gold))
    h = hashlib.sha256(open(path, "rb").read()).digest()
    assert np.array_equal(np.frombuffer(h, dtype=np.uint8), g["model_sha256"])
    audio = synth.batch(0, 3, 160000, 32000)
    logits, e = ModelOracle(spec, load_initializers(path)).logits_and_embeddings(audio)
    assert np.abs(logits[:, ::32] - g["logits_every_32"][:3]).max() < 2e-3
    assert np.abs(e[:, ::8] - g["emb_every_8"][:3]).max() < 1e-3
    _, conf, counts = po.top_k_batch(logits, 5, 0.1)
    assert np.array_equal(counts, g["top5_count"][:3])
    assert np.abs(conf - g["top5_conf"][:3]).max() < 1e-3
