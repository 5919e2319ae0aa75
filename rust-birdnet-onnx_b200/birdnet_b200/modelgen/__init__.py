"""Graph specs, seeded weights, ONNX writer and synthetic audio for the build-authored models."""
from .graphspec import GraphSpec, get_spec, make_weights  # noqa: F401
from .onnx_writer import build_model_bytes, parse_model, write_model  # noqa: F401
