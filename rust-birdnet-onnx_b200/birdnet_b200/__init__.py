"""birdnet_b200 — B200-native drop-in for the batched inference path of
tphakala/rust-birdnet-onnx.  Public names mirror the reference's re-exports (src/lib.rs:93-111).

Names resolve lazily (PEP 562): touching any product name imports ``_ffi``, which maps
``lib/libbirdnet_b200.so`` and raises if it is missing (there is no CPU fallback).  The
``modelgen`` sub-package (graph specs, ONNX writer, synthetic audio) never touches the library, so
the CPU reference arm of ``bench.py`` and the oracle can use it without mapping the product ``.so``.
"""
import importlib

_EXPORTS = {
    "errors": ("AudioFormat", "AudioRead", "BatchInputSize", "Cancelled", "Error", "Inference", "InputSize",
               "InvalidCoordinates", "InvalidDate", "LabelCount", "LabelLoad", "LabelParse", "LabelsRequired",
               "ModelDetection", "ModelLoad", "ModelPathRequired", "RangeFilterInference", "RuntimeInit", "Timeout"),
    "types": ("ExecutionProviderInfo", "LabelFormat", "LocationScore", "ModelConfig", "ModelType", "Prediction",
              "PredictionResult", "available_execution_providers"),
    "inference_options": ("CancellationToken", "InferenceOptions"),
    "classifier": ("BatchInferenceContext", "Classifier", "ClassifierBuilder", "pinned_array", "registered"),
    "rangefilter": ("RangeFilter", "RangeFilterBuilder", "calculate_week", "validate_coordinates", "validate_date"),
}
_WHERE = {name: mod for mod, names in _EXPORTS.items() for name in names}
__all__ = sorted(_WHERE)


def __getattr__(name):
    mod = _WHERE.get(name)
    if mod is None:
        raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
    value = getattr(importlib.import_module(f".{mod}", __name__), name)
    globals()[name] = value
    return value


def __dir__():
    return sorted(list(globals()) + __all__)
