"""birdnet_b200 — B200-native drop-in for the batched inference path of
tphakala/rust-birdnet-onnx.  Public names mirror the reference's re-exports (src/lib.rs:93-111).
"""
from .errors import (AudioFormat, AudioRead, BatchInputSize, Cancelled, Error, Inference,  # noqa: F401
                     InputSize, InvalidCoordinates, InvalidDate, LabelCount, LabelLoad, LabelParse,
                     LabelsRequired, ModelDetection, ModelLoad, ModelPathRequired,
                     RangeFilterInference, RuntimeInit, Timeout)
from .types import (ExecutionProviderInfo, LabelFormat, LocationScore, ModelConfig, ModelType,  # noqa: F401
                    Prediction, PredictionResult, available_execution_providers)
from .inference_options import CancellationToken, InferenceOptions  # noqa: F401
from .classifier import BatchInferenceContext, Classifier, ClassifierBuilder, pinned_array  # noqa: F401
from .rangefilter import (RangeFilter, RangeFilterBuilder, calculate_week,  # noqa: F401
                          validate_coordinates, validate_date)
