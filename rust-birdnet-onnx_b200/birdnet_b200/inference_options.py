"""InferenceOptions / CancellationToken (src/inference_options.rs:24-114).

The token owns a C int32 the engine polls while it waits for the device
(bn_run_opts.cancel_flag), the role `RunOptions::terminate()` plays in the reference's monitor
thread (src/classifier.rs:527-554).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional


class CancellationToken:
    def __init__(self):
        self._flag = ctypes.c_int32(0)

    def clone(self) -> "CancellationToken":
        t = CancellationToken.__new__(CancellationToken)
        t._flag = self._flag             # shared state, like Arc<AtomicBool>
        return t

    def cancel(self) -> None:
        self._flag.value = 1

    def is_cancelled(self) -> bool:
        return self._flag.value != 0


@dataclass
class InferenceOptions:
    timeout: Optional[float] = None                      # seconds (std::time::Duration)
    cancellation_token: Optional[CancellationToken] = None

    @staticmethod
    def new() -> "InferenceOptions":
        return InferenceOptions()

    @staticmethod
    def with_only_timeout(duration: float) -> "InferenceOptions":   # InferenceOptions::timeout(d)
        return InferenceOptions(timeout=duration)

    def with_timeout(self, duration: float) -> "InferenceOptions":
        return InferenceOptions(duration, self.cancellation_token)

    def with_cancellation_token(self, token: CancellationToken) -> "InferenceOptions":
        return InferenceOptions(self.timeout, token)

    def needs_monitor(self) -> bool:
        return self.timeout is not None or self.cancellation_token is not None
