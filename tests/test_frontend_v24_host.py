"""Host logic of the fused v2.4 front-end kernel (csrc/frontend_v24.cu), checked on CPU: the K-step schedule (cell-column
pair outer, hop-block inner, zero-weight pad step), the block-Toeplitz patch addressing the kernel's UMMA descriptors use
and the packed hi / lo basis must add up to the plain dot product of a frame with the basis."""
import ctypes as C

import pytest

from birdnet_b200 import _ffi


def _check(n_fft, hop, n_mels, samples):
    fn = _ffi.lib.bn_debug_spec_v24_selfcheck
    fn.restype = C.c_double
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    out = (C.c_int * 3)()
    return fn(n_fft, hop, n_mels, samples, out), list(out)


@pytest.mark.parametrize("n_fft,hop,ksteps", [(2048, 278, 134), (1024, 280, 66)])
def test_v24_branches_schedule_equals_direct_sum(n_fft, hop, ksteps):
    worst, (n_k, slots, smem) = _check(n_fft, hop, 96, 144000)
    assert worst >= 0 and worst < 1e-9, worst          # same hi + lo weights on both sides: only double rounding differs
    assert n_k == ksteps and n_k % 2 == 0              # two K steps per ring slot
    assert slots >= 2 and smem <= 224 * 1024


@pytest.mark.parametrize("n_fft,hop,n_mels", [(512, 128, 64), (800, 200, 80), (1024, 256, 128), (400, 160, 40), (2048, 300, 96)])
def test_other_shapes_schedule_equals_direct_sum(n_fft, hop, n_mels):
    worst, (n_k, slots, smem) = _check(n_fft, hop, n_mels, 48000)
    assert worst >= 0 and worst < 1e-9, (worst, n_k, slots, smem)


def test_shapes_outside_the_kernel_are_refused():
    assert _check(1024, 2048, 96, 144000)[0] == -1.0   # hop longer than the frame
    assert _check(1024, 256, 200, 144000)[0] == -1.0   # more mel bands than two accumulator sets hold
    assert _check(2048, 512, 96, 144000)[0] == -1.0    # a 512-sample hop: the sample patch would not fit shared memory
