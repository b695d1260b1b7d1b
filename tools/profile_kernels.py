#!/usr/bin/env python
"""Small driver for ncu captures of the two hot kernels at BASELINE cfg-2 size:
one persistent recurrence launch (N=16000, 256 steps) and one Gram SYRK launch (N=16000, K=512)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ESN_KW  # noqa: E402
from xesn_b200 import ESN, _lib  # noqa: E402

N = int(os.environ.get("PROF_N", 16000))
STEPS = int(os.environ.get("PROF_STEPS", 256))
K = int(os.environ.get("PROF_K", 512))
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
res = esn._device_reservoir()
lib = res._lib
u = torch.randn((40, STEPS), dtype=torch.float64, device="cuda")
carry = torch.zeros((1, N), dtype=torch.float64, device="cuda")
states = torch.empty((1, STEPS + 1, N), dtype=torch.float64, device="cuda")
series = _lib.Series(u.data_ptr(), u.stride(0), None, None, None)
nb = lib.xesn_forced_scratch_bytes(res.handle, 1, STEPS, 0)
scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
for _ in range(2):
    _lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, STEPS, 1, carry.data_ptr(), states.data_ptr(),
                                   scratch.data_ptr(), nb, None))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, STEPS, 1, carry.data_ptr(), states.data_ptr(),
                               scratch.data_ptr(), nb, None))
e1.record()
torch.cuda.synchronize()
print(f"forced_run N={N} steps={STEPS}: {e0.elapsed_time(e1) * 1e3 / STEPS:.2f} us/step (incl. drive GEMM + memset)")

R = states[0, :K].contiguous()
G = torch.zeros((N, N), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.check(lib.xesn_gemm_f64(R.data_ptr(), N, R.data_ptr(), N, G.data_ptr(), N, N, N, K, 1.0, 1.0, 4, None))
torch.cuda.synchronize()
e0.record()
_lib.check(lib.xesn_gemm_f64(R.data_ptr(), N, R.data_ptr(), N, G.data_ptr(), N, N, N, K, 1.0, 1.0, 4, None))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"gram N={N} K={K}: {ms:.3f} ms, {N * (N + 1) * K / ms / 1e9:.2f} TFLOP/s algorithmic")
