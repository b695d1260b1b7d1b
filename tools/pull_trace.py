#!/usr/bin/env python
"""Debug build (-DXESN_RECUR_TRACE): cycle breakdown of one pull-model recurrence step (thread 0 of 64 CTAs)."""
import ctypes, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ESN_KW
from xesn_b200 import ESN, _lib

N = int(os.environ.get("PROF_N", 16000))
STEPS = 512
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
res = esn._device_reservoir()
lib = res._lib
_lib.check(lib.xesn_set_recur_mode(res.handle, 2))
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
buf = (ctypes.c_longlong * 512)()
lib.xesn_debug_trace_fused(buf)
a = np.array(buf[:], dtype=np.float64).reshape(64, 8) / STEPS
print("cycles per step: [0]stores+digits+drive | [1]head poll | [2]dot+tails | [3]tanh+leak | [4]hint wait | - - | [7] extra poll rounds")
for name, v in (("mean", a.mean(0)), ("min", a.min(0)), ("max", a.max(0))):
    print(f"   {name:5s} " + " ".join(f"{x:8.1f}" for x in v) + f"   total {v[:4].sum():8.0f}")
