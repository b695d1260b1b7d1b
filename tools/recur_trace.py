#!/usr/bin/env python
"""Debug-build helper (library compiled with -DXESN_RECUR_TRACE): per-phase cycle breakdown of one
recurrence step, stand-alone kernel and inside the fused training kernel."""
import ctypes
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ESN_KW  # noqa: E402
from xesn_b200 import ESN, _lib  # noqa: E402

N = int(os.environ.get("PROF_N", 16000))
STEPS = int(os.environ.get("PROF_STEPS", 2048))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
res = esn._device_reservoir()
lib = res._lib


def show(fn, steps, label):
    buf = (ctypes.c_longlong * 512)()
    rc = getattr(lib, fn)(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(64, 8) / max(steps - 1, 1)
    live = a[a.sum(1) > 0]
    print(f"{label}: {len(live)} CTAs, cycles per step  [0]rest-of-compute(digits, ldg) | [1]wait warps | [2]own poll | [3]wait polls | [4]gather+dot | [5]tanh+leak | [6]state store | [7]- | total (rc={rc})")
    for name, v in (("mean", live.mean(0)), ("min", live.min(0)), ("max", live.max(0))):
        print(f"   {name:5s} " + " ".join(f"{x:8.0f}" for x in v) + f" {v.sum():8.0f}")


FULL = STEPS
STEPS = 512  # xesn_forced_run streams in chunks of 512 steps: one launch per call here
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
show("xesn_debug_trace", STEPS, "stand-alone")

STEPS = FULL
T = 3 * STEPS
ut = torch.randn((40, T), dtype=torch.float64, device="cuda") * 0.5
esn.train(ut, n_spinup=0, batch_size=STEPS)
torch.cuda.synchronize()
show("xesn_debug_trace_fused", STEPS, "fused (last chunk)")

buf = (ctypes.c_ulonglong * 640)()
lib.xesn_debug_cta_ns(buf)
both = np.array(buf[:], dtype=np.float64).reshape(2, 160, 2)[:, :148]
nrec = int(os.environ.get("XESN_B200_REC_CAP", 16))
a = both[0]
t0 = a[:, 0].min()
dur = (a[:, 1] - t0) / 1e6
print(f"fused mixed launch: recurrence CTAs end at {dur[:nrec].min():.2f}..{dur[:nrec].max():.2f} ms;"
      f" Gram CTAs end at {dur[nrec:].min():.2f}..{dur[nrec:].max():.2f} ms (mean {dur[nrec:].mean():.2f})")
a = both[1]
t0 = a[:, 0].min()
dur = (a[:, 1] - t0) / 1e6
print(f"fused Gram-only launch (148 CTAs): end at {dur.min():.2f}..{dur.max():.2f} ms (mean {dur.mean():.2f})")

# stand-alone int8 Gram kernel (256 threads): slope between two sizes removes the allocation + digitise cost
import time
S = torch.rand((4096, N), dtype=torch.float64, device="cuda") * 2 - 1
G = torch.zeros((N, N), dtype=torch.float64, device="cuda")
ts = {}
for K in (2048, 4096, 2048, 4096):
    torch.cuda.synchronize()
    t = time.perf_counter()
    _lib.check(lib.xesn_gram_i8(S.data_ptr(), N, K, N, G.data_ptr(), N, None))
    torch.cuda.synchronize()
    ts[K] = (time.perf_counter() - t) * 1e3
print(f"stand-alone gram_i8 (256 threads, 148 CTAs): {ts[4096] - ts[2048]:.2f} ms per 2048 steps (incl. ~0.2 ms digitise)")
