#!/usr/bin/env python
"""Debug build (-DXESN_RECUR_TRACE): cycle breakdown of one step of the barrier-free forecast kernel."""
import ctypes, os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ESN_KW
from xesn_b200 import ESN, _lib
from xesn_b200.synthetic import lorenz96
N = int(os.environ.get("PROF_N", 16000)); STEPS = 256
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()}); esn.build()
u = lorenz96(40, 1400)
esn.train(u[:, :1000], n_spinup=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    torch.cuda.synchronize(); e0.record()
    esn.predict(u[:, 1000:], n_steps=STEPS, n_spinup=100)
    e1.record(); torch.cuda.synchronize()
print(f"predict(100 spin-up + {STEPS} steps): {e0.elapsed_time(e1):.2f} ms")
buf = (ctypes.c_longlong * 512)()
try:
    esn._device_reservoir()._lib.xesn_debug_trace_predict(buf)
except AttributeError:
    sys.exit(0)  # not a trace build
a = np.array(buf[:], dtype=np.float64).reshape(64, 8) / STEPS
print("cycles per step: [0]hint wait | [1]gathers issue + contributions | [2]input vector | [3]gather wait + dot | [4]Win u, tanh, store | [5]readout part")
for name, v in (("mean", a.mean(0)), ("min", a.min(0)), ("max", a.max(0))):
    print(f"   {name:5s} " + " ".join(f"{x:8.1f}" for x in v[:5]) + f"   total {v[:5].sum():8.0f}   | thread 128: v poll {v[5]:8.1f}, hint wait {v[6]:8.1f} cycles, {v[7]:6.1f} spins")
