#!/usr/bin/env python
"""One training pass in the pull-model fused pipeline at a chosen size; prints time and status."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bench import ESN_KW
from xesn_b200 import ESN, _lib
from xesn_b200.synthetic import lorenz96
N = int(os.environ.get("PROF_N", 2000)); T = int(os.environ.get("PROF_T", 768)); B = int(os.environ.get("PROF_B", 256))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()}); esn.build()
u = lorenz96(40, T)
res = esn._device_reservoir()
_lib.check(res._lib.xesn_set_recur_mode(res.handle, 0))
esn.train(u, n_spinup=0, batch_size=B); w0 = np.array(esn.Wout)
_lib.check(res._lib.xesn_set_recur_mode(res.handle, 2))
t0 = time.time(); esn.train(u, n_spinup=0, batch_size=B); torch.cuda.synchronize()
print(os.environ.get("TAG", ""), f"N={N} T={T} B={B}: pull train {time.time()-t0:.3f} s, Wout maxdiff vs team {np.abs(esn.Wout-w0).max():.2e}", flush=True)

import ctypes
if hasattr(res._lib, "xesn_debug_trace_fused") or True:
    try:
        buf = (ctypes.c_longlong * 512)()
        res._lib.xesn_debug_trace_fused(buf)
        a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
        print("late hints per CTA (first 8):", a[:8, 5], " last (hint, s_sync0):", [(int(x) >> 32, int(x) & 0xffffffff) for x in a[:8, 6]])
        print("cycles/step cols 0..4:", (a[:4, :5] / B).round(0))
    except Exception as e:
        print("no trace:", e)
