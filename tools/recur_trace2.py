#!/usr/bin/env python
"""Debug-build helper (library compiled with -DXESN_RECUR_TRACE): cycles per phase of one team-model recurrence
step inside the fused training kernel, for the multi-reservoir (PROF_KIND=lazy: 16 groups x N=2000) and the
small single-reservoir (PROF_KIND=eager, PROF_N=1000) configurations."""
import ctypes
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ESN_KW  # noqa: E402
from xesn_b200 import ESN, LazyESN, _lib  # noqa: E402
from xesn_b200.synthetic import lorenz96  # noqa: E402

KIND = os.environ.get("PROF_KIND", "lazy")
N = int(os.environ.get("PROF_N", 2000))
STEPS = int(os.environ.get("PROF_STEPS", 2048))
CHUNKS = int(os.environ.get("PROF_CHUNKS", 3))   # 1: the traced launch runs the recurrence role alone (no Gram role beside it)
lib = _lib.load()


def show(fn, steps, label):
    buf = (ctypes.c_longlong * 512)()
    rc = getattr(lib, fn)(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(64, 8) / max(steps - 1, 1)
    live = a[a.sum(1) > 0]
    print(f"{label}: {len(live)} CTAs, cycles per step  [0]digits+ldg | [1]wait warps | [2]own poll | [3]wait polls | "
          f"[4]gather+dot | [5]tanh+leak | [6]state store | [7]- | total (rc={rc})")
    for name, v in (("mean", live.mean(0)), ("min", live.min(0)), ("max", live.max(0))):
        print(f"   {name:5s} " + " ".join(f"{x:8.0f}" for x in v) + f" {v.sum():8.0f}")


with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    if KIND == "lazy":
        G = int(os.environ.get("PROF_G", 16))
        esn = LazyESN({"x": 16}, {"x": 1}, "periodic", N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
        data = torch.from_numpy(lorenz96(16 * G, CHUNKS * STEPS)).cuda()
    else:
        esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
        data = torch.from_numpy(lorenz96(40, CHUNKS * STEPS)).cuda()
    esn.build()
for _ in range(2):
    esn.train(data, n_spinup=0, batch_size=STEPS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.profile_enable(True)
e0.record()
esn.train(data, n_spinup=0, batch_size=STEPS)
e1.record()
torch.cuda.synchronize()
prof = _lib.profile_read()
launch_ms = _lib.profile_regions("fused_chunk")
_lib.profile_enable(False)
print(f"{KIND} N={N}: train of {CHUNKS * STEPS} steps {e0.elapsed_time(e1):.2f} ms; fused kernels {prof['fused_chunk'][1]:.2f} ms "
      f"= {prof['fused_chunk'][1] * 1e3 / (CHUNKS * STEPS):.3f} us per step")
print("fused launches, ms: " + " ".join(f"{x:.3f}" for x in launch_ms))
show("xesn_debug_trace_fused", STEPS, f"fused {KIND} N={N} (last chunk{', recurrence role alone' if CHUNKS == 1 else ''})")
