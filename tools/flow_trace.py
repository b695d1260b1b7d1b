#!/usr/bin/env python
"""Time of the dataflow forecast (predict_flow.cu) for the cfg-3 shape and, with the -DXESN_RECUR_TRACE build
(XESN_B200_LIB=.../libxesn_b200_trace.so), cycles per phase of one step as thread 0 of each CTA sees them."""
import ctypes
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import ESN_KW  # noqa: E402
from xesn_b200 import LazyESN, _lib  # noqa: E402
from xesn_b200.synthetic import lorenz96  # noqa: E402

N = int(os.environ.get("PROF_N", 2000))
G = int(os.environ.get("PROF_G", 16))
STEPS = int(os.environ.get("PROF_STEPS", 1000))
lib = _lib.load()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = LazyESN({"x": 16}, {"x": 1}, "periodic", N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
data = torch.from_numpy(lorenz96(16 * G, 2500 + STEPS + 60)).cuda()
esn.train(data[:, :2500])
test = data[:, 2500:]
for _ in range(2):
    esn._predict_core(test, STEPS, 50)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
esn._predict_core(test, STEPS, 50)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"flow forecast G={G} N={N}: {STEPS} steps (+50 spin-up) {ms:.3f} ms = {ms * 1e3 / STEPS:.2f} us per step")
if hasattr(lib, "xesn_debug_trace_flow"):
    buf = (ctypes.c_longlong * 512)()
    lib.xesn_debug_trace_flow(buf)
    last = STEPS % 256 or 256
    a = np.array(buf[:], dtype=np.float64).reshape(64, 8) / last
    io = a[32:]
    io = io[io.sum(1) > 0]
    a = a[:32]
    live = a[a.sum(1) > 0]
    print(f"input-polling warp of {len(io)} CTAs, cycles per step: [0] polling the inputs | [1] waiting for the owners at the "
          f"inputs barrier | [2] the owners' update | [3] contributions | total")
    for name, v in (("mean", io.mean(0)), ("min", io.min(0)), ("max", io.max(0))):
        print(f"   {name:5s} " + " ".join(f"{x:8.0f}" for x in v[:4]) + f" {v.sum():8.0f}")
    print(f"{len(live)} CTAs (last launch, {last} steps), cycles per step: [0] state refresh | [1] wait inputs/other warps | "
          f"[2] row-dot head | [3] row tails | [4] Win u | [5] tanh+leak | [6] store + wait owners | [7] contributions | total")
    for name, v in (("mean", live.mean(0)), ("min", live.min(0)), ("max", live.max(0))):
        print(f"   {name:5s} " + " ".join(f"{x:8.0f}" for x in v[:8]) + f" {v.sum():8.0f}")
