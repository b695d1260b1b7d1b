#!/usr/bin/env python
"""Where the HOST time of a bench step goes: cProfile over a few device-resident steps of a bench workload
(`python tools/host_profile.py [workload]`), and the GPU-idle share (wall time of a step against the sum of its kernels'
phase timers)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from xesn_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "lazy"
ctx = bench.Ctx(torch, None, 1, 0, 0)
wl = bench.make_workload(ctx, name, dict(bench.WORKLOADS[name]), 1)
step = wl["step_device"]
for _ in range(3):
    step()
torch.cuda.synchronize()
n = 5
_lib.profile_enable(True)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(n):
    step()
torch.cuda.synchronize()
pr.disable()
wall = (time.perf_counter() - t0) / n * 1e3
prof = _lib.profile_read()
_lib.profile_enable(False)
busy = sum(v[1] for v in prof.values()) / n
print(f"{name}: {wall:.2f} ms per step on the wall clock, {busy:.2f} ms inside the library's phase timers")
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
