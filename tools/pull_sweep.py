#!/usr/bin/env python
"""A/B of the recurrence engines at BASELINE cfg-2 size: stand-alone forced run (us per step) and the fused
training chunk (ms per 2048-step launch) for the team model and the pull model with polling knobs."""
import ctypes, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ESN_KW
from xesn_b200 import ESN, _lib
from xesn_b200.synthetic import lorenz96

N = int(os.environ.get("PROF_N", 16000))
STEPS = 2048
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
res = esn._device_reservoir()
lib = res._lib
u = lorenz96(40, 3 * STEPS)
ud = torch.as_tensor(u[:, :STEPS], device="cuda").contiguous()
series = _lib.Series(ud.data_ptr(), ud.stride(0), None, None, None)
nb = lib.xesn_forced_scratch_bytes(res.handle, 1, STEPS, 0)
scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
states = torch.empty((1, STEPS + 1, N), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None


def forced():
    carry = torch.zeros((1, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, STEPS, 1, carry.data_ptr(), states.data_ptr(),
                                   scratch.data_ptr(), nb, None))


def run(tag, mode, env):
    global ref
    for k in ("XESN_B200_NO_HINT", "XESN_B200_POLL_SLEEP1", "XESN_B200_PULL_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    _lib.check(lib.xesn_set_recur_mode(res.handle, mode))
    forced(); torch.cuda.synchronize()
    e0.record(); forced(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / STEPS
    print(f"   [{tag}] forced done: {us:.2f} us/step, status {res.status() if hasattr(res, 'status') else '?'}", flush=True)
    if ref is None:
        ref = states.clone()
    err = float((states - ref).abs().max())
    fused = None
    if not env.get("XESN_B200_PULL_CTAS"):
        esn.train(u, n_spinup=0)
        _lib.profile_enable(True)
        esn.train(u, n_spinup=0)
        torch.cuda.synchronize()
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        fused = {k: (v[0], round(v[1], 2)) for k, v in pr.items() if k == "fused_chunk"}
    print(f"{tag:28s} forced {us:6.2f} us/step  maxdiff {err:.1e}  fused(3 chunks) {fused}", flush=True)


run("team", 0, {})
run("pull+hint", 2, {})
run("pull no hint", 2, {"XESN_B200_NO_HINT": "1"})
for ctas in (74, 111):
    run(f"pull+hint ctas={ctas}", 2, {"XESN_B200_PULL_CTAS": str(ctas)})
