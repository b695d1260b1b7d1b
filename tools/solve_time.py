#!/usr/bin/env python
"""Time of the ridge-solve phase alone at BASELINE cfg-2 size (short training series, library event timers)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ESN_KW
from xesn_b200 import ESN, _lib
from xesn_b200.synthetic import lorenz96

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, int(os.environ.get("PROF_N", 16000)), 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
u = lorenz96(40, 2048)
for i in range(4):
    if i == 2:
        _lib.profile_enable(True)
    esn.train(u, n_spinup=0)
torch.cuda.synchronize()
print(os.environ.get("TAG", ""), {k: round(v[1] / 2, 2) for k, v in _lib.profile_read().items() if v[1] > 0.5})
