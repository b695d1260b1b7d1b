#!/usr/bin/env python
"""Train on a short series at BASELINE cfg-2 size so that the ridge solve dominates (for ncu launch lists)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import ESN_KW
from xesn_b200 import ESN
from xesn_b200.synthetic import lorenz96

N = int(os.environ.get("PROF_N", 16000))
T = int(os.environ.get("PROF_T", 2048))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
    esn.build()
u = lorenz96(40, T)
for _ in range(int(os.environ.get("PROF_REPS", 1))):
    esn.train(u, n_spinup=0)
torch.cuda.synchronize()
print("done")
