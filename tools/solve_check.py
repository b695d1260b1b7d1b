#!/usr/bin/env python
"""Accuracy of the ridge solve at BASELINE cfg-2 size (N=16000, beta=1e-6, Lorenz-96): int8 trailing updates
vs the FP64 DMMA path; residual of the normal equations for both."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import ESN_KW
from xesn_b200 import ESN
from xesn_b200.synthetic import lorenz96

N = int(os.environ.get("PROF_N", 16000))
T = int(os.environ.get("PROF_T", 20000))
u = lorenz96(40, T)
out = {}
for mode in ("1", "0"):
    os.environ["XESN_B200_SOLVE_I8"] = mode
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = ESN(40, 40, N, 0.5, 1e-6, **{k: dict(v) for k, v in ESN_KW.items()})
        esn.build()
    esn.train(u, n_spinup=0)
    out[mode] = np.array(esn.Wout)
    print(f"SOLVE_I8={mode}: |Wout|max {np.abs(out[mode]).max():.4e}  info {esn.last_info.cpu().tolist()}", flush=True)
d = np.abs(out["1"] - out["0"]).max() / np.abs(out["0"]).max()
print(f"max-normalised difference int8-solve vs FP64-solve: {d:.3e}")

# independent check: rebuild A = R^T R + beta I and rhs = Y R with cuBLAS FP64, solve with cuSOLVER, and compare
import ctypes
from xesn_b200 import _lib
lib = _lib.load()
res = esn._device_reservoir()
u_dev = torch.from_numpy(np.ascontiguousarray(u)).cuda()
carry = torch.zeros((1, N), dtype=torch.float64, device="cuda")
states = torch.empty((1, T + 1, N), dtype=torch.float64, device="cuda")
series = _lib.Series(u_dev.data_ptr(), u_dev.stride(0), None, None, None)
nb = lib.xesn_forced_scratch_bytes(res.handle, 1, T, 0)
scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
_lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, T, 1, carry.data_ptr(), states.data_ptr(),
                               scratch.data_ptr(), nb, None))
torch.cuda.synchronize()
del scratch
R = states[0, :T]
A = R.T @ R
A.diagonal().add_(1e-6)
rhs = u_dev @ R
del states, R
for mode in ("1", "0"):
    W = torch.from_numpy(out[mode]).cuda()
    resid = float((W @ A - rhs).abs().max() / rhs.abs().max())
    print(f"SOLVE_I8={mode}: normal-equation residual max|Wout A - rhs| / max|rhs| = {resid:.3e}")
Lc = torch.linalg.cholesky(A)
Wc = torch.cholesky_solve(rhs.T.contiguous(), Lc).T
print(f"cuSOLVER Cholesky: residual {float((Wc @ A - rhs).abs().max() / rhs.abs().max()):.3e}")
for mode in ("1", "0"):
    W = torch.from_numpy(out[mode]).cuda()
    print(f"SOLVE_I8={mode}: max-normalised difference to cuSOLVER: {float((W - Wc).abs().max() / Wc.abs().max()):.3e}")
