#!/usr/bin/env python
"""Quick check + timing of the int8 tensor-core Gram against an FP64 torch reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from xesn_b200 import _lib

lib = _lib.load()
torch.manual_seed(0)
for (K, N) in [(32, 128), (64, 192), (96, 300), (512, 1000), (2048, 4000)]:
    S = (torch.rand((K, N), dtype=torch.float64, device="cuda") * 2 - 1)
    S[0, 0] = 1.0
    S[1, 1] = -1.0
    G = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.xesn_gram_i8(S.data_ptr(), N, K, N, G.data_ptr(), N, None))
    ref = S.T @ S
    low = torch.tril(torch.ones_like(ref)).bool()
    err = float((G - ref)[low].abs().max())
    up = float(G[~low].abs().max()) if N > 1 else 0.0
    print(f"K={K} N={N}: max abs err (lower) {err:.3e}  (scale {float(ref.abs().max()):.1f}), upper touched max {up:.3e}", flush=True)
    # exactness: states on a 2^-20 grid -> Gram exactly representable, must match bit for bit
    Sq = torch.round(S * 2**20) / 2**20
    G.zero_()
    _lib.check(lib.xesn_gram_i8(Sq.data_ptr(), N, K, N, G.data_ptr(), N, None))
    qi = torch.round(Sq * 2**20).to(torch.int64)
    exact = ((qi.cpu().T @ qi.cpu()).to(torch.float64) / 2.0**40).cuda() if K * N < 3e6 else (Sq.T @ Sq)
    print(f"   grid inputs: max abs diff vs exact {float((G - exact)[low].abs().max()):.3e}", flush=True)
N, K = 16000, 2048
S = (torch.rand((K, N), dtype=torch.float64, device="cuda") * 2 - 1)
G = torch.zeros((N, N), dtype=torch.float64, device="cuda")
_lib.check(lib.xesn_gram_i8(S.data_ptr(), N, K, N, G.data_ptr(), N, None))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
_lib.check(lib.xesn_gram_i8(S.data_ptr(), N, K, N, G.data_ptr(), N, None))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"N={N} K={K}: {ms:.2f} ms incl. digitise+alloc -> {N*(N+1)*K/ms/1e9:.1f} TFLOP/s FP64-equivalent")
