#!/usr/bin/env python
"""Tiny pass over every persistent kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
fused training chunk (recurrence role beside the int8 and the FP64 Gram role), pull-model spin-up, ridge solve,
barrier-free forecast (one reservoir), persistent forecast loop (several reservoirs), per-step fallback.
    XESN_B200_POLL_TIMEOUT_MS=600000 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are small because the sanitizer slows the polling kernels by orders of magnitude."""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xesn_b200 import ESN, LazyESN  # noqa: E402
from xesn_b200.synthetic import lorenz96  # noqa: E402

KW = dict(
    input_kwargs=dict(factor=0.5, distribution="uniform", normalization="svd", random_seed=0),
    adjacency_kwargs=dict(factor=0.9, distribution="uniform", normalization="svd", is_sparse=True,
                          connectedness=5, format="csr", random_seed=1),
    bias_kwargs=dict(factor=0.5, distribution="uniform", random_seed=2))


def main():
    steps = int(os.environ.get("SAN_STEPS", "96"))
    data = lorenz96(8, steps + 60)
    for gram in ("i8", "f64"):
        os.environ["XESN_B200_GRAM"] = gram
        esn = ESN(8, 8, 260, 0.5, 1e-4, **{k: dict(v) for k, v in KW.items()})
        esn.build()
        esn.train(data[:, :steps], n_spinup=8, batch_size=40)
        pred = esn.predict(data[:, steps:], n_steps=6, n_spinup=10)
        assert np.isfinite(pred).all() and np.isfinite(esn.Wout).all()
        print(f"eager gram={gram}: ok", flush=True)
    os.environ["XESN_B200_GRAM"] = "i8"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lazy = LazyESN({"x": 2}, {"x": 1}, "periodic", 136, 0.5, 1e-4, **{k: dict(v) for k, v in KW.items()})
    lazy.build()
    lazy.train(data[:, :steps], n_spinup=4, batch_size=48)
    lp = lazy.predict(data[:, steps:], n_steps=5, n_spinup=8)
    assert np.isfinite(lp).all()
    print("lazy persistent loop: ok", flush=True)
    os.environ["XESN_B200_PREDICT"] = "steps"
    lazy2 = LazyESN({"x": 2, "time": -1}, {"x": 1, "time": 0}, "periodic", 136, 0.5, 1e-4, **{k: dict(v) for k, v in KW.items()})
    lazy2.build()
    lazy2.Wout = lazy.Wout
    lazy2._geometry(data[:, steps:])
    lp2 = lazy2.predict(data[:, steps:], n_steps=5, n_spinup=8)
    np.testing.assert_allclose(lp2, lp, rtol=0, atol=1e-9)
    print("lazy per-step fallback: ok", flush=True)
    print("SANITIZE SMOKE OK")


if __name__ == "__main__":
    main()
