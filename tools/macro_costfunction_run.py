#!/usr/bin/env python
"""BASELINE configs[4] through the reference-facing entry point: 64 hyper-parameter candidates (N=4000, Lorenz-96 with 40
variables) handed to `xesn_b200.CostFunction.__call__` exactly like smt's EGO hands them to xesn.CostFunction
(xesn/cost.py:94-122); under torchrun the candidates are dealt round-robin to the ranks (8 per GPU at 8 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        tools/macro_costfunction_run.py
Prints one JSON line (rank 0): wall time of the first call (host build of the random matrices included, as in the
reference) and of a second call (matrices cached), unit-steps/s of the second call, the costs."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import ESN_KW  # noqa: E402
from xesn_b200 import ESN, CostFunction  # noqa: E402
from xesn_b200.synthetic import lorenz96  # noqa: E402


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n_cand = int(os.environ.get("MACRO_CANDIDATES", 64))
    N = int(os.environ.get("MACRO_N", 4000))
    n_train, n_samples, n_spinup, n_steps = 20000, 5, 100, 500
    data = lorenz96(40, n_train + n_samples * (n_spinup + n_steps + 1))
    seg = n_spinup + n_steps + 1
    train = data[:, :n_train]
    macro = [data[:, n_train + k * seg: n_train + (k + 1) * seg] for k in range(n_samples)]
    config = {
        "esn": dict(n_input=40, n_output=40, n_reservoir=N, leak_rate=0.5, tikhonov_parameter=1e-6,
                    **{k: dict(v) for k, v in ESN_KW.items()}),
        "training": {"n_spinup": 0},
        "macro_training": {
            "parameters": {"input_factor": None, "adjacency_factor": None, "bias_factor": None, "leak_rate": None,
                           "tikhonov_parameter": None},
            "transformations": {"tikhonov_parameter": "log10"},
            "forecast": {"n_samples": n_samples, "n_spinup": n_spinup, "n_steps": n_steps},
            "cost_terms": {"nrmse": 1.0}, "cost_upper_bound": 1e9},
    }
    rs = np.random.RandomState(0)   # the search space of SURVEY.md section 8(d)
    X = np.column_stack([rs.uniform(0.01, 2.0, n_cand), rs.uniform(0.01, 2.0, n_cand), rs.uniform(0.01, 2.0, n_cand),
                         rs.uniform(0.05, 1.0, n_cand), rs.uniform(-8, 0, n_cand)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cf = CostFunction(ESN, train, macro, config)
        times = []
        for _ in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            costs = cf(X, is_transformed=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            times.append(time.perf_counter() - t0)
    if rank == 0:
        units = n_cand * N * (n_train + n_samples * (n_spinup + n_steps))
        print(json.dumps({"what": "CostFunction.__call__ on BASELINE configs[4]", "n_gpus": world, "candidates": n_cand,
                          "n_reservoir": N, "first_call_s": times[0], "second_call_s": times[1],
                          "unit_steps_per_s_second_call": units / times[1], "costs_shape": list(costs.shape),
                          "n_clamped": int((costs >= 1e9).sum()), "cost_min": float(costs.min()),
                          "cost_median": float(np.median(costs))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
