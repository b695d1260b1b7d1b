#!/usr/bin/env python
"""Benchmark of the ESN hot path: reservoir unit-steps/s (train + predict).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload eager16k|eager1k|lazy|lazy2d|macro]

A "step" is one full pass of the hot path over one synthetic series:
`train` (recurrence + Gram accumulation + ridge solve) followed by `predict` (spin-up + closed
loop).  Default workload = BASELINE.json configs[1]: eager ESN, n_reservoir=16000, Lorenz-96
(40 variables), 20 000 training steps, predict with 500 spin-up + 1000 forecast steps.
unit-steps = G * N * (T_train + n_spinup_pred + n_steps_pred); `build()` and
the upload of the reservoir matrices are outside the timed region.

  value : K steps timed with CUDA events, inputs already resident in HBM
  e2e   : the same K steps through the public API (ESN.train / ESN.predict) with pinned HOST
          inputs -> H2D inside the timed region, readout + forecast copied back to the host
  N > 1 : eager ESN does not shard (all-to-all of the state every ~microsecond), so the ranks run
          independent replicas ("replicas only", weak scaling).  `--workload lazy` shards LazyESN
          groups over the ranks instead (weak scaling, halo exchange per forecast step).
  --impl reference : the reference's numpy algorithm (oracle port of xesn/esn.py) on the host cores,
          timed on a bounded sample of the same workload and extrapolated linearly in T.
  --workload : eager16k = BASELINE configs[1] (default, the configuration the metric is quoted on);
          eager1k = configs[0]; lazy = configs[2] (16 groups per GPU); lazy2d = the share of configs[3] one of
          8 GPUs holds (32 groups of N=6000, I=1296, O=1024); macro = the share of configs[4] one of 8 GPUs
          holds (8 candidates of N=4000 as one ESNEnsemble batch, 5 forecast windows, device-side NRMSE cost).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reservoir unit-steps/sec (train+predict)"
UNIT = "unit-steps/s"

ESN_KW = dict(
    input_kwargs=dict(factor=0.5, distribution="uniform", normalization="svd", random_seed=0),
    adjacency_kwargs=dict(factor=0.9, distribution="uniform", normalization="svd", is_sparse=True,
                          connectedness=5, format="csr", random_seed=1),
    bias_kwargs=dict(factor=0.5, distribution="uniform", random_seed=2),
)

WORKLOADS = {
    # BASELINE.json configs[1]
    "eager16k": dict(kind="eager", n_vars=40, n_reservoir=16000, n_train=20000, n_spinup_train=0,
                     pred_spinup=500, pred_steps=1000, leak=0.5, beta=1e-6),
    # BASELINE.json configs[0] (the reference's own CPU-sized case), handy for quick checks
    "eager1k": dict(kind="eager", n_vars=40, n_reservoir=1000, n_train=20000, n_spinup_train=0,
                    pred_spinup=500, pred_steps=1000, leak=0.5, beta=1e-6),
    # BASELINE.json configs[2] per GPU: 16 groups of 16 variables, overlap 1, N=2000 (weak scaling)
    "lazy": dict(kind="lazy", groups_per_gpu=16, chunk=16, overlap=1, n_reservoir=2000, n_train=20000,
                 n_spinup_train=0, pred_spinup=500, pred_steps=1000, leak=0.5, beta=1e-6),
    # BASELINE.json configs[3] per GPU at 8 GPUs: 2 x 16 groups of 32 x 32 cells, overlap 2 (I=1296, O=1024), N=6000;
    # the domain is (64 x world) x 512, doubly periodic, sharded by group rows (weak scaling)
    "lazy2d": dict(kind="lazy2d", groups_per_gpu=32, group_rows_per_gpu=2, group_cols=16, chunk=32, overlap=2,
                   n_reservoir=6000, n_train=4096, n_spinup_train=0, pred_spinup=64, pred_steps=256, leak=0.5, beta=1e-6),
    # BASELINE.json configs[4] per GPU: 8 of the 64 hyper-parameter candidates (N=4000, Lorenz-96 with 40 variables), each
    # trained on 20 000 steps and scored on 5 forecasts of 100 spin-up + 500 steps (docs/config.yaml:70-74)
    "macro": dict(kind="macro", groups_per_gpu=8, n_vars=40, n_reservoir=4000, n_train=20000, n_spinup_train=0,
                  n_forecasts=5, pred_spinup=100, pred_steps=500),
}


def unit_steps(w, n_groups):
    return n_groups * w["n_reservoir"] * (w["n_train"] + w.get("n_forecasts", 1) * (w["pred_spinup"] + w["pred_steps"]))


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(w, data, W, Win, b, sample_train=300, sample_pred=100, time_solve=True):
    """Oracle port (numpy restatement of xesn/esn.py `_train_1d` + predict loop) on the host cores,
    on a bounded sample, extrapolated linearly in the number of time steps.  Eager workload only."""
    from oracle import esn_oracle as oc
    N = w["n_reservoir"]
    u = data[:, :sample_train]
    t0 = time.perf_counter()
    rr, yr = oc.ridge_accumulate(u, u, 0, sample_train, W, Win, b, w["leak"])
    t_acc = time.perf_counter() - t0
    t_solve = None
    if time_solve:
        t0 = time.perf_counter()
        Wout = oc.ridge_solve(rr, yr, w["beta"])
        t_solve = time.perf_counter() - t0
    else:
        Wout = np.zeros((u.shape[0], N))
    del rr
    t0 = time.perf_counter()
    oc.closed_loop(data[:, :sample_pred + 1], sample_pred, sample_pred // 2, W, Win, b, w["leak"], Wout)
    t_pred = (time.perf_counter() - t0) / (sample_pred + sample_pred // 2)
    return {"per_train_step_s": t_acc / sample_train, "solve_s": t_solve, "per_predict_step_s": t_pred}


def cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, sample_train=400, sample_pred=150):
    """CPU baseline of the multi-reservoir workloads: ONE reservoir of the workload's shape (n_in inputs, n_out
    targets) through the oracle port on a bounded sample -- Gram accumulation, the full solve, and update + readout
    steps -- extrapolated linearly in T and multiplied by the reservoirs per GPU (the reference runs them as
    independent tasks: `_train_nd` per dask block, one ESN per candidate in `cost._cost`)."""
    from oracle import esn_oracle as oc
    rs = np.random.RandomState(0)
    N = w["n_reservoir"]
    u = rs.normal(size=(n_in, sample_train))
    y = np.ascontiguousarray(u[:n_out]) if n_out <= n_in else rs.normal(size=(n_out, sample_train))
    t0 = time.perf_counter()
    rr, yr = oc.ridge_accumulate(u, y, 0, sample_train, W, Win, b, leak)
    t_acc = (time.perf_counter() - t0) / sample_train
    t0 = time.perf_counter()
    Wout = oc.ridge_solve(rr, yr, max(beta, 1e-6))
    t_solve = time.perf_counter() - t0
    del rr
    r = np.zeros(N)
    t0 = time.perf_counter()
    for t in range(sample_pred):
        r = oc.leaky_step(r, u[:, t], W, Win, b, leak)
        v = Wout @ r
    t_pred = (time.perf_counter() - t0) / sample_pred
    n_pred = w.get("n_forecasts", 1) * (w["pred_spinup"] + w["pred_steps"])
    per_reservoir = w["n_train"] * t_acc + t_solve + n_pred * t_pred
    total = per_reservoir * w["groups_per_gpu"]
    return {"value": unit_steps(w, w["groups_per_gpu"]) / total, "total_s": total, "per_train_step_s": t_acc,
            "solve_s": t_solve, "per_predict_step_s": t_pred, "sample_train": sample_train, "sample_pred": sample_pred}


def cpu_extrapolate(w, parts, solve_s):
    total = (w["n_train"] * parts["per_train_step_s"] + solve_s
             + (w["pred_spinup"] + w["pred_steps"]) * parts["per_predict_step_s"])
    return unit_steps(w, 1) / total, total


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count()])
    except Exception:
        return os.cpu_count()


def build_eager_matrices(w):
    """Host build of (W, Win, b) with the reference's conventions (outside every timed region)."""
    from xesn_b200 import ESN
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = ESN(w["n_vars"], w["n_vars"], w["n_reservoir"], w["leak"], w["beta"], **{k: dict(v) for k, v in ESN_KW.items()})
        esn.build()
    return esn


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process that should use the box's cores."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:
        pass


def run_reference(args, w, rank):
    if rank != 0:
        return
    use_all_host_threads()
    if w["kind"] != "eager":
        run_reference_grouped(args, w)
        return
    from xesn_b200.synthetic import lorenz96
    import scipy.sparse as sp
    esn = build_eager_matrices(w)
    W = sp.csr_matrix(esn.W)
    n_t = 400
    data = lorenz96(w["n_vars"], n_t + 200)
    sample_train, sample_pred = (200, 60) if w["n_reservoir"] > 4000 else (2000, 500)
    # the N^3/3 solve does not depend on T: time it once (first warm-up step) and reuse it
    first = cpu_baseline(w, data, W, esn.Win, esn.bias_vector, sample_train, sample_pred, time_solve=True)
    solve_s = first["solve_s"]
    for _ in range(max(args.warmup - 1, 0)):
        cpu_baseline(w, data, W, esn.Win, esn.bias_vector, sample_train, sample_pred, time_solve=False)
    vals, totals = [], []
    for _ in range(args.steps):
        parts = cpu_baseline(w, data, W, esn.Win, esn.bias_vector, sample_train, sample_pred, time_solve=False)
        v, total = cpu_extrapolate(w, parts, solve_s)
        vals.append(v)
        totals.append(total)
    value = float(np.mean(vals))
    cores = blas_threads()
    sample = (f"oracle port of xesn/esn.py on {cores} host threads: {sample_train} train steps + Gram, "
              f"{sample_pred + sample_pred // 2} predict steps per bench step, ridge solve (N^3/3) timed once "
              f"({solve_s:.1f} s); extrapolated linearly to {w['n_train']} train + {w['pred_spinup'] + w['pred_steps']} predict steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(totals)) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, w, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def build_grouped_host(w):
    """Host build of the shared (W, Win, b) of a multi-reservoir workload; returns (model, n_in, n_out, leak, beta)."""
    from xesn_b200 import ESN, LazyESN
    kw = {k: dict(v) for k, v in ESN_KW.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if w["kind"] == "macro":
            m = ESN(w["n_vars"], w["n_vars"], w["n_reservoir"], 0.5, 1e-6, **kw)
            n_in = n_out = w["n_vars"]
            leak, beta = 0.5, 1e-6
        else:
            names = ("x", "y")[:2 if w["kind"] == "lazy2d" else 1]
            m = LazyESN({n: w["chunk"] for n in names}, {n: w["overlap"] for n in names}, "periodic", w["n_reservoir"],
                        w["leak"], w["beta"], **kw)
            nd = len(names)
            n_in, n_out = (w["chunk"] + 2 * w["overlap"]) ** nd, w["chunk"] ** nd
            leak, beta = w["leak"], w["beta"]
        m.build()
    return m, n_in, n_out, leak, beta


def run_reference_grouped(args, w):
    """`--impl reference` for the multi-reservoir workloads: the oracle port on the host cores, one reservoir of
    the workload's shape on a bounded sample, extrapolated (cpu_baseline_grouped)."""
    import scipy.sparse as sp
    m, n_in, n_out, leak, beta = build_grouped_host(w)
    W, Win, b = sp.csr_matrix(m.W), np.asarray(m.Win), np.asarray(m.bias_vector).reshape(-1)
    st, sp_ = (200, 60) if w["n_reservoir"] > 4000 else (400, 150)
    vals, totals = [], []
    for _ in range(max(args.warmup - 1, 0)):
        cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, st, sp_)
    res = None
    for _ in range(args.steps):
        res = cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, st, sp_)
        vals.append(res["value"])
        totals.append(res["total_s"])
    value, cores = float(np.mean(vals)), blas_threads()
    sample = (f"oracle port of xesn/esn.py on {cores} host threads, ONE reservoir of this shape (I={n_in}, O={n_out}): {st} train "
              f"steps + Gram, full solve, {sp_} update+readout steps; extrapolated linearly in T and x {w['groups_per_gpu']} "
              f"reservoirs per GPU")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(totals)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, w, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(name, w, world):
    gram_gb = w.get("groups_per_gpu", 1) * w["n_reservoir"] ** 2 * 8 / 1e9
    cfg = {"workload": name,
           "inputs": (f"larger than L2 (Gram accumulators {gram_gb:.2f} GB per GPU >> 126 MB L2)" if gram_gb > 0.5 else
                      f"working set {gram_gb * 1e3:.0f} MB fits the 126 MB L2 (latency-bound configuration, no flush)")}
    cfg.update({k: v for k, v in w.items() if k != "kind"})
    cfg["parallelism"] = "single GPU" if world == 1 else (
        f"{world} independent replicas (eager ESN does not shard)" if w["kind"] == "eager"
        else f"{w['groups_per_gpu']} independent candidates per rank, no communication" if w["kind"] == "macro"
        else f"LazyESN groups sharded over {world} ranks, halo exchange per forecast step")
    return cfg


# --------------------------------------------------------------------------- FP64 tensor peak probe
def dgemm_peak(torch, n=8192, reps=6):
    """cuBLAS DGEMM n^3 (the FP64 counterpart of MEASURED_PEAKS.json's bf16 probe): burst best-of-reps."""
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="eager16k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    import torch
    import torch.distributed as dist
    from xesn_b200 import _lib, launch_count
    from xesn_b200.synthetic import lorenz96

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ build the workload
    t_total = w["n_train"] + w.get("n_forecasts", 1) * (w["pred_spinup"] + w["pred_steps"] + 1)
    if w["kind"] == "macro":
        # candidates drawn like the reference's macro-training search space; every candidate
        # is an ESN rebuilt from the same seeds with its own factors, leak rate and Tikhonov parameter (cost.py:190-203)
        from xesn_b200 import ESN
        rs = np.random.RandomState(rank)
        esns = []
        for _ in range(w["groups_per_gpu"]):
            f_in, f_adj, f_b = rs.uniform(0.01, 2.0, size=3)
            leak, beta = rs.uniform(0.05, 1.0), 10.0 ** rs.uniform(-8, 0)
            kw = {k: dict(v) for k, v in ESN_KW.items()}
            kw["input_kwargs"]["factor"], kw["adjacency_kwargs"]["factor"], kw["bias_kwargs"]["factor"] = f_in, f_adj, f_b
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e = ESN(w["n_vars"], w["n_vars"], w["n_reservoir"], leak, beta, **kw)
                e.build()
            esns.append(e)
        from xesn_b200 import ESNEnsemble
        ens = ESNEnsemble(esns)
        ens.build()
        esn = esns[0]
        data = lorenz96(w["n_vars"], t_total)
        n_groups_total = w["groups_per_gpu"] * world
        seg = w["pred_spinup"] + w["pred_steps"] + 1
        train_host = torch.from_numpy(np.ascontiguousarray(data[:, :w["n_train"]])).pin_memory()
        tests_host = [torch.from_numpy(np.ascontiguousarray(data[:, w["n_train"] + k * seg: w["n_train"] + (k + 1) * seg])).pin_memory()
                      for k in range(w["n_forecasts"])]
        train_dev, tests_dev = train_host.to(dev), [t.to(dev) for t in tests_host]
        tests_stacked = torch.cat(tests_dev, dim=0)

        # all candidates of the rank as ONE batch (ESNEnsemble: per-group parameters on a shared sparsity pattern)
        def step_device():
            ens.train(train_dev, n_spinup=w["n_spinup_train"])
            return ens._predict_device(tests_stacked, w["pred_steps"], w["pred_spinup"], w["n_forecasts"])

        def step_e2e():
            # what cost._cost returns per candidate (cost.py:195-225): the NRMSE of the forecasts, averaged over the
            # windows and clamped -- evaluated on the device, only the costs come back
            ens.train(train_host, n_spinup=w["n_spinup_train"])
            costs, _ = ens.cost(tests_host, n_steps=w["pred_steps"], n_spinup=w["pred_spinup"])
            return costs

        h2d = (train_host.numel() + sum(t.numel() for t in tests_host)) * 8  # one upload serves every candidate
        d2h = w["groups_per_gpu"] * (1 + w["n_forecasts"]) * 8  # costs and per-window NRMSE
    elif w["kind"] == "lazy2d":
        from xesn_b200 import LazyESN
        from xesn_b200.synthetic import wave_field_2d
        nx, ny = w["group_rows_per_gpu"] * w["chunk"] * world, w["group_cols"] * w["chunk"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            esn = LazyESN({"x": w["chunk"], "y": w["chunk"]}, {"x": w["overlap"], "y": w["overlap"]}, "periodic",
                          w["n_reservoir"], w["leak"], w["beta"], **{k: dict(v) for k, v in ESN_KW.items()})
            esn.build()
        data = wave_field_2d(nx, ny, t_total)
        n_groups_total = w["groups_per_gpu"] * world
        train_host = torch.from_numpy(np.ascontiguousarray(data[..., :w["n_train"]])).pin_memory()
        test_host = torch.from_numpy(np.ascontiguousarray(data[..., w["n_train"]:])).pin_memory()
        train_dev, test_dev = train_host.to(dev), test_host.to(dev)
        esn._device_reservoir()

        def step_device():
            esn.train(train_dev, n_spinup=w["n_spinup_train"])
            return esn._predict_core(test_dev, w["pred_steps"], w["pred_spinup"])

        def step_e2e():
            esn.train(train_host, n_spinup=w["n_spinup_train"])
            return esn.predict(test_host, n_steps=w["pred_steps"], n_spinup=w["pred_spinup"])

        h2d = (train_host.numel() + test_host.numel()) * 8 // world
        d2h = nx * ny * (w["pred_steps"] + 1) * 8
    elif w["kind"] == "eager":
        esn = build_eager_matrices(w)
        data = lorenz96(w["n_vars"], t_total)
        n_groups_total = world  # replicas
        train_host = torch.from_numpy(np.ascontiguousarray(data[:, :w["n_train"]])).pin_memory()
        test_host = torch.from_numpy(np.ascontiguousarray(data[:, w["n_train"]:])).pin_memory()
        train_dev, test_dev = train_host.to(dev), test_host.to(dev)
        esn._device_reservoir()

        def step_device():
            esn._wout_dev = esn._train_device(train_dev, train_dev, 1, w["n_train"], w["n_spinup_train"], None)
            esn._wout_host = None
            return esn._predict_device(test_dev, w["pred_steps"], w["pred_spinup"])

        def step_e2e():
            esn.train(train_host, n_spinup=w["n_spinup_train"])
            wout = esn.Wout                       # D2H of the readout
            pred = esn.predict(test_host, n_steps=w["pred_steps"], n_spinup=w["pred_spinup"])  # D2H of the forecast
            return wout, pred

        h2d = train_host.numel() * 8 + test_host.numel() * 8
        d2h = w["n_vars"] * w["n_reservoir"] * 8 + w["n_vars"] * (w["pred_steps"] + 1) * 8
    else:
        from xesn_b200 import LazyESN
        n_vars = w["groups_per_gpu"] * w["chunk"] * world
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            esn = LazyESN({"x": w["chunk"]}, {"x": w["overlap"]}, "periodic", w["n_reservoir"], w["leak"], w["beta"],
                          **{k: dict(v) for k, v in ESN_KW.items()})
            esn.build()
        data = lorenz96(n_vars, t_total)
        n_groups_total = w["groups_per_gpu"] * world
        train_host = torch.from_numpy(np.ascontiguousarray(data[:, :w["n_train"]])).pin_memory()
        test_host = torch.from_numpy(np.ascontiguousarray(data[:, w["n_train"]:])).pin_memory()
        train_dev, test_dev = train_host.to(dev), test_host.to(dev)
        esn._device_reservoir()

        def step_device():
            esn.train(train_dev, n_spinup=w["n_spinup_train"])
            return esn._predict_core(test_dev, w["pred_steps"], w["pred_spinup"])

        def step_e2e():
            esn.train(train_host, n_spinup=w["n_spinup_train"])
            pred = esn.predict(test_host, n_steps=w["pred_steps"], n_spinup=w["pred_spinup"])
            return pred

        h2d = (train_host.numel() + test_host.numel()) * 8 // world
        d2h = n_vars * (w["pred_steps"] + 1) * 8

    units = unit_steps(w, n_groups_total)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    launches0 = launch_count()
    _lib.profile_enable(True)
    with ClockSampler(local_rank) as clocks:
        ms_total = timed(step_device, args.steps)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    launches = launch_count() - launches0
    ms_per_step = ms_total / args.steps
    value = units / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end timing (host buffers)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = units / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ roofline of the dominant kernel
    N = w["n_reservoir"]
    G_local = 1 if w["kind"] == "eager" else w["groups_per_gpu"]
    # The Gram SYRK runs inside the fused chunk kernel (Gram role on ~132 SMs beside the recurrence role);
    # with xesn_set_fused(0) it is a launch of its own.  Either way the dominant kernel is timed with
    # CUDA events recorded by the library on the launching stream.
    fused = prof.get("fused_chunk", (0, 0.0))[0] > 0
    n_gram, ms_gram = prof["fused_chunk"] if fused else prof["gram_syrk"]
    n_trains = args.steps * (w["groups_per_gpu"] if w["kind"] == "macro" else 1)
    n_gram_launches = max(n_gram - n_trains, 1) if fused else n_gram  # first fused launch of a train() has no Gram role
    peak_tf = dgemm_peak(torch)
    # algorithmic FP64 flops of the Gram accumulation: N(N+1) per reservoir per time step (one triangle)
    gram_flops = float(G_local) * N * (N + 1) * w["n_train"] * args.steps
    achieved = gram_flops / (ms_gram * 1e-3) / 1e12 if ms_gram > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dominant_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    i8 = fused and os.environ.get("XESN_B200_GRAM", "i8").lower() not in ("f64", "fp64", "dmma", "0")
    if i8:
        # the Gram role runs on the int8 tensor cores: 28 digit-pair MMAs replace each FP64 FMA
        int_ops = 28.0 * gram_flops            # 28 pairs x (2 ops per MAC) x N(N+1)/2 MACs per step
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
        int8_peak = 2.0 * bf16
        ach = int_ops / (ms_gram * 1e-3) / 1e12
        roofline = {
            "bound": "tensor",
            "kernel": "fused_chunk_i8_kernel (Gram role: tcgen05.mma kind::i8 on fixed-point digit planes, beside "
                      "the recurrence role)",
            "achieved": ach, "peak": int8_peak, "unit": "TOP/s", "frac": ach / int8_peak,
            "traffic": traffic,
            "peak_source": ("2 x the measured sustained bf16 tensor peak of MEASURED_PEAKS.json (int8 runs at twice "
                            "the bf16 rate on B200; no int8 entry is measured)" if peaks else
                            "2 x the fallback bf16 peak of B200_PROFILING.md (MEASURED_PEAKS.json absent)"),
            "launches": n_gram_launches, "avg_launch_ms": ms_gram / n_gram_launches,
            "ops_per_launch": int_ops / n_gram_launches,
            "fp64_equivalent_tflops": achieved, "fp64_dgemm_peak_tflops": peak_tf,
            "fp64_equivalent_frac_of_dgemm_peak": achieved / peak_tf if peak_tf else None,
            "note": "achieved = 28 x N(N+1) x T int8 ops (the digit-pair MMAs of one triangle) / summed event time of "
                    "the fused launches; the same kernel hosts the recurrence on its other SMs and is currently "
                    "bounded by whichever role finishes last",
        }
    else:
        roofline = {
            "bound": "tensor",
            "kernel": ("fused_chunk_kernel (Gram SYRK role, DMMA.8x8x4, beside the recurrence role)" if fused
                       else "gemm_f64_kernel<MN,MN> lower (Gram SYRK, DMMA.8x8x4)"),
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf if peak_tf else None,
            "traffic": traffic,
            "peak_source": "cuBLAS DGEMM 8192^3 burst measured in this run (MEASURED_PEAKS.json has no FP64 entry); "
                           "hardware DMMA peak 128 flop/clk/SM = 37.2 TFLOP/s at 1965 MHz",
            "launches": n_gram_launches, "avg_launch_ms": ms_gram / n_gram_launches,
            "flops_per_launch": gram_flops / n_gram_launches,
            "note": "achieved = algorithmic Gram flops N(N+1) per step / summed event time of the kernel launches "
                    "that carry them (the fused kernel also hosts the recurrence on its other SMs)",
        }
    phases = {k: {"regions": v[0], "ms_per_step": v[1] / args.steps} for k, v in prof.items() if v[0]}

    cpu = None
    # the CPU baseline is a single-process measurement: rank 0 at N=1 only (torchrun also pins OMP_NUM_THREADS to 1)
    if not args.no_cpu_baseline and world == 1 and w["kind"] == "eager":
        import scipy.sparse as sp
        big = N > 4000
        st, sp_ = (200, 60) if big else (2000, 500)
        parts = cpu_baseline(w, data, sp.csr_matrix(esn.W), esn.Win, esn.bias_vector, st, sp_, time_solve=True)
        v, total = cpu_extrapolate(w, parts, parts["solve_s"])
        cores = blas_threads()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle port of xesn/esn.py:_train_1d + predict loop, numpy/scipy on {cores} threads: "
                          f"{st} train steps + Gram ({parts['per_train_step_s'] * 1e3:.2f} ms/step), full N^3/3 solve "
                          f"({parts['solve_s']:.1f} s), {sp_ + sp_ // 2} predict steps "
                          f"({parts['per_predict_step_s'] * 1e3:.2f} ms/step); extrapolated linearly to the full "
                          f"workload ({total:.1f} s)")}

    if not args.no_cpu_baseline and world == 1 and w["kind"] != "eager":
        import scipy.sparse as sp
        if w["kind"] == "macro":
            n_in = n_out = w["n_vars"]
            leak_c, beta_c = float(esn.leak_rate), float(esn.tikhonov_parameter)
        else:
            nd = 2 if w["kind"] == "lazy2d" else 1
            n_in, n_out = (w["chunk"] + 2 * w["overlap"]) ** nd, w["chunk"] ** nd
            leak_c, beta_c = w["leak"], w["beta"]
        big = N > 4000
        st, sp_ = (200, 60) if big else (400, 150)
        res_c = cpu_baseline_grouped(w, sp.csr_matrix(esn.W), np.asarray(esn.Win), np.asarray(esn.bias_vector).reshape(-1),
                                     n_in, n_out, leak_c, beta_c, st, sp_)
        cores = blas_threads()
        cpu = {"value": res_c["value"], "unit": UNIT, "cores": cores, "kind": "port",
               "sample": (f"oracle port of xesn/esn.py on {cores} threads, ONE reservoir of this shape (I={n_in}, O={n_out}): "
                          f"{st} train steps + Gram ({res_c['per_train_step_s'] * 1e3:.2f} ms/step), full solve "
                          f"({res_c['solve_s']:.2f} s), {sp_} update+readout steps ({res_c['per_predict_step_s'] * 1e3:.2f} "
                          f"ms/step); extrapolated linearly in T and x {w['groups_per_gpu']} reservoirs per GPU "
                          f"({res_c['total_s']:.1f} s)")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "gram_engine": ("int8 fixed-point digit planes on tcgen05 (error-free), FP64 everywhere else" if i8
                        else "FP64 DMMA"),
        "config": workload_config(args.workload, w, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "phases_ms_per_step": phases,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
