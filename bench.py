#!/usr/bin/env python
"""Benchmark of the ESN hot path: reservoir unit-steps/s (train + predict).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload eager16k|eager1k|lazy|lazy2d|macro]

A "step" is one full pass of the hot path over one synthetic series:
`train` (recurrence + Gram accumulation + ridge solve) followed by `predict` (spin-up + closed
loop).  Default workload = `lazy`, BASELINE.json configs[2] per GPU (LazyESN on Lorenz-96, 16 groups of
16 variables with overlap 1, n_reservoir=2000, 20 000 training steps, predict with 500 spin-up + 1000
forecast steps), weak-scaled: the ONLY configuration the metric's "1/2/4/8 B200" spans -- the groups
are sharded over the ranks and every forecast step exchanges the halo cells between neighbouring ranks.
At N=1 the line also carries the eager N=16000 configuration (configs[1], the tensor-bound one) as the
nested object `eager16k` (value, e2e, roofline, phases); at N>1 it carries `multi_gpu_parity`: a short
sharded forecast through BOTH exchange engines (in-kernel P2P stores, NCCL) compared with the same
forecast run single-rank on rank 0 -- the run fails if they differ.
unit-steps = G * N * (T_train + n_spinup_pred + n_steps_pred); `build()` and
the upload of the reservoir matrices are outside the timed region.

  value : K steps timed with CUDA events, inputs already resident in HBM
  e2e   : the same K steps through the public API (ESN.train / ESN.predict) with pinned HOST
          inputs -> H2D inside the timed region, readout + forecast copied back to the host
  N > 1 : eager ESN does not shard (all-to-all of the state every ~microsecond), so the ranks run
          independent replicas ("replicas only", weak scaling).  `--workload lazy` shards LazyESN
          groups over the ranks instead (weak scaling, halo exchange per forecast step).
  --impl reference : the reference's numpy algorithm (oracle port of xesn/esn.py) on the host cores,
          timed on a bounded sample of the same workload (>= 2000 training steps per sample, so the N^2
          Gram update is amortised as in the real run) and extrapolated linearly in T.
  --workload : lazy = configs[2] (default, 16 groups per GPU); eager16k = BASELINE configs[1];
          eager1k = configs[0]; lazy2d = the share of configs[3] one of
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
SAMPLE_TRAIN = 2000   # training steps per CPU sample: SURVEY.md section 8(d); amortises the N^2 Gram update


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count()])
    except Exception:
        return os.cpu_count()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process that should use the box's cores."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=host_cores())
    except Exception:
        pass


def cpu_eager_parts(w, data, W, Win, b, sample_train, sample_pred, time_solve=True):
    """Oracle port (numpy restatement of xesn/esn.py `_train_1d` + predict loop) of ONE eager reservoir on all host
    threads: `sample_train` training steps as one batch (update loop, then ybar/rbar GEMMs exactly as esn.py:502-513),
    the full N^3/3 symmetric solve, and a short closed loop.  Arrays are pre-touched by an untimed 8-step pass."""
    from oracle import esn_oracle as oc
    N = w["n_reservoir"]
    oc.ridge_accumulate(data[:, :8], data[:, :8], 0, 8, W, Win, b, w["leak"])          # warm BLAS / page in
    u = np.ascontiguousarray(data[:, :sample_train])
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


def cpu_extrapolate(w, parts, solve_s):
    total = (w["n_train"] * parts["per_train_step_s"] + solve_s
             + (w["pred_spinup"] + w["pred_steps"]) * parts["per_predict_step_s"])
    return unit_steps(w, 1) / total, total


def _grouped_worker(job):
    """One reservoir of a multi-reservoir workload through the oracle port, single-threaded (one worker process per
    reservoir, the paper's best CPU mode for LazyESN: `scaling/lazy-scaling.pdf`, one dask worker per group)."""
    (W, Win, b, n_in, n_out, leak, beta, sample_train, sample_pred, seed) = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import esn_oracle as oc
    rs = np.random.RandomState(seed)
    N = Win.shape[0]
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
    return t_acc, t_solve, t_pred


def cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, n_groups, sample_train, sample_pred):
    """CPU baseline of the multi-reservoir workloads.  The reference runs the reservoirs as independent tasks
    (`_train_nd` per dask block, one ESN per candidate in `cost._cost`); its best published CPU mode is one worker
    process per group.  So: min(groups, host cores) worker processes side by side, each running ONE reservoir of the
    workload's shape through the oracle port on a bounded sample (Gram accumulation of `sample_train` steps as one
    batch, the full solve, `sample_pred` update+readout steps); per-reservoir time extrapolated linearly in T, the
    job = ceil(groups / workers) waves of that."""
    import multiprocessing as mp
    workers = max(1, min(n_groups, host_cores()))
    jobs = [(W, Win, b, n_in, n_out, leak, beta, sample_train, sample_pred, k) for k in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        res = [_grouped_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.map(_grouped_worker, jobs)
    wall = time.perf_counter() - t0
    t_acc, t_solve, t_pred = (float(np.max([r[k] for r in res])) for k in range(3))   # the slowest worker bounds a wave
    n_pred = w.get("n_forecasts", 1) * (w["pred_spinup"] + w["pred_steps"])
    per_reservoir = w["n_train"] * t_acc + t_solve + n_pred * t_pred
    waves = (n_groups + workers - 1) // workers
    total = per_reservoir * waves
    return {"value": unit_steps(w, n_groups) / total, "total_s": total, "per_train_step_s": t_acc, "solve_s": t_solve,
            "per_predict_step_s": t_pred, "sample_train": sample_train, "sample_pred": sample_pred, "workers": workers,
            "waves": waves, "sample_wall_s": wall}


def grouped_sample_text(res, n_in, n_out, n_groups):
    return (f"oracle port of xesn/esn.py, {res['workers']} single-threaded worker processes side by side (one reservoir of this "
            f"shape each, I={n_in}, O={n_out}): {res['sample_train']} train steps + Gram as one batch "
            f"({res['per_train_step_s'] * 1e3:.3f} ms/step), full solve ({res['solve_s']:.2f} s), {res['sample_pred']} "
            f"update+readout steps ({res['per_predict_step_s'] * 1e3:.3f} ms/step); sample wall {res['sample_wall_s']:.1f} s; "
            f"extrapolated linearly in T, {res['waves']} wave(s) for {n_groups} reservoirs ({res['total_s']:.1f} s per bench step)")


def build_eager_matrices(w):
    """Host build of (W, Win, b) with the reference's conventions (outside every timed region)."""
    from xesn_b200 import ESN
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = ESN(w["n_vars"], w["n_vars"], w["n_reservoir"], w["leak"], w["beta"], **{k: dict(v) for k, v in ESN_KW.items()})
        esn.build()
    return esn


def eager_cpu_sample(w, data, esn, sample_pred, solve_s=None):
    """One CPU sample of the eager workload -> (value, total_s, parts); the solve is timed unless `solve_s` is given."""
    import scipy.sparse as sp
    st = min(SAMPLE_TRAIN, w["n_train"])
    parts = cpu_eager_parts(w, data, sp.csr_matrix(esn.W), esn.Win, esn.bias_vector, st, sample_pred,
                            time_solve=solve_s is None)
    if solve_s is not None:
        parts["solve_s"] = solve_s
    v, total = cpu_extrapolate(w, parts, parts["solve_s"])
    parts["sample_train"], parts["sample_pred"] = st, sample_pred + sample_pred // 2
    return v, total, parts


def eager_sample_text(parts, total, cores):
    return (f"oracle port of xesn/esn.py:_train_1d + predict loop, numpy/scipy on {cores} threads: {parts['sample_train']} "
            f"train steps as one batch incl. the ybar/rbar GEMMs ({parts['per_train_step_s'] * 1e3:.2f} ms/step), full N^3/3 "
            f"symmetric solve ({parts['solve_s']:.1f} s, timed once), {parts['sample_pred']} predict steps "
            f"({parts['per_predict_step_s'] * 1e3:.2f} ms/step); extrapolated linearly to the full workload "
            f"({total:.1f} s per bench step)")


def reference_line(args, name, w, value, total_s, cpu, n_gpus):
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(name, w, n_gpus, host=True),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args, w, rank):
    """`--impl reference`: the reference's CPU algorithm (oracle port) on the host cores, same config/metric as our arm.
    Under torchrun only rank 0 works; the other ranks exit."""
    if rank != 0:
        return
    use_all_host_threads()
    if w["kind"] != "eager":
        run_reference_grouped(args, w)
        return
    from xesn_b200.synthetic import lorenz96
    esn = build_eager_matrices(w)
    data = lorenz96(w["n_vars"], SAMPLE_TRAIN + 200)
    sample_pred = 60 if w["n_reservoir"] > 4000 else 500
    # every bench step is one full sample (>= 2000 training steps); the N^3/3 solve does not depend on T, so it is
    # timed once, in the first pass, and reused -- the arm finishes in minutes at N = 16000
    v, total, parts = eager_cpu_sample(w, data, esn, sample_pred)
    solve_s = parts["solve_s"]
    vals, totals = ([v], [total]) if args.warmup == 0 else ([], [])
    while len(vals) < args.steps:
        v, total, parts = eager_cpu_sample(w, data, esn, sample_pred, solve_s=solve_s)
        vals.append(v)
        totals.append(total)
    value, total = float(np.mean(vals)), float(np.mean(totals))
    cores = blas_threads()
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": eager_sample_text(parts, total, cores)}
    # eager ESN does not shard: at --gpus N our arm runs N replicas, the host arm still has one set of cores
    print(json.dumps(reference_line(args, args.workload, w, value, total, cpu, 1)))


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


def grouped_samples(w):
    big = w["n_reservoir"] > 4000
    return (min(SAMPLE_TRAIN, w["n_train"]), 60 if big else 150)


def run_reference_grouped(args, w):
    """`--impl reference` for the multi-reservoir workloads (cpu_baseline_grouped).  The workload at --gpus N is
    N x groups_per_gpu reservoirs (weak scaling), all of them on the one set of host cores."""
    import scipy.sparse as sp
    m, n_in, n_out, leak, beta = build_grouped_host(w)
    W, Win, b = sp.csr_matrix(m.W), np.asarray(m.Win), np.asarray(m.bias_vector).reshape(-1)
    st, sp_ = grouped_samples(w)
    n_groups = w["groups_per_gpu"] * max(args.gpus, 1)
    vals, totals, res = [], [], None
    for _ in range(max(args.warmup - 1, 0)):
        cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, n_groups, st, sp_)
    for _ in range(args.steps):
        res = cpu_baseline_grouped(w, W, Win, b, n_in, n_out, leak, beta, n_groups, st, sp_)
        vals.append(res["value"])
        totals.append(res["total_s"])
    value, total = float(np.mean(vals)), float(np.mean(totals))
    cpu = {"value": value, "unit": UNIT, "cores": res["workers"], "kind": "port",
           "sample": grouped_sample_text(res, n_in, n_out, n_groups)}
    print(json.dumps(reference_line(args, args.workload, w, value, total, cpu, max(args.gpus, 1))))


def workload_config(name, w, world, host=False):
    gram_gb = w.get("groups_per_gpu", 1) * w["n_reservoir"] ** 2 * 8 / 1e9
    cfg = {"workload": name,
           "inputs": (f"larger than L2 (Gram accumulators {gram_gb:.2f} GB per GPU >> 126 MB L2)" if gram_gb > 0.5 else
                      f"working set {gram_gb * 1e3:.0f} MB fits the 126 MB L2 (latency-bound configuration, no flush)")}
    cfg.update({k: v for k, v in w.items() if k != "kind"})
    if host:
        cfg["parallelism"] = ("host cores (one process, BLAS threads)" if w["kind"] == "eager" else
                              f"{w['groups_per_gpu'] * world} reservoirs on the host cores, one worker process per reservoir")
        return cfg
    cfg["parallelism"] = "single GPU" if world == 1 else (
        f"{world} independent replicas (eager ESN does not shard)" if w["kind"] == "eager"
        else f"{w['groups_per_gpu']} independent candidates per rank, no communication" if w["kind"] == "macro"
        else f"LazyESN groups sharded over {world} ranks, halo exchange per forecast step")
    return cfg


# --------------------------------------------------------------------------- tensor peak probes
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


def int8_peak(torch, n=8192, reps=10, sustain_s=1.0):
    """int8 x int8 -> int32 tensor-core GEMM n^3 through cuBLASLt (torch._int_mm), the int8 counterpart of the bf16
    probe of MEASURED_PEAKS.json: (burst best-of-reps, sustained back-to-back for ~1 s) in TOP/s, or None."""
    try:
        a = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
        torch._int_mm(a, b)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        n_rep = max(4, int(sustain_s * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_rep):
            torch._int_mm(a, b)
        e1.record()
        e1.synchronize()
        sustained = 2.0 * n ** 3 * n_rep / (e0.elapsed_time(e1) * 1e-3) / 1e12
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12, sustained
    except Exception as exc:  # pragma: no cover - probe only
        sys.stderr.write(f"int8 peak probe failed: {exc}\n")
        return None


# --------------------------------------------------------------------------- our arm
class Ctx:
    """Process-wide handles of one bench run (torch, the distributed job, the device of this rank)."""

    def __init__(self, torch, dist, world, rank, local_rank):
        self.torch, self.dist, self.world, self.rank, self.local_rank = torch, dist, world, rank, local_rank
        self.dev = torch.device(f"cuda:{local_rank}")

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, k):
        """K calls bracketed by barrier + synchronize on both sides, CUDA events, MAX over the ranks (ms)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def make_workload(ctx, name, w, world):
    """Builds the model and the synthetic series of workload `name`; returns a dict with step_device / step_e2e
    callables, the byte counts of the end-to-end copies and what the roofline / CPU legs need."""
    torch, dev, rank = ctx.torch, ctx.dev, ctx.rank
    from xesn_b200.synthetic import lorenz96
    t_total = w["n_train"] + w.get("n_forecasts", 1) * (w["pred_spinup"] + w["pred_steps"] + 1)
    out = {"extra": {}}
    if w["kind"] == "macro":
        # candidates drawn like the reference's macro-training search space; every candidate
        # is an ESN rebuilt from the same seeds with its own factors, leak rate and Tikhonov parameter (cost.py:190-203)
        from xesn_b200 import ESN, ESNEnsemble
        rs = np.random.RandomState(rank)
        esns = []
        for _ in range(w["groups_per_gpu"]):
            f_in, f_adj, f_b = rs.uniform(0.01, 2.0, size=3)
            leak, beta = rs.uniform(0.05, 1.0), 10.0 ** rs.uniform(-8, 0)
            kw = {k: dict(v) for k, v in ESN_KW.items()}
            kw["input_kwargs"]["factor"], kw["adjacency_kwargs"]["factor"], kw["bias_kwargs"]["factor"] = f_in, f_adj, f_b
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                esns.append(ESN(w["n_vars"], w["n_vars"], w["n_reservoir"], leak, beta, **kw))
        ens = ESNEnsemble(esns)
        ens.build()
        esn = esns[0]
        data = lorenz96(w["n_vars"], t_total)
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

        out.update(n_groups_total=w["groups_per_gpu"] * world, g_local=w["groups_per_gpu"],
                   h2d=(train_host.numel() + sum(t.numel() for t in tests_host)) * 8,  # one upload serves every candidate
                   d2h=w["groups_per_gpu"] * (1 + w["n_forecasts"]) * 8)                # costs and per-window NRMSE
        out["cpu_shape"] = (w["n_vars"], w["n_vars"], float(esn.leak_rate), float(esn.tikhonov_parameter))
    elif w["kind"] in ("lazy", "lazy2d"):
        from xesn_b200 import LazyESN
        two_d = w["kind"] == "lazy2d"
        names = ("x", "y") if two_d else ("x",)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            esn = LazyESN({n: w["chunk"] for n in names}, {n: w["overlap"] for n in names}, "periodic",
                          w["n_reservoir"], w["leak"], w["beta"], **{k: dict(v) for k, v in ESN_KW.items()})
            esn.build()
        if two_d:
            from xesn_b200.synthetic import wave_field_2d
            nx, ny = w["group_rows_per_gpu"] * w["chunk"] * world, w["group_cols"] * w["chunk"]
            data = wave_field_2d(nx, ny, t_total)
            n_cells = nx * ny
        else:
            n_cells = w["groups_per_gpu"] * w["chunk"] * world
            data = lorenz96(n_cells, t_total)
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

        nd = len(names)
        out.update(n_groups_total=w["groups_per_gpu"] * world, g_local=w["groups_per_gpu"],
                   h2d=(train_host.numel() + test_host.numel()) * 8 // world, d2h=n_cells * (w["pred_steps"] + 1) * 8)
        out["cpu_shape"] = ((w["chunk"] + 2 * w["overlap"]) ** nd, w["chunk"] ** nd, w["leak"], w["beta"])
        out["lazy"] = (esn, names, test_dev)
    else:
        esn = build_eager_matrices(w)
        data = lorenz96(w["n_vars"], t_total)
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

        out.update(n_groups_total=world, g_local=1,  # replicas
                   h2d=train_host.numel() * 8 + test_host.numel() * 8,
                   d2h=w["n_vars"] * w["n_reservoir"] * 8 + w["n_vars"] * (w["pred_steps"] + 1) * 8)
    out.update(esn=esn, data=data, step_device=step_device, step_e2e=step_e2e)
    return out


def multi_gpu_parity(ctx, w, esn, names, test_dev, n_steps=48, n_spinup=32):
    """Forecast `n_steps` steps of the trained sharded model through EVERY exchange engine -- the dataflow kernel with
    neighbour-only P2P stores of the halo cells (xesn_predict_flow, the default), the persistent kernel that stores
    every cell to every rank (xesn_predict_p2p) and the per-step NCCL exchange -- and compare each with the same
    forecast run on rank 0 alone (all groups on one GPU through xesn_predict, no exchange).  Collective."""
    import xesn_b200.lazyesn as lz
    from xesn_b200 import LazyESN
    torch, dist = ctx.torch, ctx.dist
    y = test_dev[..., :n_spinup + n_steps + 1].contiguous()
    preds = {}
    prev = os.environ.get("XESN_B200_EXCHANGE")
    for engine in ("flow", "p2p", "nccl"):
        os.environ["XESN_B200_EXCHANGE"] = engine
        preds[engine] = esn._predict_core(y, n_steps, n_spinup).clone()
    if prev is None:
        os.environ.pop("XESN_B200_EXCHANGE", None)
    else:
        os.environ["XESN_B200_EXCHANGE"] = prev
    blocks = esn.Wout            # collective: all-gather of the per-group readouts, reference block layout
    ctx.barrier()
    result = None
    if ctx.rank == 0:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            solo = LazyESN({n: w["chunk"] for n in names}, {n: w["overlap"] for n in names}, "periodic", w["n_reservoir"],
                           w["leak"], w["beta"], **{k: dict(v) for k, v in ESN_KW.items()})
        solo.W, solo.Win, solo.bias_vector = esn.W, esn.Win, esn.bias_vector      # same matrices, no second host build
        lz._FORCE_SINGLE_RANK = True
        try:
            solo._geometry(y)
            solo.Wout = blocks
            ref = solo._predict_core(y, n_steps, n_spinup)
            torch.cuda.synchronize()
        finally:
            lz._FORCE_SINGLE_RANK = False
        scale = float(ref.abs().max())
        errs = {k: float((v - ref).abs().max()) for k, v in preds.items()}
        tol = 1e-6 * max(scale, 1.0)
        result = {"what": (f"{n_steps}-step forecast after {n_spinup} spin-up steps of the trained workload model, sharded over "
                           f"{ctx.world} ranks, against the same forecast on rank 0 alone (xesn_predict, all "
                           f"{w['groups_per_gpu'] * ctx.world} groups on one GPU)"),
                  "max_abs_err": max(errs.values()), "max_abs_err_flow": errs["flow"], "max_abs_err_p2p": errs["p2p"],
                  "max_abs_err_nccl": errs["nccl"], "engine_timed": os.environ.get("XESN_B200_EXCHANGE", "flow"),
                  "field_scale": scale, "tolerance": tol, "ok": bool(max(errs.values()) <= tol)}
        solo._drop_device_state()
    ctx.barrier()
    return result


def roofline_of(ctx, args, name, w, prof, g_local, peaks_int8):
    """Roofline object of the dominant kernel (the fused chunk kernel: recurrence role beside the Gram role)."""
    torch = ctx.torch
    N = w["n_reservoir"]
    fused = prof.get("fused_chunk", (0, 0.0))[0] > 0
    n_gram, ms_gram = prof["fused_chunk"] if fused else prof["gram_syrk"]
    n_trains = args.steps
    n_gram_launches = max(n_gram - n_trains, 1) if fused else max(n_gram, 1)  # first fused launch of a train() has no Gram role
    peak_tf = dgemm_peak(torch)
    # algorithmic FP64 flops of the Gram accumulation: N(N+1) per reservoir per time step (one triangle)
    gram_flops = float(g_local) * N * (N + 1) * w["n_train"] * args.steps
    achieved = gram_flops / (ms_gram * 1e-3) / 1e12 if ms_gram > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(name, {}).get("dominant_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    i8 = fused and os.environ.get("XESN_B200_GRAM", "i8").lower() not in ("f64", "fp64", "dmma", "0")
    if not i8:
        return {
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
                    "that carry them (the fused kernel also hosts the recurrence on its other SMs)"}, False
    # the Gram role runs on the int8 tensor cores: 28 digit-pair MMAs replace each FP64 FMA
    int_ops = 28.0 * gram_flops            # 28 pairs x (2 ops per MAC) x N(N+1)/2 MACs per step
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    derived = 2.0 * bf16
    if peaks_int8 is not None:
        burst, sustained = peaks_int8
        # the fused kernel runs inside a long step: compare with the sustained figure; never with less than the
        # bf16-derived one (a poorly tuned library IGEMM must not flatter the fraction)
        peak, src = max(sustained, derived), (
            f"measured in this run: torch._int_mm (cuBLASLt int8 GEMM) 8192^3, sustained {sustained:.0f} TOP/s (burst "
            f"{burst:.0f}); floor = 2 x the sustained bf16 peak of MEASURED_PEAKS.json ({derived:.0f}); the larger is used")
    else:
        peak, src = derived, ("2 x the measured sustained bf16 tensor peak of MEASURED_PEAKS.json (the int8 probe "
                              "torch._int_mm is unavailable here)")
    ach = int_ops / (ms_gram * 1e-3) / 1e12
    return {
        "bound": "tensor",
        "kernel": "fused_chunk_i8_kernel (Gram role: tcgen05.mma kind::i8 on fixed-point digit planes) beside the recurrence "
                  "role -- the same kernel's other CTAs for large reservoirs, the cluster kernel fused_cluster_i8_kernel launched "
                  "next to it for small ones (DSMEM state exchange)",
        "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak, "traffic": traffic, "peak_source": src,
        "int8_peak_measured": None if peaks_int8 is None else {"burst": peaks_int8[0], "sustained": peaks_int8[1]},
        "int8_peak_derived_2x_bf16": derived,
        "launches": n_gram_launches, "avg_launch_ms": ms_gram / n_gram_launches, "ops_per_launch": int_ops / n_gram_launches,
        "fp64_equivalent_tflops": achieved, "fp64_dgemm_peak_tflops": peak_tf,
        "fp64_equivalent_frac_of_dgemm_peak": achieved / peak_tf if peak_tf else None,
        "note": "achieved = 28 x N(N+1) x T x G int8 ops (the digit-pair MMAs of one triangle) / summed event time of the "
                "chunk launches (event pair around both roles); a launch is bounded by whichever role finishes last (small "
                "reservoirs: the serial recurrence chain, so the fraction is informational there)"}, True


def run_ours(ctx, args, name, steps, warmup, with_cpu=True, peaks_int8=None):
    """One workload through our arm; returns the JSON line (rank 0) or None (other ranks)."""
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    from xesn_b200 import _lib, launch_count
    w = dict(WORKLOADS[name])
    wl = make_workload(ctx, name, w, world)
    step_device, step_e2e = wl["step_device"], wl["step_e2e"]
    units = unit_steps(w, wl["n_groups_total"])

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    parity = None
    if world > 1 and "lazy" in wl:
        parity = multi_gpu_parity(ctx, w, *wl["lazy"])
        if rank == 0 and not parity["ok"]:
            print(json.dumps({"error": "multi-GPU parity failed", "multi_gpu_parity": parity}))
            sys.stdout.flush()
        ok = torch.tensor([1 if (parity is None or parity["ok"]) else 0], device=ctx.dev)
        dist.broadcast(ok, 0)
        if int(ok.item()) == 0:
            raise SystemExit(3)
        step_device()  # re-warm after the check
    launches0 = launch_count()
    _lib.profile_enable(True)
    with ClockSampler(ctx.local_rank) as clocks:
        ms_total = ctx.timed(step_device, steps)
    prof = _lib.profile_read()
    fused_ms = _lib.profile_regions("fused_chunk")
    _lib.profile_enable(False)
    launches = launch_count() - launches0
    ms_per_step = ms_total / steps
    value = units / (ms_per_step * 1e-3)

    # ------------------------------------------------------------------ end-to-end timing (host buffers)
    step_e2e()
    ms_e2e = ctx.timed(step_e2e, steps) / steps
    e2e_value = units / (ms_e2e * 1e-3)
    if rank != 0:
        return None

    sub = argparse.Namespace(steps=steps)
    roofline, i8 = roofline_of(ctx, sub, name, w, prof, wl["g_local"], peaks_int8)
    phases = {k: {"regions": v[0], "ms_per_step": v[1] / steps} for k, v in prof.items() if v[0]}
    if len(fused_ms) >= 3 * steps and len(fused_ms) % steps == 0:
        # launches of one training pass: the first runs the recurrence role alone (nothing to accumulate yet), the last
        # the Gram role alone (on every SM), the others both side by side
        per = len(fused_ms) // steps
        one = fused_ms[-per:]
        phases["fused_chunk"]["launches_ms"] = {"recurrence_only_first": round(one[0], 3),
                                                "both_median": round(float(np.median(one[1:-1])), 3) if per > 2 else None,
                                                "gram_only_last": round(one[-1], 3), "per_pass": per}

    cpu = None
    # the CPU baseline is a single-process measurement: rank 0 at N=1 only (torchrun also pins OMP_NUM_THREADS to 1)
    if with_cpu and world == 1:
        esn = wl["esn"]
        if w["kind"] == "eager":
            v, total, parts = eager_cpu_sample(w, wl["data"], esn, 60 if w["n_reservoir"] > 4000 else 500)
            cores = blas_threads()
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": eager_sample_text(parts, total, cores)}
        else:
            import scipy.sparse as sp
            n_in, n_out, leak_c, beta_c = wl["cpu_shape"]
            st, sp_ = grouped_samples(w)
            res_c = cpu_baseline_grouped(w, sp.csr_matrix(esn.W), np.asarray(esn.Win), np.asarray(esn.bias_vector).reshape(-1),
                                         n_in, n_out, leak_c, beta_c, w["groups_per_gpu"], st, sp_)
            cpu = {"value": res_c["value"], "unit": UNIT, "cores": res_c["workers"], "kind": "port",
                   "sample": grouped_sample_text(res_c, n_in, n_out, w["groups_per_gpu"])}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "gram_engine": ("int8 fixed-point digit planes on tcgen05 (absolute error <= K 2^-50 per entry, i.e. FP64-DGEMM "
                        "level for |r| <= 1), FP64 everywhere else" if i8 else "FP64 DMMA"),
        "config": workload_config(name, w, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(wl["h2d"]),
                "d2h_bytes_per_step": int(wl["d2h"])},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "phases_ms_per_step": phases,
    }
    if parity is not None:
        line["multi_gpu_parity"] = parity
    return line


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)   # ~0.4 s timed at the default workload: several clock samples
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lazy", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nested", action="store_true", help="skip the nested eager16k measurement of the default run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    import torch
    import torch.distributed as dist

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = Ctx(torch, dist, world, rank, local_rank)
    peaks_int8 = int8_peak(torch) if rank == 0 else None

    line = run_ours(ctx, args, args.workload, args.steps, args.warmup, with_cpu=not args.no_cpu_baseline,
                    peaks_int8=peaks_int8)
    # the tensor-bound configuration (BASELINE configs[1]) rides along on the default single-GPU run
    if args.workload == "lazy" and world == 1 and not args.no_nested:
        torch.cuda.empty_cache()
        sub = run_ours(ctx, args, "eager16k", max(2, min(args.steps, 3)), 3, with_cpu=not args.no_cpu_baseline,
                       peaks_int8=peaks_int8)
        line["eager16k"] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e",
                                                 "gpu_launches", "roofline", "cpu_baseline", "phases_ms_per_step",
                                                 "gram_engine")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
