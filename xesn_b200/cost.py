"""CostFunction -- the macro-training cost of the reference, evaluated for whole batches of candidates.

Drop-in for `xesn.CostFunction` on the eager `ESN` (reference xesn/cost.py:23-172: constructor, `__call__`;
`_cost` :175-225).  The reference evaluates the parameter sets one after the other: build an `ESN(**params)`
from the same seeds, train it, forecast `n_samples` windows, NRMSE (+ PSD-NRMSE) of each, average, clamp.  Here
the sets of a call are one or more `ESNEnsemble` batches (per-group parameters on a shared sparsity pattern,
forecast windows batched too, both cost terms on the device), and with `torch.distributed` initialised the
sets are dealt round-robin to the ranks -- BASELINE configs[4]: 64 candidates over the 8 GPUs of a box -- with
one `all_gather` of the costs at the end.  Only the costs leave the GPUs.

The small host-side helpers the reference keeps in `optim.py` / `utils.py` are restated here with the same
behaviour: `inverse_transform` (optim.py:100-137; its docstring example is the known-answer test in
tests/test_host_logic.py) and `update_esn_kwargs` (utils.py:65-100).
"""
import copy
import math

import numpy as np

from . import _lib
from .ensemble import ESNEnsemble
from .esn import ESN as _ESN
from .lazyesn import LazyESN as _LazyESN

_TERMS = ("nrmse", "psd_nrmse")


def _apply(fun, val):
    return fun(val) if isinstance(val, (float, int)) else [fun(x) for x in val]


def transform(params, transformations):
    """log / log10 of the named parameters, the others untouched (reference optim.py:46-97)."""
    return _transform_with(params, transformations, {"log": math.log, "log10": math.log10})


def inverse_transform(transformed_params, transformations):
    """exp / 10** of the named parameters, the others untouched (reference optim.py:100-137)."""
    return _transform_with(transformed_params, transformations, {"log": math.exp, "log10": lambda x: 10.0 ** x})


def _transform_with(params, transformations, funs):
    out = dict(params)
    for key, name in transformations.items():
        if key not in params:
            raise KeyError(f"Could not find '{key}' in optimization parameters, was provided {tuple(params.keys())}")
        if name not in funs:
            raise NotImplementedError(f"transformation {name} unrecognized for parameter {key}, only 'log' and 'log10' "
                                      f"implemented")
        out[key] = _apply(funs[name], params[key])
    return out


def update_esn_kwargs(params, original=None):
    """Constructor kwargs with the optimised parameters filled in: `input_factor` -> `input_kwargs["factor"]` (same for
    `adjacency_` and `bias_`), every other name is a top-level kwarg (reference utils.py:65-100)."""
    esnc = {} if original is None else copy.deepcopy(original)
    for key, val in params.items():
        val = float(val) if isinstance(val, (float, np.floating)) else val
        front, _, back = key.partition("_")
        if back and front in ("input", "adjacency", "bias"):
            esnc.setdefault(f"{front}_kwargs", {})[back] = val
        else:
            esnc[key] = val
    return esnc


def shard(n_sets, rank, world):
    """Indices of the parameter sets rank `rank` of `world` evaluates (round robin: equal counts up to one)."""
    return list(range(rank, n_sets, world))


class CostFunction:
    __slots__ = ("ESN", "train_data", "macro_data", "config", "test_data", "max_batch")

    def __init__(self, ESN, train_data, macro_data, config, test_data=None, max_batch=None):
        if not (isinstance(ESN, type) and issubclass(ESN, _ESN)):
            raise TypeError("CostFunction: ESN must be xesn_b200.ESN or xesn_b200.LazyESN")
        for key in config["macro_training"].get("cost_terms", {}):
            if key not in _TERMS:
                raise NotImplementedError(f"CostFunction.__init__: found '{key}' in config['macro_training']['cost_terms'], "
                                          f"only 'nrmse' and 'psd_nrmse' are implemented")
        self.ESN, self.train_data, self.macro_data = ESN, train_data, macro_data
        self.config, self.test_data = config, test_data
        self.max_batch = max_batch

    # ------------------------------------------------------------------ host logic
    def candidate_kwargs(self, macro_param_sets, is_transformed=True):
        """Constructor kwargs of every parameter set (cost.py:177-190)."""
        sets = np.atleast_2d(np.asarray(macro_param_sets, dtype=np.float64))
        names = tuple(self.config["macro_training"]["parameters"].keys())
        out = []
        for x in sets:
            params = dict(zip(names, (float(v) for v in x)))
            if is_transformed:
                params = inverse_transform(params, self.config["macro_training"].get("transformations", {}))
            out.append(update_esn_kwargs(params, self.config[self.ESN.__name__.lower()]))
        return out

    def batch_size(self, n_reservoir, n_sets):
        """Members per ESNEnsemble: bounded by the Gram matrices (8 N^2 bytes per member) in ~40 GB."""
        if self.max_batch is not None:
            return max(1, min(int(self.max_batch), n_sets))
        return max(1, min(n_sets, int((40 << 30) // (8 * n_reservoir * n_reservoir + 1))))

    # ------------------------------------------------------------------ evaluation
    def _evaluate_members(self, members):
        """Costs of a list of unbuilt ESN members (one ESNEnsemble batch on this rank's GPU)."""
        mt = self.config["macro_training"]
        fc = mt["forecast"]
        ub = mt.get("cost_upper_bound", 1e9)
        ub = 1e9 if ub is None else ub
        ens = ESNEnsemble(members)
        try:
            ens.build()
        except _lib.XesnError as err:
            if "sparsity pattern" not in str(err) or len(members) == 1:
                raise
            # candidates that change the pattern of W (e.g. an optimised connectedness) cannot share a batch
            return np.concatenate([self._evaluate_members([m]) for m in members])
        ens.train(self.train_data, **self.config.get("training", {}))
        costs, _ = ens.cost(self.macro_data, n_steps=fc["n_steps"], n_spinup=fc["n_spinup"], cost_upper_bound=ub,
                            cost_terms=mt.get("cost_terms", {"nrmse": 1.0}))
        return np.asarray(costs, dtype=np.float64)

    # ------------------------------------------------------------------ LazyESN candidates: one model at a time
    def _terms_device(self, pred, truth, space_shape):
        """Per-window total of the cost terms (cost.py:204-210) of ONE model: pred / truth (F, O, S) device tensors."""
        import torch
        lib = _lib.load()
        terms = self.config["macro_training"].get("cost_terms", {"nrmse": 1.0})
        F, O, S = (int(n) for n in pred.shape)
        total = torch.zeros((F, 1), dtype=torch.float64, device=pred.device)
        with torch.cuda.device(pred.device):
            st = torch.cuda.current_stream().cuda_stream
            if "nrmse" in terms:
                fg = torch.empty((F, 1), dtype=torch.float64, device=pred.device)
                scratch = torch.empty(1, dtype=torch.float64, device=pred.device)
                tr = truth.reshape(F * O, S).contiguous()
                _lib.check(lib.xesn_cost_nrmse(pred.contiguous().data_ptr(), tr.data_ptr(), S, 0, F, 1, O, S, 1e300,
                                               fg.data_ptr(), scratch.data_ptr(), st))
                total = total + float(terms["nrmse"]) * fg
            if "psd_nrmse" in terms:
                from .psd import psd_nrmse_device
                total = total + float(terms["psd_nrmse"]) * psd_nrmse_device(pred.reshape(F, 1, O, S), truth, space_shape)
        return total[:, 0]

    def _evaluate_lazy(self, kwargs):
        """`_cost` (cost.py:175-225) of one LazyESN parameter set: build, train, forecast every macro sample, cost terms
        on the device.  Under torch.distributed the groups of the model are sharded over the ranks (LazyESN does
        that itself), so every rank evaluates every set."""
        import torch
        mt = self.config["macro_training"]
        fc = mt["forecast"]
        esn = self.ESN(**copy.deepcopy(kwargs))
        esn.build()
        esn.train(self.train_data, **self.config.get("training", {}))
        preds, truths = [], []
        for truth in self.macro_data:
            pred = esn._predict_core(truth, fc["n_steps"], fc["n_spinup"])              # (n_cells, n_steps + 1) on the device
            t = truth if isinstance(truth, torch.Tensor) else torch.from_numpy(
                np.ascontiguousarray(np.asarray(getattr(truth, "data", truth), dtype=np.float64)))
            space_shape = tuple(int(n) for n in t.shape[:-1])
            t = t.reshape(-1, t.shape[-1])[:, fc["n_spinup"]:fc["n_spinup"] + fc["n_steps"] + 1].to(pred.device)
            preds.append(pred)
            truths.append(t)
        per_window = self._terms_device(torch.stack(preds), torch.stack(truths).contiguous(), space_shape)
        cost = float(per_window.mean().item())
        ub = mt.get("cost_upper_bound", 1e9)
        ub = 1e9 if ub is None else ub
        esn._drop_device_state()
        return ub if (math.isnan(cost) or math.isinf(cost) or cost > ub) else cost

    def __call__(self, macro_param_sets, is_transformed=True):
        """Cost of every parameter set, shape (n_sets, 1) like the reference (cost.py:94-122)."""
        kwargs = self.candidate_kwargs(macro_param_sets, is_transformed)
        n_sets = len(kwargs)
        if issubclass(self.ESN, _LazyESN):
            return np.asarray([self._evaluate_lazy(kw) for kw in kwargs], dtype=np.float64).reshape(-1, 1)
        rank, world, dist = 0, 1, None
        try:
            import torch.distributed as dist_mod
            if dist_mod.is_available() and dist_mod.is_initialized():
                dist, rank, world = dist_mod, dist_mod.get_rank(), dist_mod.get_world_size()
        except ImportError:  # pragma: no cover
            pass
        mine = shard(n_sets, rank, world)
        local = np.full(n_sets, np.nan)
        if mine:
            members = [self.ESN(**kwargs[i]) for i in mine]
            step = self.batch_size(members[0].n_reservoir, len(members))
            for lo in range(0, len(members), step):
                local[mine[lo:lo + step]] = self._evaluate_members(members[lo:lo + step])
        if dist is not None and world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (mine, local[mine].tolist()))
            for idx, vals in gathered:
                local[idx] = vals
        return local.reshape(-1, 1)


    # ------------------------------------------------------------------ diagnostics of one parameter set
    def evaluate(self, parameters, use_test_data=False):
        """`CostFunction.evaluate` of the reference (cost.py:124-172): build and train one model from `parameters`
        (names as in the config, untransformed), forecast every sample of `macro_data` (or `test_data`), and return
        per sample the prediction, the truth and the time-resolved terms named in `cost_terms`: "nrmse" (cost.py:227-254
        with drop_time=False: RMS over space of the error normalised by the truth's temporal standard deviation) and
        "psd_truth" / "psd_prediction" / "psd_nrmse" (psd.py, cost.py:257-272).  Returns an `xarray.Dataset`
        concatenated along "sample" when xarray is importable and the data are DataArrays, else a dict of numpy arrays
        with a leading sample axis."""
        from .esn import _host_array, _is_dataarray, xr
        from .psd import psd as psd_fn
        kwargs = update_esn_kwargs(parameters, self.config[self.ESN.__name__.lower()])
        esn = self.ESN(**kwargs)
        esn.build()
        esn.train(self.train_data, **self.config.get("training", {}))
        if not use_test_data:
            datasets = self.macro_data
            n_spinup = self.config["macro_training"]["forecast"]["n_spinup"]
            n_steps = self.config["macro_training"]["forecast"]["n_steps"]
        else:
            datasets = self.test_data
            n_spinup, n_steps = self.config["testing"]["n_spinup"], self.config["testing"]["n_steps"]
        terms = self.config["macro_training"].get("cost_terms", {"nrmse": 1.0})
        results = []
        for i, truth in enumerate(datasets):
            result = esn.test(truth, n_steps=n_steps, n_spinup=n_spinup)
            if _is_dataarray(truth):
                pred, tr = result["prediction"].values, result["truth"].values
            else:
                pred, tr = np.asarray(result["prediction"]), np.asarray(_host_array(truth))[..., n_spinup:n_spinup + n_steps + 1]
            out = {"prediction": pred, "truth": tr}
            space_axes = tuple(range(pred.ndim - 1))
            if "nrmse" in terms:
                w = 1.0 / np.nanstd(tr, axis=-1, keepdims=True)
                out["nrmse"] = np.sqrt(np.nanmean(((pred - tr) * w) ** 2, axis=space_axes))
            if "psd_nrmse" in terms:
                out["psd_truth"], out["psd_prediction"] = psd_fn(tr), psd_fn(pred)
                with np.errstate(all="ignore"):
                    w = 1.0 / np.nanstd(out["psd_truth"], axis=-1, keepdims=True)
                    out["psd_nrmse"] = np.sqrt(np.nanmean(((out["psd_prediction"] - out["psd_truth"]) * w) ** 2, axis=0))
            if _is_dataarray(truth):
                ds = result
                tdim = "ftime"
                if "nrmse" in out:
                    ds["nrmse"] = xr.DataArray(out["nrmse"], coords={tdim: ds[tdim]}, dims=(tdim,))
                if "psd_nrmse" in out:
                    k = np.arange(1.0, out["psd_truth"].shape[0] + 1)
                    for name in ("psd_truth", "psd_prediction"):
                        ds[name] = xr.DataArray(out[name], coords={"k1d": k, tdim: ds[tdim]}, dims=("k1d", tdim))
                    ds["psd_nrmse"] = xr.DataArray(out["psd_nrmse"], coords={tdim: ds[tdim]}, dims=(tdim,))
                results.append(ds.expand_dims({"sample": [i]}))
            else:
                results.append(out)
        if results and not isinstance(results[0], dict):
            return xr.concat(results, dim="sample")
        return {key: np.stack([r[key] for r in results]) for key in results[0]} if results else {}
