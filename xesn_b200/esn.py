"""`xesn_b200.ESN` -- drop-in for `xesn.ESN` (reference xesn/esn.py:21-455) whose
numerical layer runs on one B200 through libxesn_b200.so.

Kept from the reference: constructor signature and option handling (user dicts
are *filtered*, not merged with defaults, esn.py:341-387), warnings and errors,
attribute names, `build()` (same RandomState draw order: W, Win, bias --
esn.py:162-187), `train(u, y=None, n_spinup=0, batch_size=None)`,
`predict(y, n_steps, n_spinup)`, `test(...)`, `to_xds()`.

Inputs may be `xarray.DataArray`s (time must be the last dimension, named
"time") or bare arrays / torch tensors with time last.  xarray results are
produced when xarray is importable and the input was a DataArray; otherwise
plain numpy arrays are returned.

There is no CPU fallback: every numerical call goes to the CUDA library.
"""
import inspect
import warnings
from decimal import Decimal

import numpy as np
import scipy.sparse

from . import _lib
from .matrix import RandomMatrix, SparseRandomMatrix

try:  # optional packaging layer
    import xarray as xr
except Exception:  # pragma: no cover - xarray is absent in the build container
    xr = None

_MATRIX_KEYS = ("factor", "distribution", "normalization", "is_sparse", "density", "sparsity",
                "connectedness", "format", "random_seed")
_ALLOWED = {"input_kwargs": _MATRIX_KEYS, "adjacency_kwargs": _MATRIX_KEYS,
            "bias_kwargs": ("factor", "distribution", "random_seed")}
_DEFAULTS = {
    "input_kwargs": {"factor": 1.0, "distribution": "uniform", "normalization": "multiply",
                     "is_sparse": False, "random_seed": None},
    "adjacency_kwargs": {"factor": 1.0, "distribution": "uniform", "normalization": "eig",
                         "is_sparse": True, "connectedness": 5, "format": "csr", "random_seed": None},
    "bias_kwargs": {"factor": 1.0, "distribution": "uniform", "random_seed": None},
}


def _torch():
    import torch
    return torch


def _ptr(t):
    return None if t is None else t.data_ptr()


def _is_dataarray(x):
    return xr is not None and isinstance(x, xr.DataArray)


def _time_last(x, method, name):
    dims = getattr(x, "dims", None)
    if dims is not None:
        assert dims[-1] == "time", f"{method}: {name} must have 'time' as the final dimension"


def _host_array(x):
    """numpy view of a DataArray-like / tensor / array input."""
    if hasattr(x, "cpu") and hasattr(x, "numpy"):
        return x.cpu().numpy()
    data = x.data if getattr(x, "dims", None) is not None else x
    if hasattr(data, "compute"):
        data = data.compute()
    return np.asarray(data)


class DeviceReservoir:
    """Owner of the `xesn_reservoir` handle (W as CSR, Win, bias on the device)."""

    def __init__(self, W, Win, bias, leak_rate, device):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.XesnError("xesn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        lib = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        csr = scipy.sparse.csr_matrix(W) if not scipy.sparse.issparse(W) else W.tocsr()
        csr.sum_duplicates()
        csr.sort_indices()
        self.N, self.I = int(Win.shape[0]), int(Win.shape[1])
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        values = np.ascontiguousarray(csr.data, dtype=np.float64)
        win = Win.toarray() if scipy.sparse.issparse(Win) else np.asarray(Win)
        win = np.ascontiguousarray(win, dtype=np.float64)
        b = np.ascontiguousarray(np.asarray(bias, dtype=np.float64).reshape(-1))
        assert b.shape[0] == self.N and csr.shape == (self.N, self.N)
        import ctypes
        self._h = ctypes.c_void_p()
        _lib.check(lib.xesn_reservoir_create(
            ctypes.byref(self._h), self.device.index or 0, self.N, self.I,
            indptr.ctypes.data, indices.ctypes.data, values.ctypes.data, win.ctypes.data, b.ctypes.data,
            float(leak_rate)))
        self._lib = lib
        # A/B switches (environment): XESN_B200_GRAM=f64|i8, XESN_B200_FUSED=0|1
        import os
        if os.environ.get("XESN_B200_GRAM", "i8").lower() in ("f64", "fp64", "dmma", "0"):
            _lib.check(lib.xesn_set_gram_mode(self._h, 0))
        if os.environ.get("XESN_B200_FUSED", "1") == "0":
            _lib.check(lib.xesn_set_fused(self._h, 0))

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.xesn_reservoir_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def auto_chunk(n_reservoir, n_groups, n_time, batch_size, budget_bytes=12 << 30):
    """Internal time-chunk length: bounds the (chunk+1, N) state ring + drive buffers
    (the reference's `batch_size` has the same purpose, esn.py:479-489).  Only the
    floating-point summation order of the Gram terms depends on it.  At most 2048 steps (longer chunks lose more at the
    two ends of the recurrence || Gram pipeline than they save) and at most 12 GB of the 180 GB: every chunk launch
    reads and writes the FP64 Gram triangles once, so for many large reservoirs (32 x N=6000: 18 GB per launch) few long
    chunks beat many short ones -- measured 556 -> 526 ms per bench step from 3 GB to 12 GB, 538 ms at 24 GB."""
    import os
    per_step = 48 * n_reservoir * n_groups  # drive + state ring + int8 digit planes, double-buffered
    cap = int(os.environ.get("XESN_B200_CHUNK_MAX", 2048))
    budget_bytes = int(os.environ.get("XESN_B200_CHUNK_BUDGET", budget_bytes))
    c = max(64, min(cap, budget_bytes // max(per_step, 1)))
    c = (c // 64) * 64
    if batch_size is not None:
        c = min(c, max(int(batch_size), 1))
    return int(max(1, min(c, max(n_time, 1))))


class ESN:
    """Echo State Network on one GPU; see the module docstring."""

    __slots__ = (
        "W", "Win", "n_input", "n_output", "n_reservoir", "leak_rate", "tikhonov_parameter", "bias_vector",
        "input_kwargs", "adjacency_kwargs", "bias_kwargs",
        "device", "_reservoir", "_wout_dev", "_wout_host", "last_info",
    )

    input_factor = property(lambda self: self.input_kwargs["factor"])
    adjacency_factor = property(lambda self: self.adjacency_kwargs["factor"])
    bias_factor = property(lambda self: self.bias_kwargs["factor"])

    def __init__(self, n_input, n_output, n_reservoir, leak_rate, tikhonov_parameter,
                 input_kwargs=None, adjacency_kwargs=None, bias_kwargs=None, *, device=None):
        self.n_input, self.n_output, self.n_reservoir = n_input, n_output, n_reservoir
        self.leak_rate, self.tikhonov_parameter = leak_rate, tikhonov_parameter
        self.input_kwargs = self._set_matrix_options(input_kwargs, "input_kwargs")
        self.adjacency_kwargs = self._set_matrix_options(adjacency_kwargs, "adjacency_kwargs")
        self.bias_kwargs = self._set_matrix_options(bias_kwargs, "bias_kwargs")
        if not self.bias_factor >= 0.0:
            raise ValueError(f"ESN.__init__: bias_factor must be non-negative, got {self.bias_factor}")
        self.device = device
        self.W = self.Win = self.bias_vector = None
        self._reservoir = None
        self._wout_dev = self._wout_host = None
        self.last_info = None

    # ------------------------------------------------------------------ options / repr
    @staticmethod
    def _set_matrix_options(user, name):
        if user is None:
            warnings.warn(f"ESN.__init__: Did not find '{name}' options, using default.")
            return dict(_DEFAULTS[name])
        kept = {}
        for key, val in user.items():
            if key in _ALLOWED[name]:
                kept[key] = val
            else:
                warnings.warn(f"ESN.__init__: '{key}' not allowed for in {name}, ignoring.")
        return kept

    @staticmethod
    def _dictstr(d):
        return "".join(f"        {k:<20s}{v}\n" for k, v in d.items())

    def __str__(self):
        parts = [
            "ESN\n",
            f'    {"n_input:":<24s}{self.n_input}\n',
            f'    {"n_output:":<24s}{self.n_output}\n',
            f'    {"n_reservoir:":<24s}{self.n_reservoir}\n',
            "---\n",
            f'    {"leak_rate:":<24s}{self.leak_rate}\n',
            f'    {"tikhonov_parameter:":<24s}{self.tikhonov_parameter}\n',
            "---\n",
            f"    Input Matrix:\n{self._dictstr(self.input_kwargs)}",
            "---\n",
            f"    Adjacency Matrix:\n{self._dictstr(self.adjacency_kwargs)}",
            "---\n",
            f"    Bias Vector:\n{self._dictstr(self.bias_kwargs)}",
        ]
        return "".join(parts)

    __repr__ = __str__

    # ------------------------------------------------------------------ build
    @staticmethod
    def _make(kwargs, n_rows, n_cols):
        kw = dict(kwargs)
        maker = (SparseRandomMatrix if kw.pop("is_sparse", False) else RandomMatrix)(n_rows=n_rows, n_cols=n_cols, **kw)
        return maker, maker()

    def build(self):
        """Generate W, Win and the bias vector from their seeds (host, scipy) --
        identical matrices to the reference's `ESN.build` (esn.py:147-187)."""
        maker, self.W = self._make(self.adjacency_kwargs, self.n_reservoir, self.n_reservoir)
        if isinstance(maker, SparseRandomMatrix) and maker.density > 0.2:
            warnings.warn("ESN.__init__: adjacency matrix density is >20% but adjacency_kwargs['is_sparse'] = True. "
                          "Performance could suffer from dense matrix operations with scipy.sparse.", RuntimeWarning)
        _, self.Win = self._make(self.input_kwargs, self.n_reservoir, self.n_input)
        self.bias_vector = RandomMatrix(n_rows=1, n_cols=self.n_reservoir, **self.bias_kwargs)().squeeze()
        self._drop_device_state()

    def _drop_device_state(self):
        if self._reservoir is not None:
            self._reservoir.close()
        self._reservoir = None

    def _device_reservoir(self):
        if self.W is None:
            raise _lib.XesnError("ESN: call build() before train()/predict()")
        if self._reservoir is None:
            self._reservoir = DeviceReservoir(self.W, self.Win, self.bias_vector, self.leak_rate, self.device)
        return self._reservoir

    # ------------------------------------------------------------------ Wout (host view of the device readout)
    @property
    def Wout(self):
        if self._wout_host is None and self._wout_dev is not None:
            self._wout_host = self._wout_to_host(self._wout_dev)
        return self._wout_host

    @Wout.setter
    def Wout(self, value):
        self._wout_host = None if value is None else np.asarray(value, dtype=np.float64)
        self._wout_dev = None

    def _wout_to_host(self, wout_dev):
        return wout_dev[0].cpu().numpy()

    def _wout_from_host(self, wout_host):
        torch = _torch()
        res = self._device_reservoir()
        w = np.ascontiguousarray(np.asarray(wout_host, dtype=np.float64).reshape(1, self.n_output, self.n_reservoir))
        return torch.from_numpy(w).to(res.device)

    def _wout_device(self):
        if self._wout_dev is None:
            if self._wout_host is None:
                raise _lib.XesnError("ESN: no readout matrix; call train() first")
            self._wout_dev = self._wout_from_host(self._wout_host)
        return self._wout_dev

    # ------------------------------------------------------------------ data movement helpers
    def _to_device(self, x, name):
        """host array / DataArray / tensor -> contiguous float64 CUDA tensor (H2D from pinned memory)."""
        torch = _torch()
        res = self._device_reservoir()
        if isinstance(x, torch.Tensor):
            t = x.to(device=res.device, dtype=torch.float64)
            return t.contiguous()
        data = getattr(x, "data", x) if getattr(x, "dims", None) is not None else x
        if hasattr(data, "compute"):  # dask-backed DataArray
            data = data.compute()
        arr = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
        host = torch.from_numpy(arr)
        try:
            host = host.pin_memory()
        except Exception:  # pragma: no cover
            pass
        return host.to(res.device, non_blocking=True)

    @staticmethod
    def _stream():
        return _torch().cuda.current_stream().cuda_stream

    # ------------------------------------------------------------------ train
    def train(self, u, y=None, n_spinup=0, batch_size=None):
        """Ridge-regression readout (reference esn.py:190-220 -> `_train_1d` :463-520)."""
        _time_last(u, "ESN.train", "u")
        if y is not None:
            _time_last(y, "ESN.train", "y")
        n_time = (u if y is None else y).shape[1]
        assert n_time >= n_spinup
        u_dev = self._to_device(u, "u")
        y_dev = u_dev if y is None else self._to_device(y, "y")
        assert u_dev.shape[0] == self.n_input and y_dev.shape[0] == self.n_output
        self._wout_dev = self._train_device(u_dev, y_dev, 1, n_time, n_spinup, batch_size)
        self._wout_host = None
        self._check_info()

    def _train_device(self, u_dev, y_dev, n_groups, n_time, n_spinup, batch_size,
                      u_series=None, y_series=None, keepalive=()):
        """u_dev: (G*I, T) rows / y_dev: (G*O, T) rows on the device, or explicit Series."""
        import ctypes
        torch = _torch()
        res = self._device_reservoir()
        lib = res._lib
        N, O, G = self.n_reservoir, self.n_output, n_groups
        chunk = auto_chunk(N, G, max(n_time - n_spinup, n_spinup, 1), batch_size)
        if u_series is None:
            u_series = _lib.Series(_ptr(u_dev), u_dev.stride(0), None, None, None)
        if y_series is None:
            y_series = _lib.Series(_ptr(y_dev), y_dev.stride(0), None, None, None)
        gu = int(bool(u_series.map_dev or u_series.zero_row_dev))
        gy = int(bool(y_series.map_dev or y_series.zero_row_dev))
        with torch.cuda.device(res.device):
            nbytes = lib.xesn_train_scratch_bytes(res.handle, O, G, chunk, gu, gy)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=res.device)
            gram, wout, both = gram_and_readout(G, N, O, res.device)
            info = torch.zeros(G, dtype=torch.int32, device=res.device)
            st = self._stream()

            def accumulate():
                _lib.check(lib.xesn_train_accumulate(
                    res.handle, ctypes.byref(u_series), ctypes.byref(y_series), O, G, n_time, n_spinup, chunk,
                    _ptr(gram), _ptr(wout), _ptr(scratch), nbytes, st))

            accumulate()
            _lib.check(lib.xesn_ridge_solve(
                res.handle, O, G, None, float(self.tikhonov_parameter), _ptr(gram), _ptr(wout), _ptr(info),
                _ptr(scratch), nbytes, st))
            pivoted_fallback(lib, res.handle, accumulate, O, [float(self.tikhonov_parameter)] * G, gram, wout, info,
                             scratch, nbytes, st, type(self).__name__)
            if both is not None:
                wout = wout.clone()  # do not keep the 8 N^2 bytes of the Gram matrices alive through a view
                del gram, both
        self.last_info = info
        del keepalive
        return wout

    def _check_status(self):
        """Raise if a persistent-kernel poll timed out (see xesn_reservoir_status)."""
        import ctypes
        res = self._device_reservoir()
        st = ctypes.c_int32(0)
        _lib.check(res._lib.xesn_reservoir_status(res.handle, ctypes.byref(st)))
        if st.value:
            raise _lib.XesnError("xesn_b200: the recurrence kernel timed out waiting for a peer CTA; results invalid")

    def _check_info(self):
        self._check_status()
        info = self.last_info.cpu().numpy()
        if info.any():
            bad = np.nonzero(info)[0]
            warnings.warn(f"ESN.train: the ridge system of {bad.size} reservoir(s) is not positive definite "
                          f"(first failing minor {int(info[bad[0]])}); the readout of those reservoirs is NaN. "
                          f"Increase tikhonov_parameter.", RuntimeWarning)
            self._wout_dev[torch_index(bad, self._wout_dev)] = float("nan")

    # ------------------------------------------------------------------ predict / test
    def predict(self, y, n_steps, n_spinup):
        """Closed-loop forecast (reference esn.py:223-276).  Returns (n_output, n_steps+1),
        column 0 being y[:, n_spinup]; packaged as a DataArray with an `ftime` axis when the
        input was one."""
        _time_last(y, "ESN.predict", "y")
        n_time = y.shape[1]
        assert n_time >= n_spinup
        if n_time < n_spinup + 1:
            raise IndexError(f"ESN.predict: need n_spinup+1={n_spinup + 1} time samples, got {n_time}")
        assert self.n_input == self.n_output, "ESN.predict: closed loop needs n_input == n_output"
        y_dev = self._to_device(y, "y")
        pred = self._predict_device(y_dev, n_steps, n_spinup)
        out = pred.cpu().numpy()
        self._check_status()
        if _is_dataarray(y):
            fdims, fcoords = self._get_fcoords(y.dims, y.coords, n_steps, n_spinup)
            return xr.DataArray(out, coords=fcoords, dims=fdims)
        return out

    def _predict_device(self, y_dev, n_steps, n_spinup):
        import ctypes
        torch = _torch()
        res = self._device_reservoir()
        lib = res._lib
        N, O, I = self.n_reservoir, self.n_output, self.n_input
        wout = self._wout_device()
        with torch.cuda.device(res.device):
            st = self._stream()
            r = torch.zeros((1, N), dtype=torch.float64, device=res.device)
            series = _lib.Series(_ptr(y_dev), y_dev.stride(0), None, None, None)
            nb = lib.xesn_forced_scratch_bytes(res.handle, 1, n_spinup, 0)
            nb2 = lib.xesn_predict_scratch_bytes(res.handle, O, 1, O)
            scratch = torch.empty(max(nb, nb2), dtype=torch.uint8, device=res.device)
            _lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, n_spinup, 1, _ptr(r), None,
                                           _ptr(scratch), scratch.numel(), st))
            ident = torch.arange(I, dtype=torch.int32, device=res.device)
            fill = torch.zeros(1, dtype=torch.float64, device=res.device)
            field = y_dev[:, n_spinup].contiguous().clone()
            pred = torch.empty((O, n_steps + 1), dtype=torch.float64, device=res.device)
            _lib.check(lib.xesn_predict(res.handle, _ptr(wout), O, 1, O, _ptr(ident), _ptr(fill), _ptr(ident),
                                        _ptr(field), _ptr(r), n_steps, _ptr(pred), _ptr(scratch), scratch.numel(), st))
        return pred

    def test(self, y, n_steps, n_spinup):
        """Prediction plus the matching truth (reference esn.py:279-306)."""
        pred = self.predict(y, n_steps, n_spinup)
        if _is_dataarray(y):
            xds = xr.Dataset()
            xds["prediction"] = pred
            xds["truth"] = xr.DataArray(y.sel(time=xds.prediction.time).data, coords=xds.prediction.coords,
                                        dims=xds.prediction.dims)
            xds.attrs.update(self._get_attrs())
            xds.attrs["esn_type"] = self.__class__.__name__
            xds.attrs["description"] = "Contains a test prediction and matching truth trajectory"
            return xds
        truth = _host_array(y)[..., n_spinup:n_spinup + n_steps + 1]
        return {"prediction": pred, "truth": truth, "attrs": dict(self._get_attrs(), esn_type=self.__class__.__name__)}

    # ------------------------------------------------------------------ packaging (xarray only)
    def to_xds(self):
        """Readout + constructor kwargs as an `xarray.Dataset` (reference esn.py:309-338);
        W and Win are not stored, their seeds are."""
        if self.Wout is None:
            raise Exception("ESN.to_xds: Wout has not been computed yet, so it's not worth storing this model")
        if xr is None:
            raise ImportError("ESN.to_xds needs xarray")
        w = np.asarray(self.Wout).squeeze()
        ds = xr.Dataset()
        ir = np.arange(w.shape[-1])
        iy = np.arange(w.shape[0])
        ds["ir"] = xr.DataArray(ir, coords={"ir": ir}, dims=("ir",),
                                attrs={"description": "logical index for reservoir coordinate"})
        ds["iy"] = xr.DataArray(iy, coords={"iy": iy}, dims=("iy",),
                                attrs={"description": "logical index for flattened output axis"})
        ds["Wout"] = xr.DataArray(w, coords={"iy": ds["iy"], "ir": ds["ir"]}, dims=("iy", "ir"))
        ds.attrs.update(self._get_attrs())
        return ds

    def _get_attrs(self):
        names = [k for k in inspect.signature(type(self).__init__).parameters if k not in ("self", "device")]
        return {k: getattr(self, k) for k in names}

    def _get_fcoords(self, dims, coords, n_steps, n_spinup):
        """Forecast coordinates: drop spin-up, rename time -> ftime (time since the initial
        condition) and keep the absolute time as a coordinate (reference esn.py:398-444)."""
        fdims = tuple("ftime" if d == "time" else d for d in dims)
        window = coords["time"].isel(time=slice(n_spinup, n_spinup + n_steps + 1))
        fcoords = {k: coords[k] for k in coords.keys() if k != "time"}
        fcoords["ftime"] = self._get_ftime(window)
        fcoords["time"] = xr.DataArray(window.values, coords={"ftime": fcoords["ftime"]}, dims="ftime",
                                       attrs=coords["time"].attrs.copy())
        return fdims, fcoords

    @staticmethod
    def _get_ftime(time):
        ftime = time.values - time.values[0]
        if isinstance(time.data[0], float) and "delta_t" in time.attrs:
            digits = abs(Decimal(str(time.attrs["delta_t"])).as_tuple().exponent)
            ftime = np.round(ftime, digits)
        attrs = {"long_name": "forecast_time",
                 "description": "time passed since prediction initial condition, not including ESN spinup"}
        if "units" in time.attrs:
            attrs["units"] = time.attrs["units"]
        return xr.DataArray(ftime, coords={"ftime": ftime}, dims="ftime", attrs=attrs)


def gram_and_readout(n_groups, n_reservoir, n_output, device):
    """Device buffers of the Gram matrices (G, N, N) and the readouts (G, O, N) of a training call.  With few readout rows
    both are views of ONE array (G, N + O, N) -- the readout rows of a group right below its Gram matrix -- which the
    library recognises (wout == gram + N*N elements, include/xesn_b200.h): the solve then carries the right-hand sides
    through the factorisation as extra rows and needs no forward substitution.  Returns (gram, wout, owner or None)."""
    torch = _torch()
    G, N, O = n_groups, n_reservoir, n_output
    if O <= 256:
        both = torch.empty((G, N + O, N), dtype=torch.float64, device=device)
        return both[:, :N, :], both[:, N:, :], both
    return (torch.empty((G, N, N), dtype=torch.float64, device=device),
            torch.empty((G, O, N), dtype=torch.float64, device=device), None)


def pivoted_fallback(lib, handle, accumulate, n_output, betas, gram, wout, info, scratch, nbytes, stream, who):
    """Systems whose Cholesky factorisation failed (info > 0: numerically not positive definite -- tiny
    tikhonov_parameter, rank-deficient Gram) are solved again with partial pivoting, as the reference does on every
    system (scipy `assume_a="sym"` = pivoted LDL^T, cupy = LU; xesn/esn.py:516-520).  The failed factorisation
    destroyed the Gram matrices, so the accumulation is repeated first (rare path; nothing is paid when every system
    is positive definite, apart from reading `info`).  gram (G, N, N), wout (G, O, N), info (G,) device tensors.
    XESN_B200_SOLVE_FALLBACK=0 disables it (the readout of such systems is then NaN)."""
    import os
    torch = _torch()
    bad = torch.nonzero(info).flatten().tolist()
    if not bad or os.environ.get("XESN_B200_SOLVE_FALLBACK", "1") == "0":
        return bad
    warnings.warn(f"{who}.train: {len(bad)} ridge system(s) are not positive definite in double precision "
                  f"(ill-conditioned: tikhonov_parameter too small for this Gram matrix); solved with partial pivoting "
                  f"like scipy.linalg.solve(assume_a='sym')", RuntimeWarning)
    good = wout.clone() if len(bad) < wout.shape[0] else None
    accumulate()
    for g in bad:
        _lib.check(lib.xesn_ridge_solve_pivoted(handle, n_output, float(betas[g]), gram[g].data_ptr(), wout[g].data_ptr(),
                                                info[g:g + 1].data_ptr(), _ptr(scratch), nbytes, stream))
    if good is not None:
        keep = torch.ones(wout.shape[0], dtype=torch.bool, device=wout.device)
        keep[torch.as_tensor(bad, device=wout.device)] = False
        wout[keep] = good[keep]
    return bad


def torch_index(idx, like):
    torch = _torch()
    return torch.as_tensor(np.asarray(idx), device=like.device, dtype=torch.long)
