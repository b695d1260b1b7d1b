"""ESNEnsemble -- G eager ESNs that differ only in their hyper-parameters, trained and forecast as ONE batch.

The reference's macro training (`xesn/cost.py:190-203`, `xesn/optim.py:97-137`) evaluates every candidate
of the search by building a fresh `ESN(**params)` from the SAME seeds with its own `input_factor`,
`adjacency_factor`, `bias_factor`, `leak_rate` and `tikhonov_parameter`, training it and forecasting a few
test windows -- one candidate after the other.  A reservoir of a few thousand units is latency-bound on a
B200 (one step is a serial dependency), so the candidates are independent work that belongs side by side:
they share the sparsity pattern of W (same seed), so one device reservoir holds G value sets
(`xesn_reservoir_set_group_params`, include/xesn_b200.h) and every kernel of the hot path runs the G
reservoirs as groups -- the recurrence of all of them in one persistent kernel, their Gram matrices on the
tensor cores back to back, their ridge systems as one batched Cholesky, their forecasts in one loop.

    esns = [ESN(40, 40, 4000, leak, beta, input_kwargs=..., ...) for leak, beta, ... in candidates]
    ens = ESNEnsemble(esns); ens.build()
    ens.train(u, n_spinup=0)                 # every member's Wout, as ESN.train would have produced it
    preds = ens.predict(y, n_steps, n_spinup)  # (G, n_output, n_steps+1)

Each member remains a normal `ESN` (its `Wout` is filled in), so the results can be checked member by member
against the single-reservoir path and against the reference.
"""
import ctypes

import numpy as np
import scipy.sparse

from . import _lib
from .esn import ESN, DeviceReservoir, _ptr, _time_last, _torch, auto_chunk, gram_and_readout, pivoted_fallback
from .matrix import RandomMatrix, SparseRandomMatrix, _dense_scale, _sparse_scale


class ESNEnsemble:
    # unscaled random matrices + their normalisation scalars per (kind, shape, options incl. seed); see build_host
    _MATRIX_CACHE = {}

    def __init__(self, members, device=None):
        members = list(members)
        assert members, "ESNEnsemble: at least one member"
        m0 = members[0]
        for m in members:
            assert isinstance(m, ESN)
            assert (m.n_input, m.n_output, m.n_reservoir) == (m0.n_input, m0.n_output, m0.n_reservoir), \
                "ESNEnsemble: members must have the same sizes"
        self.members = members
        self.n_input, self.n_output, self.n_reservoir = m0.n_input, m0.n_output, m0.n_reservoir
        self.device = device
        self._reservoir = None
        self._wout_dev = None
        self.last_info = None

    def __len__(self):
        return len(self.members)

    # ------------------------------------------------------------------ build
    def build(self):
        """Build every member on the host exactly as `ESN.build` does (reference esn.py:147-187; same seeds,
        own factors) and upload them as the parameter groups of one device reservoir."""
        self.build_host()
        self._upload()

    def build_host(self):
        """Host part of build().  The members of a search generation are built from the SAME seeds and differ only
        in their factors (cost.py:190-193), so the unscaled random matrix and its normalisation constant (the
        25 s of `scipy.sparse.random` + ARPACK `svds` at N=16000) are computed once per
        distinct (shape, distribution, normalization, density, format, seed) and every member then applies its own
        `factor / scale * A` -- the reference's expression (matrix.py:183-193, :296-316), hence the same matrices
        up to ARPACK's ~1e-15 run-to-run jitter in `scale`, which the cache removes."""
        cache = ESNEnsemble._MATRIX_CACHE     # kept across calls: successive batches of a search reuse the draws
        while len(cache) > 16:
            cache.pop(next(iter(cache)))

        def make(kwargs, n_rows, n_cols):
            kw = dict(kwargs)
            sparse = kw.pop("is_sparse", False)
            factor = kw.pop("factor")
            key = (sparse, n_rows, n_cols, tuple(sorted((k, repr(v)) for k, v in kw.items())))
            if key not in cache:
                maker = (SparseRandomMatrix if sparse else RandomMatrix)(n_rows=n_rows, n_cols=n_cols, factor=1.0, **kw)
                A = maker.create_matrix()
                scale = (_sparse_scale if sparse else _dense_scale)(A, maker.normalization)
                cache[key] = (maker, A, scale)
            maker, A, scale = cache[key]
            if sparse:
                B = A.copy()
                B.data = factor / scale * A.data
                return maker, B
            return maker, factor / scale * A

        for m in self.members:
            if m.W is not None:
                continue
            maker, m.W = make(m.adjacency_kwargs, m.n_reservoir, m.n_reservoir)
            if isinstance(maker, SparseRandomMatrix) and maker.density > 0.2:
                import warnings
                warnings.warn("ESN.__init__: adjacency matrix density is >20% but adjacency_kwargs['is_sparse'] = True. "
                              "Performance could suffer from dense matrix operations with scipy.sparse.", RuntimeWarning)
            _, m.Win = make(m.input_kwargs, m.n_reservoir, m.n_input)
            _, b = make(dict(m.bias_kwargs), 1, m.n_reservoir)
            m.bias_vector = b.squeeze()
            m._drop_device_state()

    def _upload(self):
        csrs = []
        for m in self.members:
            c = scipy.sparse.csr_matrix(m.W) if not scipy.sparse.issparse(m.W) else m.W.tocsr()
            c.sum_duplicates()
            c.sort_indices()
            csrs.append(c)
        c0 = csrs[0]
        for c in csrs[1:]:
            if not (np.array_equal(c.indptr, c0.indptr) and np.array_equal(c.indices, c0.indices)):
                raise _lib.XesnError("ESNEnsemble: the members' adjacency matrices must share one sparsity pattern "
                                     "(same adjacency seed, density and size)")
        m0 = self.members[0]
        self._reservoir = DeviceReservoir(c0, m0.Win, m0.bias_vector, m0.leak_rate, self.device)
        values = np.ascontiguousarray(np.stack([c.data for c in csrs]), dtype=np.float64)
        win = np.ascontiguousarray(np.stack([np.asarray(m.Win, dtype=np.float64) for m in self.members]))
        bias = np.ascontiguousarray(np.stack([np.asarray(m.bias_vector, dtype=np.float64).reshape(-1)
                                              for m in self.members]))
        leak = np.ascontiguousarray([float(m.leak_rate) for m in self.members], dtype=np.float64)
        res = self._reservoir
        _lib.check(res._lib.xesn_reservoir_set_group_params(
            res.handle, len(self.members), values.ctypes.data, win.ctypes.data, bias.ctypes.data, leak.ctypes.data))

    def _device_reservoir(self):
        if self._reservoir is None:
            raise _lib.XesnError("ESNEnsemble: call build() first")
        return self._reservoir

    def _to_device(self, x):
        torch = _torch()
        res = self._device_reservoir()
        if isinstance(x, torch.Tensor):
            return x.to(device=res.device, dtype=torch.float64).contiguous()
        data = getattr(x, "data", x) if getattr(x, "dims", None) is not None else x
        host = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64)))
        try:
            host = host.pin_memory()
        except Exception:  # pragma: no cover
            pass
        return host.to(res.device, non_blocking=True)

    def _shared_series(self, x_dev, rows):
        """All G groups read the same `rows` rows of x_dev: an index map instead of G copies."""
        torch = _torch()
        G = len(self.members)
        idx = torch.arange(rows, dtype=torch.int32, device=x_dev.device).repeat(G)
        return _lib.Series(_ptr(x_dev), x_dev.stride(0), _ptr(idx), None, None), idx

    # ------------------------------------------------------------------ train
    def train(self, u, y=None, n_spinup=0, batch_size=None):
        """`ESN.train(u, y, n_spinup, batch_size)` of every member (reference esn.py:190-220 -> `_train_1d`)."""
        _time_last(u, "ESNEnsemble.train", "u")
        torch = _torch()
        res = self._device_reservoir()
        lib = res._lib
        G, N, O, I = len(self.members), self.n_reservoir, self.n_output, self.n_input
        u_dev = self._to_device(u)
        y_dev = u_dev if y is None else self._to_device(y)
        n_time = y_dev.shape[1]
        assert n_time >= n_spinup and u_dev.shape[0] == I and y_dev.shape[0] == O
        chunk = auto_chunk(N, G, max(n_time - n_spinup, n_spinup, 1), batch_size)
        u_series, ku = self._shared_series(u_dev, I)
        y_series, ky = self._shared_series(y_dev, O)
        with torch.cuda.device(res.device):
            st = torch.cuda.current_stream().cuda_stream
            nbytes = lib.xesn_train_scratch_bytes(res.handle, O, G, chunk, 1, 1)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=res.device)
            gram, wout, both = gram_and_readout(G, N, O, res.device)
            info = torch.zeros(G, dtype=torch.int32, device=res.device)
            beta = torch.tensor([float(m.tikhonov_parameter) for m in self.members], dtype=torch.float64,
                                device=res.device)
            def accumulate():
                _lib.check(lib.xesn_train_accumulate(
                    res.handle, ctypes.byref(u_series), ctypes.byref(y_series), O, G, n_time, n_spinup, chunk,
                    _ptr(gram), _ptr(wout), _ptr(scratch), nbytes, st))

            accumulate()
            _lib.check(lib.xesn_ridge_solve(res.handle, O, G, _ptr(beta), 0.0, _ptr(gram), _ptr(wout), _ptr(info),
                                            _ptr(scratch), nbytes, st))
            pivoted_fallback(lib, res.handle, accumulate, O, [float(m.tikhonov_parameter) for m in self.members], gram,
                             wout, info, scratch, nbytes, st, "ESNEnsemble")
            if both is not None:
                wout = wout.clone()   # the Gram matrices must not stay alive through a view
            del gram, both, ku, ky
        self._wout_dev = wout
        self.last_info = info
        bad = np.nonzero(info.cpu().numpy())[0]
        if bad.size:
            import warnings
            warnings.warn(f"ESNEnsemble.train: the ridge system of member(s) {bad.tolist()} is not positive definite; "
                          f"their readout is NaN", RuntimeWarning)
            wout[torch.as_tensor(bad, device=wout.device, dtype=torch.long)] = float("nan")
        for g, m in enumerate(self.members):
            m._wout_dev = wout[g:g + 1]
            m._wout_host = None
            m.last_info = info[g:g + 1]
        return self

    @property
    def Wout(self):
        """(G, n_output, n_reservoir) on the host."""
        return None if self._wout_dev is None else self._wout_dev.cpu().numpy()

    # ------------------------------------------------------------------ predict
    def predict(self, y, n_steps, n_spinup):
        """`ESN.predict(y, n_steps, n_spinup)` of every member (reference esn.py:223-276): (G, n_output, n_steps+1)."""
        _time_last(y, "ESNEnsemble.predict", "y")
        assert self.n_input == self.n_output, "closed loop needs n_input == n_output"
        y_dev = self._to_device(y)
        assert y_dev.shape[1] >= n_spinup + 1
        return self._predict_device(y_dev, n_steps, n_spinup).cpu().numpy()

    def predict_windows(self, windows, n_steps, n_spinup):
        """Several test windows at once (the reference scores a candidate on `n_samples` forecasts, cost.py:196-203):
        windows is a sequence of F arrays (n_output, T >= n_spinup+1) of equal length, or one array (F, n_output, T).
        Returns (F, G, n_output, n_steps+1); equals [predict(w, ...) for w in windows] with F x G reservoirs in one loop."""
        torch = _torch()
        res = self._device_reservoir()
        ws = [self._to_device(w) for w in windows]
        F, O = len(ws), self.n_output
        assert F >= 1 and all(w.shape == ws[0].shape for w in ws) and ws[0].shape[0] == O
        assert ws[0].shape[1] >= n_spinup + 1
        stacked = torch.cat(ws, dim=0) if F > 1 else ws[0]          # rows f*O + o
        return self._predict_device(stacked, n_steps, n_spinup, F).cpu().numpy()

    def cost(self, windows, n_steps, n_spinup, cost_upper_bound=1e9, cost_terms=None, space_shape=None):
        """The reference's macro-training cost of every member (`cost._cost`, xesn/cost.py:195-225): forecast the F
        windows, per window the factor-weighted sum of the cost terms -- "nrmse" (cost.py:227-254) and, for fields with
        one spatial axis, "psd_nrmse" (cost.py:257-272 over psd.py:12-65); default {"nrmse": 1} like the reference --
        then the average over the windows and the clamp to `cost_upper_bound`.  Forecasts, spectra and errors stay on
        the device; returns (costs (G,), per-window totals (F, G)) as numpy arrays.  `space_shape`: the spatial shape the
        n_output variables unflatten to (C order) when the field has two or three spatial axes -- the PSD term is then
        the radially binned spectrum of psd.py:42-65 (xesn_b200/psd.py); default: one spatial axis."""
        torch = _torch()
        res = self._device_reservoir()
        terms = {"nrmse": 1.0} if cost_terms is None else dict(cost_terms)
        for key in terms:
            if key not in ("nrmse", "psd_nrmse"):
                raise NotImplementedError(f"ESNEnsemble.cost: found '{key}' in cost_terms, only 'nrmse' and 'psd_nrmse' "
                                          f"are implemented")
        ws = [self._to_device(w) for w in windows]
        F, O, G = len(ws), self.n_output, len(self.members)
        assert F >= 1 and all(w.shape == ws[0].shape for w in ws) and ws[0].shape[0] == O
        assert ws[0].shape[1] >= n_spinup + n_steps + 1, "the truth must cover the forecast"
        stacked = torch.cat(ws, dim=0).contiguous() if F > 1 else ws[0]
        pred = self._predict_device(stacked, n_steps, n_spinup, F).contiguous()
        with torch.cuda.device(res.device):
            st = torch.cuda.current_stream().cuda_stream
            fg = torch.empty((F, G), dtype=torch.float64, device=res.device)
            cost = torch.empty(G, dtype=torch.float64, device=res.device)
            _lib.check(res._lib.xesn_cost_nrmse(_ptr(pred), _ptr(stacked), stacked.stride(0), n_spinup, F, G, O,
                                                n_steps + 1, float(cost_upper_bound), _ptr(fg), _ptr(cost), st))
            if terms != {"nrmse": 1.0}:
                total = float(terms.get("nrmse", 0.0)) * fg if "nrmse" in terms else torch.zeros_like(fg)
                if "psd_nrmse" in terms:
                    if space_shape is not None and len(tuple(space_shape)) > 1:
                        from .psd import psd_nrmse_device
                        truth = stacked.view(F, O, -1)[:, :, n_spinup:n_spinup + n_steps + 1].contiguous()
                        pfg = psd_nrmse_device(pred.view(F, G, O, n_steps + 1), truth, space_shape)
                    else:
                        pfg = torch.empty((F, G), dtype=torch.float64, device=res.device)
                        _lib.check(res._lib.xesn_cost_psd_nrmse(_ptr(pred), _ptr(stacked), stacked.stride(0), n_spinup, F,
                                                                G, O, n_steps + 1, _ptr(pfg), st))
                    total = total + float(terms["psd_nrmse"]) * pfg
                fg = total
                cost = fg.mean(dim=0)
                ub = torch.full_like(cost, float(cost_upper_bound))
                cost = torch.where(torch.isnan(cost) | torch.isinf(cost) | (cost > ub), ub, cost)
        return cost.cpu().numpy(), fg.cpu().numpy()

    def _predict_device(self, y_dev, n_steps, n_spinup, n_windows=1):
        """y_dev: (n_windows * n_output, T) rows on the device -> (n_windows, G, O, n_steps+1), or (G, O, n_steps+1)
        for a single window."""
        torch = _torch()
        F = int(n_windows)
        res = self._device_reservoir()  # group f*G + g uses parameter set g (xesn_reservoir_set_group_params)
        lib = res._lib
        if self._wout_dev is None:
            raise _lib.XesnError("ESNEnsemble: no readout; call train() first")
        Gm, N, O, I = len(self.members), self.n_reservoir, self.n_output, self.n_input
        G = F * Gm
        wout = self._wout_dev if F == 1 else self._wout_dev.repeat(F, 1, 1)
        with torch.cuda.device(res.device):
            st = torch.cuda.current_stream().cuda_stream
            r = torch.zeros((G, N), dtype=torch.float64, device=res.device)
            # group f*Gm + g reads the rows of window f
            rows = (torch.arange(F, dtype=torch.int32, device=res.device).repeat_interleave(Gm * I) * I
                    + torch.arange(I, dtype=torch.int32, device=res.device).repeat(G))
            series, keep = _lib.Series(_ptr(y_dev), y_dev.stride(0), _ptr(rows), None, None), rows
            nb = lib.xesn_forced_scratch_bytes(res.handle, G, n_spinup, 1)
            nb2 = lib.xesn_predict_scratch_bytes(res.handle, O, G, G * O)
            scratch = torch.empty(max(nb, nb2), dtype=torch.uint8, device=res.device)
            _lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, n_spinup, G, _ptr(r), None,
                                           _ptr(scratch), scratch.numel(), st))
            # every member forecasts its own copy of the field: cells [g*O, (g+1)*O)
            cells = torch.arange(G * O, dtype=torch.int32, device=res.device)
            fill = torch.zeros(1, dtype=torch.float64, device=res.device)
            field = y_dev[:, n_spinup].contiguous().view(F, 1, O).expand(F, Gm, O).contiguous().view(-1)
            pred = torch.empty((G * O, n_steps + 1), dtype=torch.float64, device=res.device)
            _lib.check(lib.xesn_predict(res.handle, _ptr(wout), O, G, G * O, _ptr(cells), _ptr(fill),
                                        _ptr(cells), _ptr(field), _ptr(r), n_steps, _ptr(pred), _ptr(scratch),
                                        scratch.numel(), st))
            del keep
        return pred.view(Gm, O, n_steps + 1) if F == 1 else pred.view(F, Gm, O, n_steps + 1)
