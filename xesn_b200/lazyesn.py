"""`xesn_b200.LazyESN` -- drop-in for `xesn.LazyESN` (reference xesn/lazyesn.py:18-261).

The reference chunks the domain with dask, builds halos with
`dask.array.overlap.overlap` and maps `_train_nd` / `_spinup` / `_update_nd` /
`_readout` over the blocks (lazyesn.py:264-327).  Here every chunk ("group") is
one independent reservoir of a single batched CUDA computation: all groups share
(W, Win, b), the halos are index gathers (xesn_b200/halo.py) and the per-group
ridge systems are factorised as one batch.  With `torch.distributed` initialised
(one process per GPU) the groups are split in contiguous slabs over the ranks:
training needs no communication; during `predict` the ranks exchange the
freshly predicted cells once per step.
"""
import ctypes
import os
import warnings
from functools import reduce

import numpy as np

from . import _lib
from .esn import ESN, _is_dataarray, _ptr, _time_last, _torch, auto_chunk, xr
from . import halo as _halo


def _prod(seq):
    return reduce(lambda a, b: a * b, seq)


# set by measurement code that needs an unsharded run inside a distributed job (bench.py's multi-GPU parity
# check runs the whole domain on rank 0 alone while the other ranks wait)
_FORCE_SINGLE_RANK = False


def _dist():
    """(rank, world) of the torch.distributed job, (0, 1) when not initialised."""
    if _FORCE_SINGLE_RANK:
        return 0, 1
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_groups(n_groups, rank, world):
    """Contiguous slab [g0, g1) of the C-ordered group list owned by `rank`."""
    if world == 1:
        return 0, n_groups
    if n_groups % world:
        raise ValueError(f"LazyESN: {n_groups} groups do not divide over {world} ranks")
    per = n_groups // world
    return rank * per, (rank + 1) * per


def exchange_owned(field, owned, rank, gathered):
    """All ranks publish the cells they own: field[owned[r]] of rank r is copied into `field`
    of every rank.  `owned`: (world, cells_per_rank) int64 tensor, `gathered`: scratch of the
    same shape in field's dtype.  One collective per forecast step -- the neighbour exchange
    dask's `overlap` performs in the reference (lazyesn.py:235)."""
    import torch.distributed as dist
    mine = field[owned[rank]].contiguous()
    try:
        dist.all_gather_into_tensor(gathered, mine)
    except (RuntimeError, NotImplementedError):  # backends without the fused variant
        parts = [gathered[r] for r in range(gathered.shape[0])]
        dist.all_gather(parts, mine)
    field[owned.reshape(-1)] = gathered.reshape(-1)
    return field


class LazyESN(ESN):
    __slots__ = ("esn_chunks", "overlap", "persist", "boundary",
                 "_grid", "_dims", "_maps", "_wout_blocks_host", "_p2p")

    @property
    def output_chunks(self):
        return self.esn_chunks

    @property
    def input_chunks(self):
        return {k: self.output_chunks[k] + 2 * self.overlap[k] for k in self.output_chunks.keys()}

    @property
    def ndim_state(self):
        """Number of non-time axes."""
        return len(self.overlap) - 1

    @property
    def _r_chunks(self):
        return tuple(1 for _ in range(self.ndim_state - 1)) + (self.n_reservoir,)

    @property
    def _Wout_chunks(self):
        return (self.n_output, self.n_reservoir) + tuple(1 for _ in range(self.ndim_state - 2))

    def __init__(self, esn_chunks, overlap, boundary, n_reservoir, leak_rate, tikhonov_parameter, persist=False,
                 input_kwargs=None, adjacency_kwargs=None, bias_kwargs=None, *, device=None):
        self.overlap, self.boundary, self.persist, self.esn_chunks = overlap, boundary, persist, esn_chunks
        # like the reference (lazyesn.py:98-105) the caller's dicts get a trivial "time" entry
        self.overlap.setdefault("time", 0)
        self.esn_chunks.setdefault("time", -1)
        if any(c < 0 for k, c in self.esn_chunks.items() if k != "time"):
            raise ValueError("LazyESN.__init__: Cannot have negative numbers or Nones in non-temporal axis "
                             "locations of esn_chunks. Provide the actual value please.")
        n_output = _prod([c for k, c in self.output_chunks.items() if k != "time"])
        n_input = _prod([c for k, c in self.input_chunks.items() if k != "time"])
        super().__init__(n_input=n_input, n_output=n_output, n_reservoir=n_reservoir, leak_rate=leak_rate,
                         tikhonov_parameter=tikhonov_parameter, input_kwargs=input_kwargs,
                         adjacency_kwargs=adjacency_kwargs, bias_kwargs=bias_kwargs, device=device)
        if sum(1 for d in self.overlap.values() if d > 0) > 2:
            raise NotImplementedError("LazyESN.__init__: cannot overlap more than 2 axes")
        self._grid = self._dims = self._maps = self._wout_blocks_host = None
        self._p2p = None

    def __str__(self):
        bdy = "\n" + self._dictstr(self.boundary) if isinstance(self.boundary, dict) else str(self.boundary)
        head = ("LazyESN\n"
                f"    input_chunks:\n{self._dictstr(self.input_chunks)}---\n"
                f"    output_chunks:\n{self._dictstr(self.output_chunks)}---\n"
                f"    overlap:\n{self._dictstr(self.overlap)}---\n"
                f'    {"boundary:":<24s}{bdy}\n---\n')
        return head + super().__str__().replace("ESN\n", "", 1)

    __repr__ = __str__

    def _device_reservoir(self):
        fresh = self._reservoir is None
        res = super()._device_reservoir()
        if fresh:
            # LazyESN rule: NaN inputs of the closed loop read as 0 (`_prepare_1d_inputs`, lazyesn.py:357-365);
            # the eager ESN lets them propagate, which is the library default
            _lib.check(res._lib.xesn_set_nan_input_mask(res.handle, 1))
        return res

    def _drop_device_state(self):
        self._p2p_release()
        super()._drop_device_state()

    def __del__(self):  # pragma: no cover
        try:
            self._p2p_release()
        except Exception:
            pass

    # ------------------------------------------------------------------ geometry
    def _space_dims(self, y):
        """Names of the space axes in array order (time last)."""
        dims = getattr(y, "dims", None)
        if dims is not None:
            return tuple(d for d in dims[:-1])
        names = tuple(k for k in self.esn_chunks.keys() if k != "time")
        if len(names) != y.ndim - 1:
            raise ValueError(f"LazyESN: data has {y.ndim - 1} space axes but esn_chunks names {names}")
        return names

    def _geometry(self, y):
        dims = self._space_dims(y)
        shape = tuple(int(n) for n in y.shape[:-1])
        chunks = tuple(int(self.esn_chunks[d]) for d in dims)
        depth = tuple(int(self.overlap[d]) for d in dims)
        key = (dims, shape)
        if self._maps is None or self._maps[0] != key:
            rules = _halo.normalise_boundary(self.boundary, dims, depth)
            grid, in_map, out_map, fill = _halo.build_maps(shape, chunks, depth, rules)
            assert in_map.shape[1] == self.n_input and out_map.shape[1] == self.n_output
            self._maps = (key, grid, in_map, out_map, fill)
        self._dims, self._grid = dims, self._maps[1]
        return self._maps[1:]

    def _my_groups(self, n_groups):
        rank, world = _dist()
        return partition_groups(n_groups, rank, world)

    # ------------------------------------------------------------------ Wout in the reference's block layout
    @property
    def Wout(self):
        if self._wout_blocks_host is None and self._wout_dev is not None:
            self._wout_blocks_host = _halo.wout_to_blocks(self._gather_wout_groups(), self._grid)
        return self._wout_blocks_host

    @Wout.setter
    def Wout(self, value):
        self._wout_blocks_host = None if value is None else np.asarray(value, dtype=np.float64)
        self._wout_dev = None

    @property
    def Wout_groups(self):
        """(G, O, N) readouts, group index in C order over the group grid."""
        return self._gather_wout_groups()

    def _gather_wout_groups(self):
        torch = _torch()
        w = self._wout_dev
        rank, world = _dist()
        if world > 1:
            import torch.distributed as dist
            parts = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(parts, w.contiguous())
            w = torch.cat(parts, dim=0)
        return w.cpu().numpy()

    def _wout_device_groups(self, g0, g1):
        if self._wout_dev is None:
            if self._wout_blocks_host is None:
                raise _lib.XesnError("LazyESN: no readout matrix; call train() first")
            if self._grid is None:
                raise _lib.XesnError("LazyESN: geometry unknown; call predict() with data first")
            w = _halo.blocks_to_wout(self._wout_blocks_host, self._grid, self.n_output, self.n_reservoir)
            torch = _torch()
            self._wout_dev = torch.from_numpy(np.ascontiguousarray(w[g0:g1])).to(self._device_reservoir().device)
        return self._wout_dev

    # ------------------------------------------------------------------ train
    def train(self, y, n_spinup=0, batch_size=None):
        """Per-group ridge regression (reference lazyesn.py:148-187 + `_train_nd` :264-298)."""
        _time_last(y, "LazyESN.train", "y")
        torch = _torch()
        grid, in_map, out_map, fill = self._geometry(y)
        n_time = int(y.shape[-1])
        assert n_time >= n_spinup
        G_all = in_map.shape[0]
        g0, g1 = self._my_groups(G_all)
        res = self._device_reservoir()
        dev = res.device

        field_dev, in_l, out_l, masked = self._upload_local(y, in_map[g0:g1], out_map[g0:g1], slice(None))
        zero_u = masked[np.maximum(in_l, 0)] & (in_l >= 0)
        zero_y = masked[out_l]
        t_in = torch.from_numpy(in_l.reshape(-1)).to(dev)
        t_out = torch.from_numpy(out_l.reshape(-1)).to(dev)
        t_fill = torch.from_numpy(fill).to(dev)
        t_zu = torch.from_numpy(zero_u.reshape(-1).astype(np.uint8)).to(dev) if zero_u.any() else None
        t_zy = torch.from_numpy(zero_y.reshape(-1).astype(np.uint8)).to(dev) if zero_y.any() else None
        ld = field_dev.stride(0)
        u_series = _lib.Series(_ptr(field_dev), ld, _ptr(t_in), _ptr(t_fill), _ptr(t_zu))
        y_series = _lib.Series(_ptr(field_dev), ld, _ptr(t_out), _ptr(t_fill), _ptr(t_zy))
        wout = self._train_device(None, None, g1 - g0, n_time, n_spinup, batch_size, u_series, y_series,
                                  keepalive=(field_dev, t_in, t_out, t_fill, t_zu, t_zy))
        if zero_y.any():
            # masked target cells: the reference leaves NaN rows (lazyesn.py:292-297)
            wout[torch.from_numpy(zero_y).to(dev)] = float("nan")
        self._wout_dev = wout
        self._wout_blocks_host = None
        self._check_info_groups()

    def _check_info_groups(self):
        self._check_status()
        info = self.last_info.cpu().numpy()
        if info.any():
            bad = np.nonzero(info)[0]
            warnings.warn(f"LazyESN.train: the ridge system of {bad.size} group(s) is not positive definite; "
                          f"their readout is NaN. Increase tikhonov_parameter.", RuntimeWarning)
            self._wout_dev[_torch().as_tensor(bad, device=self._wout_dev.device)] = float("nan")

    def _upload_local(self, y, in_map, out_map, tslice):
        """Upload only the cells this rank's groups touch; returns the device rows, the maps
        re-indexed to those rows and the per-row mask `isnan(first time sample)`.

        Host inputs are copied run by run: the cells of a slab of groups are a few runs of consecutive rows
        of the flattened field (the slab itself plus its halo strips), so every run is one strided
        host-to-device copy straight from the caller's (ideally pinned) buffer -- no gathered host copy."""
        torch = _torch()
        dev = self._device_reservoir().device
        used = np.unique(np.concatenate([in_map[in_map >= 0].ravel(), out_map.ravel()]))
        if isinstance(y, torch.Tensor):
            flat = y.reshape(-1, y.shape[-1])[:, tslice]
        else:
            data = getattr(y, "data", y) if getattr(y, "dims", None) is not None else y
            if hasattr(data, "compute"):
                data = data.compute()
            arr = np.asarray(data)
            if arr.dtype != np.float64:
                arr = arr.astype(np.float64)
            flat = torch.from_numpy(arr.reshape(-1, arr.shape[-1]))[:, tslice]
        if flat.is_cuda:
            rows = flat[torch.from_numpy(used).to(flat.device)].to(device=dev, dtype=torch.float64).contiguous()
            masked_rows = torch.isnan(rows[:, 0]).cpu().numpy()
        else:
            breaks = np.nonzero(np.diff(used) != 1)[0] + 1
            starts = np.concatenate([[0], breaks])
            ends = np.concatenate([breaks, [used.size]])
            if flat.dtype != torch.float64 or len(starts) > 64:
                host = flat[torch.from_numpy(used)].to(torch.float64).contiguous()
                rows = host.to(dev, non_blocking=True)
            else:
                rows = torch.empty((used.size, flat.shape[1]), dtype=torch.float64, device=dev)
                for a, b in zip(starts, ends):
                    rows[a:b].copy_(flat[int(used[a]):int(used[b - 1]) + 1], non_blocking=True)
            masked_rows = torch.isnan(flat[torch.from_numpy(used), 0]).numpy() if flat.shape[1] else np.zeros(used.size, bool)
        lookup = np.full(int(used.max()) + 1 if used.size else 1, -1, dtype=np.int64)
        lookup[used] = np.arange(used.size)
        in_l = np.where(in_map >= 0, lookup[np.maximum(in_map, 0)], in_map).astype(np.int32)
        out_l = lookup[out_map].astype(np.int32)
        return rows, in_l, out_l, masked_rows

    # ------------------------------------------------------------------ predict
    def predict(self, y, n_steps, n_spinup):
        """Closed-loop forecast of the whole domain (reference lazyesn.py:190-248)."""
        _time_last(y, "LazyESN.predict", "y")
        assert y.shape[-1] >= n_spinup + 1
        pred = self._predict_core(y, n_steps, n_spinup)
        out = pred.cpu().numpy().reshape(tuple(y.shape[:-1]) + (n_steps + 1,))
        self._check_status()
        if _is_dataarray(y):
            fdims, fcoords = self._get_fcoords(y.dims, y.coords, n_steps, n_spinup)
            return xr.DataArray(out, coords=fcoords, dims=fdims)
        return out

    def _predict_core(self, y, n_steps, n_spinup):
        """Forecast on the device: returns the (n_cells, n_steps+1) trajectory tensor."""
        torch = _torch()
        grid, in_map, out_map, fill = self._geometry(y)
        G_all = in_map.shape[0]
        g0, g1 = self._my_groups(G_all)
        G = g1 - g0
        rank, world = _dist()
        res = self._device_reservoir()
        lib, dev = res._lib, res.device
        N, O, I = self.n_reservoir, self.n_output, self.n_input
        n_cells = int(np.prod(y.shape[:-1]))
        wout = self._wout_device_groups(g0, g1)

        # ---- spin-up on the halo'd truth (lazyesn.py:215-221, `_spinup` :301-316)
        truth_dev, in_l, out_l, masked = self._upload_local(y, in_map[g0:g1], out_map[g0:g1], slice(0, n_spinup + 1))
        zero_u = masked[np.maximum(in_l, 0)] & (in_l >= 0)
        t_in_l = torch.from_numpy(in_l.reshape(-1)).to(dev)
        t_fill = torch.from_numpy(fill).to(dev)
        t_zu = torch.from_numpy(zero_u.reshape(-1).astype(np.uint8)).to(dev) if zero_u.any() else None
        with torch.cuda.device(dev):
            st = self._stream()
            r = torch.zeros((G, N), dtype=torch.float64, device=dev)
            series = _lib.Series(_ptr(truth_dev), truth_dev.stride(0), _ptr(t_in_l), _ptr(t_fill), _ptr(t_zu))
            nb = max(lib.xesn_forced_scratch_bytes(res.handle, G, n_spinup, 1),
                     lib.xesn_predict_scratch_bytes(res.handle, O, G, n_cells))
            scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
            _lib.check(lib.xesn_forced_run(res.handle, ctypes.byref(series), 0, n_spinup, G, _ptr(r), None,
                                           _ptr(scratch), nb, st))

            # ---- closed loop on the global cell vector
            field = self._initial_field(y, n_spinup, dev)
            t_in = torch.from_numpy(in_map[g0:g1].reshape(-1)).to(dev)
            t_out = torch.from_numpy(out_map[g0:g1].reshape(-1)).to(dev)
            pred = torch.empty((n_cells, n_steps + 1), dtype=torch.float64, device=dev)
            if world == 1:
                _lib.check(lib.xesn_predict(res.handle, _ptr(wout), O, G, n_cells, _ptr(t_in), _ptr(t_fill),
                                            _ptr(t_out), _ptr(field), _ptr(r), n_steps, _ptr(pred), _ptr(scratch),
                                            nb, st))
            else:
                # latency-bound shapes: one persistent kernel per rank with the exchange inside (P2P);
                # bandwidth-bound shapes (huge Win/Wout traffic per step): per-step launches + NCCL
                # engines: "flow" (default) = dataflow kernel, neighbour-only P2P stores of the halo cells;
                # "p2p" = persistent kernel with grid barriers, every cell to every rank; "nccl" = per-step launches
                # and an NCCL exchange (also the fallback for shapes whose Win/Wout traffic per step is HBM-sized)
                engine = os.environ.get("XESN_B200_EXCHANGE", "flow")
                small = G * N * I * 8 <= (64 << 20) and world <= 8
                team, ring_steps = ctypes.c_int32(0), ctypes.c_int32(0)
                if engine == "flow" and lib.xesn_flow_geometry(res.handle, O, G, ctypes.byref(team),
                                                               ctypes.byref(ring_steps)) != 0:
                    engine = "p2p"
                if engine == "flow":
                    self._predict_multi_flow(lib, res, wout, G, n_cells, in_map, out_map, t_in, t_fill, t_out, field, r,
                                             n_steps, pred, scratch, nb, st, team.value, ring_steps.value)
                else:
                    multi = self._predict_multi if (small and engine == "p2p") else self._predict_multi_nccl
                    multi(lib, res, wout, G, g0, n_cells, t_in, t_fill, t_out, out_map, field, r,
                          n_steps, pred, scratch, nb, st)
        return pred

    def _initial_field(self, y, n_spinup, dev):
        torch = _torch()
        if isinstance(y, torch.Tensor):
            return y[..., n_spinup].reshape(-1).to(device=dev, dtype=torch.float64).contiguous().clone()
        data = getattr(y, "data", y) if getattr(y, "dims", None) is not None else y
        col = data[..., n_spinup]
        if hasattr(col, "compute"):
            col = col.compute()
        return torch.from_numpy(np.ascontiguousarray(np.asarray(col, dtype=np.float64).reshape(-1))).to(dev)

    def _predict_multi(self, lib, res, wout, G, g0, n_cells, t_in, t_fill, t_out, out_map, field, r, n_steps,
                       pred, scratch, nb, st):
        """One process per GPU.  Every rank runs ONE persistent forecast kernel over its slab of groups;
        the readout stores each predicted cell into the field buffer of every rank through
        peer-mapped memory (NVLink P2P stores) and the ranks synchronise once per step on peer-mapped
        flags -- the per-step neighbour exchange dask's `overlap` performs in the reference
        (lazyesn.py:235), without leaving the kernel.  The trajectory of the owned cells is gathered
        once at the end (NCCL)."""
        torch = _torch()
        import torch.distributed as dist
        rank, world = _dist()
        O = self.n_output
        dev = res.device
        p2p = self._p2p_buffers(lib, n_cells, world, rank)
        _lib.check(lib.xesn_memcpy_async(p2p["field"][0][rank], _ptr(field), 8 * n_cells, st))
        field_ptrs = (ctypes.c_uint64 * (2 * world))(*[p2p["field"][par][k] for par in range(2) for k in range(world)])
        flag_ptrs = (ctypes.c_uint64 * world)(*p2p["flags"])
        per = out_map.shape[0] // world * out_map.shape[1]
        owned_np = out_map.reshape(world, -1)
        mine = torch.from_numpy(owned_np[rank].astype(np.int64)).to(dev)
        pred[mine, 0] = field[mine]
        _lib.check(lib.xesn_predict_p2p(res.handle, _ptr(wout), O, G, n_cells, _ptr(t_in), _ptr(t_fill), _ptr(t_out),
                                        _ptr(r), n_steps, _ptr(pred), world, rank, p2p["steps"],
                                        ctypes.cast(field_ptrs, ctypes.c_void_p), ctypes.cast(flag_ptrs, ctypes.c_void_p),
                                        _ptr(scratch), nb, st))
        p2p["steps"] += n_steps
        # assemble the trajectory: every rank contributes the rows (cells) it owns
        contiguous = all(np.array_equal(np.sort(owned_np[k]), np.arange(k * per, (k + 1) * per)) for k in range(world))
        if contiguous:
            lo, hi = rank * per, (rank + 1) * per
            dist.all_gather_into_tensor(pred, pred[lo:hi])
        else:
            owned = torch.from_numpy(owned_np.astype(np.int64)).to(dev)
            rows = pred[mine].contiguous()
            gathered = torch.empty((world,) + tuple(rows.shape), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gathered, rows)
            pred[owned.reshape(-1)] = gathered.reshape(-1, rows.shape[-1])

    def _predict_multi_flow(self, lib, res, wout, G, n_cells, in_map, out_map, t_in, t_fill, t_out, field, r, n_steps,
                            pred, scratch, nb, st, team, ring_steps):
        """One process per GPU, dataflow forecast (csrc/predict_flow.cu): every rank runs its slab of groups; per step a
        rank pushes exactly the cells that groups of OTHER ranks read (the `overlap`-deep halo strips) into the rings
        of exactly those ranks with NVLink P2P stores -- the neighbour exchange of lazyesn.py:235 without a fence, a
        flag or a collective.  The trajectory rows are gathered once at the end."""
        torch = _torch()
        rank, world = _dist()
        O = self.n_output
        dev = res.device
        plan = self._flow_plan(lib, in_map, out_map, n_cells, world, rank, G, team, ring_steps, dev, st)
        _lib.check(lib.xesn_predict_flow(res.handle, ctypes.byref(plan["c"]), _ptr(wout), O, G, n_cells, _ptr(t_in),
                                         _ptr(t_fill), _ptr(t_out), _ptr(field), _ptr(r), n_steps, _ptr(pred),
                                         _ptr(scratch), nb, st))
        self._gather_owned_rows(pred, out_map, world, rank, dev)

    def _gather_owned_rows(self, pred, out_map, world, rank, dev):
        """Assemble the trajectory on every rank: each rank contributes the rows (cells) its groups predict."""
        torch = _torch()
        import torch.distributed as dist
        per = out_map.shape[0] // world * out_map.shape[1]
        owned_np = out_map.reshape(world, -1)
        contiguous = all(np.array_equal(np.sort(owned_np[k]), np.arange(k * per, (k + 1) * per)) for k in range(world))
        if contiguous:
            lo, hi = rank * per, (rank + 1) * per
            dist.all_gather_into_tensor(pred, pred[lo:hi])
        else:
            owned = torch.from_numpy(owned_np.astype(np.int64)).to(dev)
            rows = pred[owned[rank]].contiguous()
            gathered = torch.empty((world,) + tuple(rows.shape), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(gathered, rows)
            pred[owned.reshape(-1)] = gathered.reshape(-1, rows.shape[-1])

    def _flow_plan(self, lib, in_map, out_map, n_cells, world, rank, G, team, ring_steps, dev, st):
        """Tables + peer-mapped rings of the dataflow forecast, cached per geometry.  Building them is collective
        (CUDA IPC handles go through torch.distributed once)."""
        torch = _torch()
        import torch.distributed as dist
        key = ("flow", n_cells, world, int(in_map.shape[0]), team, ring_steps)
        if self._p2p is not None and self._p2p.get("key") == key:
            return self._p2p
        self._p2p_release()
        tb = _halo.flow_tables(in_map, out_map, n_cells, world, rank)
        O = out_map.shape[1]
        n_imp = tb["n_import_all"]
        dev_t = {k: torch.from_numpy(tb[k]).to(dev) for k in ("src", "import_cell", "exp_ptr", "exp_peer", "exp_slot")}
        ring_bytes = lambda p: int((G * O + int(n_imp[p])) * ring_steps * team * 8)  # noqa: E731
        own, handles = {}, {}
        for par in range(2):
            ptr = ctypes.c_void_p()
            h = (ctypes.c_uint8 * 64)()
            _lib.check(lib.xesn_p2p_alloc(ring_bytes(rank), ctypes.byref(ptr), ctypes.cast(h, ctypes.c_void_p)))
            _lib.check(lib.xesn_memset_async(ptr, 0xFF, ring_bytes(rank), st))
            own[par], handles[par] = ptr.value, bytes(h)
        torch.cuda.synchronize(dev)
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        importers = set(int(p) for p in tb["exp_peer"])
        plan = _lib.FlowPlan()
        plan.n_import = int(n_imp[rank])
        plan.src_dev, plan.import_cell_dev = _ptr(dev_t["src"]), _ptr(dev_t["import_cell"])
        has_exports = len(tb["exp_peer"]) > 0
        plan.exp_ptr_dev = _ptr(dev_t["exp_ptr"]) if has_exports else None
        plan.exp_peer_dev = _ptr(dev_t["exp_peer"]) if has_exports else None
        plan.exp_slot_dev = _ptr(dev_t["exp_slot"]) if has_exports else None
        plan.world, plan.rank, plan.launches_done = world, rank, 0
        opened = []
        for par in range(2):
            for k in range(world):
                if k == rank:
                    plan.ring_ptrs[par * world + k] = own[par]
                elif k in importers:            # only the ranks that read cells of mine are mapped
                    ptr = ctypes.c_void_p()
                    buf = (ctypes.c_uint8 * 64).from_buffer_copy(everyone[k][par])
                    _lib.check(lib.xesn_p2p_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ptr)))
                    plan.ring_ptrs[par * world + k] = ptr.value
                    opened.append(ptr.value)
        dist.barrier()   # every ring is armed ("unset") and mapped before any rank launches
        self._p2p = {"key": key, "c": plan, "tables": dev_t, "own_flow": list(own.values()), "opened_flow": opened,
                     "importers": sorted(importers), "n_import": int(n_imp[rank])}
        return self._p2p

    def _p2p_buffers(self, lib, n_cells, world, rank):
        """Peer-mapped field buffers (2 parities) and step flags of every rank, opened through CUDA
        IPC handles exchanged with torch.distributed; cached per domain size."""
        import torch.distributed as dist
        if self._p2p is not None and self._p2p.get("key") == ("p2p", n_cells, world):
            return self._p2p
        self._p2p_release()  # other geometry: unmap the peers' buffers and free our own before allocating anew
        mine = {}
        handles = {}
        for name, nbytes in (("f0", 8 * n_cells), ("f1", 8 * n_cells), ("flags", 256)):
            ptr = ctypes.c_void_p()
            h = (ctypes.c_uint8 * 64)()
            _lib.check(lib.xesn_p2p_alloc(nbytes, ctypes.byref(ptr), ctypes.cast(h, ctypes.c_void_p)))
            mine[name] = ptr.value
            handles[name] = bytes(h)
        everyone = [None] * world
        dist.all_gather_object(everyone, handles)
        field = [[0] * world, [0] * world]
        flags = [0] * world
        for k in range(world):
            if k == rank:
                field[0][k], field[1][k], flags[k] = mine["f0"], mine["f1"], mine["flags"]
                continue
            for name, slot in (("f0", (field[0], k)), ("f1", (field[1], k)), ("flags", (flags, k))):
                ptr = ctypes.c_void_p()
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(everyone[k][name])
                _lib.check(lib.xesn_p2p_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(ptr)))
                slot[0][slot[1]] = ptr.value
        dist.barrier()
        self._p2p = {"key": ("p2p", n_cells, world), "n_cells": n_cells, "world": world, "field": field, "flags": flags,
                     "steps": 0, "own": mine}
        return self._p2p

    def _p2p_release(self):
        """Close the peer mappings and free this rank's peer-visible buffers (collective in effect: every rank
        drops its buffers when its geometry changes or the model is rebuilt / destroyed)."""
        p2p, self._p2p = getattr(self, "_p2p", None), None
        if not p2p:
            return
        lib = _lib.load()
        if "own_flow" in p2p:
            for ptr in p2p["opened_flow"]:
                lib.xesn_p2p_close(ptr)
            for ptr in p2p["own_flow"]:
                lib.xesn_p2p_free(ptr)
            return
        own = set(p2p["own"].values())
        for ptr in set(p2p["field"][0]) | set(p2p["field"][1]) | set(p2p["flags"]):
            if ptr and ptr not in own:
                lib.xesn_p2p_close(ptr)
        for ptr in own:
            lib.xesn_p2p_free(ptr)

    def _predict_multi_nccl(self, lib, res, wout, G, g0, n_cells, t_in, t_fill, t_out, out_map, field, r, n_steps,
                       pred, scratch, nb, st):
        """Per-step launches + NCCL (used for bandwidth-bound shapes).  Every rank advances its slab of groups, then the ranks
        all-gather the cells they own (NCCL over NVLink) -- the per-step neighbour exchange
        that dask's `overlap` performs in the reference (lazyesn.py:235).

        When every rank's cells form one contiguous range of the flat domain (slabs along the
        first axis, the normal case) the all-gather runs in place on the field vector and the
        readout kernel writes the trajectory of the owned cells directly; the full trajectory is
        assembled with one all-gather at the end.  Otherwise an index-based exchange is used."""
        torch = _torch()
        import torch.distributed as dist
        rank, world = _dist()
        O = self.n_output
        dev = res.device
        per = out_map.shape[0] // world * out_map.shape[1]
        owned_np = out_map.reshape(world, -1)
        contiguous = all(np.array_equal(np.sort(owned_np[k]), np.arange(k * per, (k + 1) * per)) for k in range(world))
        nxt = field.clone()
        ld = pred.stride(0)
        if contiguous:
            lo, hi = rank * per, (rank + 1) * per
            pred[lo:hi, 0] = field[lo:hi]
            for n in range(1, n_steps + 1):
                _lib.check(lib.xesn_predict_step(res.handle, _ptr(wout), O, G, _ptr(t_in), _ptr(t_fill), _ptr(t_out),
                                                 _ptr(field), _ptr(nxt), _ptr(r), pred.data_ptr() + 8 * n, ld,
                                                 _ptr(scratch), nb, st))
                dist.all_gather_into_tensor(nxt, nxt[lo:hi])
                field, nxt = nxt, field
            dist.all_gather_into_tensor(pred, pred[lo:hi])
            return
        owned = torch.from_numpy(owned_np.astype(np.int64)).to(dev)
        pred[:, 0] = field
        gathered = torch.empty(tuple(owned.shape), dtype=torch.float64, device=dev)
        for n in range(1, n_steps + 1):
            _lib.check(lib.xesn_predict_step(res.handle, _ptr(wout), O, G, _ptr(t_in), _ptr(t_fill), _ptr(t_out),
                                             _ptr(field), _ptr(nxt), _ptr(r), None, 0, _ptr(scratch), nb, st))
            exchange_owned(nxt, owned, rank, gathered)
            pred[:, n] = nxt
            field, nxt = nxt, field
