"""Checkpoint round trip: `from_zarr` of the reference (xesn/io.py:13-53) plus a dependency-free twin.

A trained model is stored as its readout matrix and its constructor kwargs -- the seeds of W, Win and the bias vector
ARE the checkpoint of those matrices (`ESN.to_xds`, esn.py:309-338): loading re-runs `build()`.  `from_zarr` needs
xarray + zarr exactly like the reference and raises ImportError where they are absent (this image); `to_npz` /
`from_npz` store the same content (Wout + JSON of the kwargs) in one .npz file, so a model trained here can be handed
to a process without xarray, and the attribute handling below is exercised by the CPU tests either way.
LazyESN: `Wout` is stored in the block layout dask assembles in the reference ((O, G N) / (Gx O, Gy N) /
(Gx O, Gy N, Gz), lazyesn.py:72-77), squeezed; singleton trailing axes are re-appended on load (io.py:46-49).
"""
import inspect
import json
import warnings

import numpy as np

from .esn import ESN, xr
from .lazyesn import LazyESN


def _model_from_attrs(attrs):
    """ESN / LazyESN from stored constructor kwargs (io.py:30-40: lazy iff the attrs hold "overlap"); built."""
    is_lazy = "overlap" in attrs
    Model = LazyESN if is_lazy else ESN
    names = [k for k in inspect.signature(Model.__init__).parameters if k not in ("self", "device")]
    args = {}
    for key in names:
        val = attrs[key]
        args[key] = dict(val) if isinstance(val, dict) else val
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = Model(**args)
    esn.build()
    return esn, is_lazy


def _set_wout(esn, is_lazy, wout):
    wout = np.asarray(wout, dtype=np.float64)
    if is_lazy:
        for _ in range(esn.ndim_state - 2):        # re-append the singleton axes `squeeze` removed (io.py:46-49)
            if wout.ndim < esn.ndim_state:
                wout = wout[..., None]
        grid = tuple(wout.shape[k] // n for k, n in enumerate((esn.n_output, esn.n_reservoir))) + tuple(wout.shape[2:])
        esn._grid = grid
    esn.Wout = wout
    return esn


def from_zarr(store, **kwargs):
    """Create an ESN or LazyESN from a zarr store written by `esn.to_xds().to_zarr(store)` -- by this package or by
    the reference (same attrs, same "Wout" variable).  Needs xarray and zarr."""
    if xr is None:
        raise ImportError("xesn_b200.from_zarr needs xarray and zarr, which are not installed; use from_npz/to_npz")
    xds = xr.open_zarr(store, **kwargs)
    esn, is_lazy = _model_from_attrs(xds.attrs)
    if "Wout" in xds:
        _set_wout(esn, is_lazy, xds["Wout"].values)
    else:
        warnings.warn("from_zarr: did not find 'Wout' in zarr store, returning an untrained network.")
    return esn


def _jsonable(v):
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    return v


def to_npz(esn, path):
    """Readout (squeezed, block layout for LazyESN) + constructor kwargs as JSON: the content of `to_xds()`."""
    if esn.Wout is None:
        raise Exception("ESN.to_xds: Wout has not been computed yet, so it's not worth storing this model")
    attrs = _jsonable(esn._get_attrs())
    np.savez_compressed(path, Wout=np.asarray(esn.Wout).squeeze(), attrs=np.asarray(json.dumps(attrs)),
                        esn_type=np.asarray(type(esn).__name__))


def from_npz(path):
    z = np.load(path, allow_pickle=False)
    attrs = json.loads(str(z["attrs"]))
    esn, is_lazy = _model_from_attrs(attrs)
    return _set_wout(esn, is_lazy, z["Wout"])
