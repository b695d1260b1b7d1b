"""Import the reference's own hot-path modules in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where `/root/reference` exists (it
does not exist on the GPU box).  `import xesn` itself fails here because
xarray/dask/smt are not installed, so stub modules are registered first and the
three source files of the hot path are imported *unmodified* from where they
lie.  Used by `oracle/make_golden.py` to produce the committed fixtures.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("XESN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xesn"))


def load():
    """Returns (matrix, esn, lazyesn) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "xesn.lazyesn" not in sys.modules:
        for name in ("xarray", "dask", "dask.array", "dask.array.overlap"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        da = sys.modules["dask.array"]
        for attr in ("map_blocks", "stack"):
            if not hasattr(da, attr):
                setattr(da, attr, None)
        if not hasattr(da, "Array"):
            # scipy's array-api probe looks this attribute up
            da.Array = type("Array", (), {})
        sys.modules["dask"].array = da
        if not hasattr(sys.modules["dask.array.overlap"], "overlap"):
            sys.modules["dask.array.overlap"].overlap = None
        pkg = types.ModuleType("xesn")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "xesn")]
        pkg._use_cupy = False
        sys.modules["xesn"] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        mods = tuple(importlib.import_module(f"xesn.{m}") for m in ("matrix", "esn", "lazyesn"))
    return mods


class _StubDataArray:
    """Just enough of xarray.DataArray for the reference's `psd` (xesn/psd.py:12-94) to run unmodified: it reads
    `.dims`, `.ndim`, `.shape`, `.load()`, `.data`, `.attrs`, indexes the time coordinate by name and packs its result
    with `xr.DataArray(data, coords=..., dims=..., attrs=...)`.  None of the numerics goes through this class."""

    def __init__(self, data=None, coords=None, dims=None, attrs=None):
        import numpy as np
        self.data = np.asarray(data)
        self.dims = tuple(dims) if dims is not None else tuple(f"dim_{k}" for k in range(self.data.ndim))
        self.coords = dict(coords or {})
        self.attrs = dict(attrs or {})

    ndim = property(lambda self: self.data.ndim)
    shape = property(lambda self: self.data.shape)

    def load(self):
        return self

    def __getitem__(self, key):
        if isinstance(key, str):
            import numpy as np
            if key not in self.coords:
                self.coords[key] = _StubDataArray(np.arange(self.data.shape[self.dims.index(key)]), dims=(key,))
            if not isinstance(self.coords[key], _StubDataArray):
                self.coords[key] = _StubDataArray(self.coords[key], dims=(key,))
            return self.coords[key]
        raise KeyError(key)

    def __setitem__(self, key, value):
        self.coords[key] = value


def load_psd():
    """The reference's `psd` function (xesn/psd.py), imported unmodified over a stub `xarray.DataArray`.
    Usage: psd(ref_shim.field_array(x, dims)).data  ->  (n_x // 2, n_time)."""
    load()
    xr = sys.modules["xarray"]
    if not hasattr(xr, "DataArray"):
        xr.DataArray = _StubDataArray
    return importlib.import_module("xesn.psd").psd


def field_array(x, dims):
    return _StubDataArray(x, dims=dims)
