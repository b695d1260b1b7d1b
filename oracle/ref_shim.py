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
