"""`xesn_b200.psd` -- the 1-D power spectral density of a time-varying field on the device.

Drop-in for `xesn.psd.psd` (reference xesn/psd.py:12-70): FFT over the first spatial axis (one spatial axis) or a
2-D FFT over the first two (two or three spatial axes, both transformed axes equally long, psd.py:42-49), |.|^2,
mean over a third spatial axis (psd.py:59), mean over the spectral cells of every unit-width wavenumber annulus
(`scipy.stats.binned_statistic` with bins .5, 1.5, ..., n_x//2 + .5) times the annulus factor pi (k2^2 - k1^2).
The reference does the binning on the CPU even under cupy ("the final binning does not run on the GPU",
psd.py:18-19); here the transform is cuFFT (through torch.fft) and the binning a CUDA kernel of
libxesn_b200.so (`xesn_psd_radial`, csrc/cost.cu), so the spectra of a batch of forecasts never leave the device.
"""
import numpy as np

from . import _lib

_TABLES = {}


def _torch():
    import torch
    return torch


def bin_tables(n_x, ndim_space):
    """(bin_ptr, pix_idx, factor, k_vals) of the wavenumber annuli of an n_x (x n_x) grid, as the reference bins them
    (psd.py:42-56): k = n_x * fftfreq(n_x); |k| per spectral cell; bins [b + .5, b + 1.5), the last one closed."""
    key = (int(n_x), 1 if ndim_space == 1 else 2)
    if key not in _TABLES:
        k = np.fft.fftfreq(n_x) * n_x
        if ndim_space == 1:
            k_tot = k            # signed, as psd.py:45-46 has it: only the positive half of the spectrum is binned
        else:
            kx, ky = np.meshgrid(k, k)
            k_tot = np.sqrt(kx ** 2 + ky ** 2).reshape(-1)
        n_k = n_x // 2 + 1
        k_bins = np.arange(.5, n_k, 1.)
        n_bins = len(k_bins) - 1
        which = np.searchsorted(k_bins, k_tot, side="right") - 1
        which[k_tot == k_bins[-1]] = n_bins - 1                 # the last bin includes its right edge
        which[(k_tot < k_bins[0]) | (k_tot > k_bins[-1])] = -1
        order = np.argsort(which, kind="stable")
        order = order[which[order] >= 0]
        counts = np.bincount(which[which >= 0], minlength=n_bins)
        bin_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        factor = np.pi * (k_bins[1:] ** 2 - k_bins[:-1] ** 2)
        _TABLES[key] = (bin_ptr, order.astype(np.int32), factor.astype(np.float64), 0.5 * (k_bins[1:] + k_bins[:-1]))
    return _TABLES[key]


def psd_device(x, n_space):
    """x: CUDA tensor (..., space axes ..., n_time) float64 with `n_space` (1, 2 or 3) spatial axes before the time axis;
    returns (..., n_x // 2, n_time).  Every leading axis is a batch axis."""
    torch = _torch()
    if n_space not in (1, 2, 3):
        raise NotImplementedError("psd will only work for up to 3D time varying arrays (i.e., 4 dims total)")
    lib = _lib.load()
    sp = x.shape[-1 - n_space:-1]
    n_x, S = int(sp[0]), int(x.shape[-1])
    if n_space >= 2 and int(sp[1]) != n_x:
        raise ValueError("psd: the two transformed spatial axes must be equally long (the reference builds both "
                         "wavenumber axes from the first one, xesn/psd.py:42-49)")
    batch = tuple(x.shape[:-1 - n_space])
    B = int(np.prod(batch)) if batch else 1
    n_z = int(sp[2]) if n_space == 3 else 1
    xb = x.reshape((B,) + tuple(sp) + (S,)).to(torch.float64)
    # image = one time of one field: (B, S, [z,] x [, y])
    if n_space == 1:
        spec = torch.fft.fft(xb.permute(0, 2, 1), dim=-1)                         # (B, S, x)
    elif n_space == 2:
        spec = torch.fft.fft2(xb.permute(0, 3, 1, 2), dim=(-2, -1))                # (B, S, x, y)
    else:
        spec = torch.fft.fft2(xb.permute(0, 4, 3, 1, 2), dim=(-2, -1))             # (B, S, z, x, y)
    spec = torch.view_as_real(spec.contiguous())                                   # interleaved (re, im)
    bin_ptr, pix_idx, factor, _ = bin_tables(n_x, n_space)
    n_bins, n_pix = len(factor), n_x if n_space == 1 else n_x * n_x
    dev = x.device
    t_ptr, t_idx, t_fac = (torch.from_numpy(a).to(dev) for a in (bin_ptr, pix_idx, factor))
    out = torch.empty((B * S, n_bins), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.xesn_psd_radial(spec.data_ptr(), B * S, n_z, n_pix, t_ptr.data_ptr(), t_idx.data_ptr(),
                                       t_fac.data_ptr(), n_bins, out.data_ptr(), st))
    return out.reshape(batch + (S, n_bins)).transpose(-1, -2).contiguous()


def psd(xda, device=None):
    """`xesn.psd.psd(xda)`: xda is a DataArray-like (`.dims` ending in "time" / "ftime") or a bare array with time last
    and 1-3 spatial axes.  Returns an xarray.DataArray over ("k1d", time) when xarray is importable and the input was
    one, else a numpy array (n_x // 2, n_time)."""
    from .esn import _is_dataarray, xr
    torch = _torch()
    dims = getattr(xda, "dims", None)
    if dims is not None:
        assert dims[-1] in ("time", "ftime"), "psd_1d requires 'time' to be on the last dimension"
    data = getattr(xda, "data", xda) if dims is not None else xda
    if hasattr(data, "compute"):
        data = data.compute()
    ndim = int(getattr(data, "ndim"))
    if not ndim < 5:
        raise NotImplementedError("psd will only work for up to 3D time varying arrays (i.e., 4 dims total)")
    if not torch.cuda.is_available():
        raise _lib.XesnError("xesn_b200.psd needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    x = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float64)))
    out = psd_device(x.to(dev), ndim - 1).cpu().numpy()
    if _is_dataarray(xda):
        tdim = "time" if "time" in xda.dims else "ftime"
        k_vals = bin_tables(int(data.shape[0]), ndim - 1)[3]
        attrs = {"long_name": "psd_of_" + xda.attrs["long_name"]} if "long_name" in xda.attrs else {}
        res = xr.DataArray(out, coords={"k1d": k_vals, "time": xda[tdim]}, dims=("k1d", tdim), attrs=attrs)
        res["k1d"].attrs = {"units": "", "description": "nondimensional 1D wavenumber"}
        return res
    return out


def psd_nrmse_device(pred, truth, space_shape):
    """NRMSE of the spectra of forecasts against the truth's (xesn/cost.py:257-272): pred (F, G, O, S), truth (F, O, S)
    CUDA tensors whose O cells unflatten (C order) to `space_shape`; returns the (F, G) tensor."""
    torch = _torch()
    lib = _lib.load()
    F, G, O, S = (int(n) for n in pred.shape)
    shape = tuple(int(n) for n in space_shape)
    assert int(np.prod(shape)) == O and truth.shape == (F, O, S)
    pp = psd_device(pred.reshape((F, G) + shape + (S,)), len(shape))       # (F, G, K, S)
    pt = psd_device(truth.reshape((F,) + shape + (S,)), len(shape))        # (F, K, S)
    K = int(pt.shape[-2])
    out = torch.empty((F, G), dtype=torch.float64, device=pred.device)
    with torch.cuda.device(pred.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.xesn_nrmse_skipnan(pp.data_ptr(), pt.data_ptr(), F, G, K, S, out.data_ptr(), st))
    return out
