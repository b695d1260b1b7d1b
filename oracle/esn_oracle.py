"""numpy restatement of the xesn ESN / LazyESN hot path -- TEST INFRASTRUCTURE ONLY.

This file is the *checker* for the CUDA path; the product never imports it.
Every function names the reference lines it restates (paths relative to the
upstream tree, v0.2.2).  It is pinned against outputs of the reference's own
source (see `oracle/make_golden.py` and `tests/test_oracle_golden.py`).

All arithmetic is float64, matching the reference's numpy path.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.linalg import solve as _sym_solve


# --------------------------------------------------------------------------- #
# eager ESN
# --------------------------------------------------------------------------- #
def leaky_step(r, u, W, Win, b, alpha):
    """One reservoir update.  Restates xesn/esn.py:458-460 (`_update`).

    r' = alpha * tanh(W r + Win u + b) + (1 - alpha) r
    """
    pre = W @ r + Win @ u + b
    return alpha * np.tanh(pre) + (1.0 - alpha) * r


def forced_states(u, n_steps, W, Win, b, alpha, r0=None):
    """States r_0 .. r_{n_steps} of a teacher-forced run (row t = state *before*
    u[:, t] is consumed).  This is the `rT` buffer of xesn/esn.py:498-507
    un-batched; used for state parity."""
    N = Win.shape[0]
    out = np.zeros((n_steps + 1, N))
    if r0 is not None:
        out[0] = r0
    for t in range(n_steps):
        out[t + 1] = leaky_step(out[t], u[:, t], W, Win, b, alpha)
    return out


def ridge_accumulate(u, y, n_spinup, batch_size, W, Win, b, alpha):
    """Spin-up + batched Gram accumulation.  Restates xesn/esn.py:476-513 (the part of
    `_train_1d` before the solve).  Returns (rbar (N,N), ybar (O,N)) without the Tikhonov term.

    u:(I,T), y:(O,T) time-last.  The state that has seen u[:, :t] is regressed onto y[:, t];
    Gram terms are accumulated batch by batch in the reference's order (batch boundaries
    change the floating-point sum order)."""
    N = Win.shape[0]
    O, T = y.shape
    B = T if batch_size is None else batch_size
    n_batches = int(np.ceil((T - n_spinup) / B))

    states = np.zeros((B + 1, N))
    yr = np.zeros((O, N))
    rr = np.zeros((N, N))

    for t in range(n_spinup):
        states[0] = leaky_step(states[0], u[:, t], W, Win, b, alpha)

    for k in range(n_batches):
        lo = n_spinup + k * B
        hi = min(lo + B, T)
        m = hi - lo
        for j in range(m):
            states[j + 1] = leaky_step(states[j], u[:, lo + j], W, Win, b, alpha)
        yr += y[:, lo:hi] @ states[:m]
        rr += states[:m].T @ states[:m]
        states[0] = states[m].copy()
    return rr, yr


def ridge_solve(rr, yr, beta):
    """Tikhonov term + symmetric solve.  Restates xesn/esn.py:516-520 (modifies rr in place
    like the reference does)."""
    rr += beta * np.eye(rr.shape[0])
    return _sym_solve(rr.T, yr.T, assume_a="sym").T


def ridge_train(u, y, n_spinup, batch_size, W, Win, b, alpha, beta, return_gram=False):
    """Ridge-regression readout.  Restates xesn/esn.py:463-520 (`_train_1d`)."""
    rr, yr = ridge_accumulate(u, y, n_spinup, batch_size, W, Win, b, alpha)
    Wout = ridge_solve(rr, yr, beta)
    if return_gram:
        return Wout, rr, yr
    return Wout


def closed_loop(y, n_steps, n_spinup, W, Win, b, alpha, Wout):
    """Forecast.  Restates the loop of xesn/esn.py:251-268 (`ESN.predict`).

    y:(O,T>=n_spinup+1).  Returns (O, n_steps+1); column 0 is y[:, n_spinup].
    """
    N = Win.shape[0]
    r = np.zeros(N)
    for t in range(n_spinup):
        r = leaky_step(r, y[:, t], W, Win, b, alpha)
    v = np.zeros((n_steps + 1, Wout.shape[0]))
    v[0] = y[:, n_spinup]
    for n in range(1, n_steps + 1):
        r = leaky_step(r, v[n - 1], W, Win, b, alpha)
        v[n] = Wout @ r
    return v.T


# --------------------------------------------------------------------------- #
# LazyESN: halo construction (dask.array.overlap.overlap restated with np.pad)
# --------------------------------------------------------------------------- #
def _axis_boundary(boundary, axis):
    if isinstance(boundary, dict):
        return boundary.get(axis, "none")
    return boundary


def pad_domain(field, depth, boundary):
    """Pad the global domain axis by axis (in axis order, so corners compose the
    way dask's `boundaries()` composes them).  `depth`/`boundary` are keyed by
    axis index as in xesn/lazyesn.py:251-261.  'periodic' -> wrap, 'reflect' ->
    edge-inclusive mirror (numpy 'symmetric'), number -> constant fill."""
    out = field
    for ax in range(field.ndim):
        d = depth.get(ax, 0)
        if d == 0:
            continue
        kind = _axis_boundary(boundary, ax)
        width = [(0, 0)] * out.ndim
        width[ax] = (d, d)
        if kind == "periodic":
            out = np.pad(out, width, mode="wrap")
        elif kind == "reflect":
            out = np.pad(out, width, mode="symmetric")
        elif isinstance(kind, str):
            raise NotImplementedError(f"boundary kind {kind!r}")
        else:
            out = np.pad(out, width, mode="constant", constant_values=kind)
    return out


def group_grid(shape, chunks):
    """Number of groups along every non-time axis."""
    grid = []
    for n, c in zip(shape, chunks):
        if n % c:
            raise ValueError("domain size must be divisible by the chunk size")
        grid.append(n // c)
    return tuple(grid)


def halo_blocks(field, chunks, depth, boundary):
    """Yield (group_index_tuple, halo'd block) for every group, C-order over the
    group grid.  field: (*space, T); chunks: per space axis; depth: {axis: d}.
    Equivalent of `overlap(y.chunk(...), depth, boundary)` at
    xesn/lazyesn.py:165-166 / :199-200 followed by iterating the blocks."""
    nsp = field.ndim - 1
    padded = pad_domain(field, depth, boundary)
    grid = group_grid(field.shape[:nsp], chunks)
    for g in itertools.product(*[range(n) for n in grid]):
        sl = []
        for ax in range(nsp):
            d = depth.get(ax, 0)
            lo = g[ax] * chunks[ax]
            sl.append(slice(lo, lo + chunks[ax] + 2 * d))
        sl.append(slice(None))
        yield g, padded[tuple(sl)]


def inner_mask(block_shape, depth):
    """True on the cells a group predicts (non-halo), broadcast over the
    non-overlapped trailing space axes.  Restates xesn/lazyesn.py:330-349;
    block_shape includes the time axis, result has a trailing length-1 axis."""
    per_axis = []
    for ax, d in depth.items():
        if d > 0:
            m = np.ones(block_shape[ax], dtype=bool)
            m[:d] = False
            m[-d:] = False
            per_axis.append(m)
    m = np.outer(*per_axis) if len(per_axis) > 1 else per_axis[0]
    target = tuple(block_shape[:-1]) + (1,)
    return np.broadcast_to(m.T, target[::-1]).T


def _flat(block):
    return block.reshape(int(np.prod(block.shape[:-1])), block.shape[-1])


def train_block(block, depth, n_spinup, batch_size, W, Win, b, alpha, beta):
    """Per-group training.  Restates xesn/lazyesn.py:264-298 (`_train_nd`)
    without the final reshape: returns (O, N) with NaN rows where the target is
    masked."""
    u = _flat(block)
    y = u[inner_mask(block.shape, depth).flatten(), :]
    O = y.shape[0]
    N = Win.shape[0]
    drop_in = np.isnan(u[:, 0])
    u = u[~drop_in]
    Win_g = Win[:, ~drop_in]
    drop_out = np.isnan(y[:, 0])
    y = y[~drop_out]
    out = np.full((O, N), np.nan)
    out[~drop_out] = ridge_train(u, y, n_spinup, batch_size, W, Win_g, b, alpha, beta)
    return out


def lazy_train(field, chunks, depth, boundary, n_spinup, batch_size, W, Win, b, alpha, beta):
    """All groups.  Restates xesn/lazyesn.py:148-185.  Returns array
    (*group_grid, O, N) -- the per-group readouts before dask lays them out as
    blocks of size `_Wout_chunks`."""
    nsp = field.ndim - 1
    grid = group_grid(field.shape[:nsp], chunks)
    O = int(np.prod(chunks))
    N = Win.shape[0]
    out = np.zeros(grid + (O, N))
    for g, block in halo_blocks(field, chunks, depth, boundary):
        out[g] = train_block(block, depth, n_spinup, batch_size, W, Win, b, alpha, beta)
    return out


def wout_block_layout(wout_groups):
    """(*grid, O, N) -> the 2-D/3-D array dask's map_blocks assembles at
    xesn/lazyesn.py:169-185 with chunks=_Wout_chunks (:72-77) and drop_axis=-1:
    1-D grid (G,)      -> (O, G*N)        [new leading axis, x blocks on axis 1]
    2-D grid (Gx,Gy)   -> (Gx*O, Gy*N)
    3-D grid (Gx,Gy,Gz)-> (Gx*O, Gy*N, Gz)
    (pinned by xesn/test/lazy.py:46-88)."""
    grid = wout_groups.shape[:-2]
    O, N = wout_groups.shape[-2:]
    if len(grid) == 1:
        return np.concatenate(list(wout_groups), axis=1)
    if len(grid) == 2:
        return np.block([[wout_groups[i, j] for j in range(grid[1])] for i in range(grid[0])])
    if len(grid) == 3:
        out = np.zeros((grid[0] * O, grid[1] * N, grid[2]))
        for i, j, k in itertools.product(*[range(n) for n in grid]):
            out[i * O:(i + 1) * O, j * N:(j + 1) * N, k] = wout_groups[i, j, k]
        return out
    raise NotImplementedError


def _masked_inputs(u_flat, Win):
    """NaN-boundary handling.  Restates xesn/lazyesn.py:357-365."""
    drop = np.isnan(u_flat[:, 0])
    return u_flat[~drop], Win[:, ~drop]


def lazy_predict(field, chunks, depth, boundary, n_steps, n_spinup, W, Win, b, alpha, wout_groups):
    """Closed-loop LazyESN forecast.  Restates xesn/lazyesn.py:190-248 with the
    block functions `_spinup` (:301-316), `_update_nd` (:319-322), `_readout`
    (:325-327).  field: (*space, T >= n_spinup+1).  Returns (*space, n_steps+1)."""
    nsp = field.ndim - 1
    space = field.shape[:nsp]
    grid = group_grid(space, chunks)
    N = Win.shape[0]
    truth = field[..., :n_spinup + 1]

    # spin-up on halo'd truth
    states = {}
    for g, block in halo_blocks(truth, chunks, depth, boundary):
        u, Win_g = _masked_inputs(_flat(block), Win)
        r = np.zeros(N)
        for t in range(n_spinup):
            r = leaky_step(r, u[:, t], W, Win_g, b, alpha)
        states[g] = r

    out = np.zeros(space + (n_steps + 1,))
    out[..., 0] = truth[..., n_spinup]
    current = truth[..., n_spinup]
    for n in range(1, n_steps + 1):
        nxt = np.zeros(space)
        for g, block in halo_blocks(current[..., None], chunks, depth, boundary):
            u, Win_g = _masked_inputs(_flat(block), Win)
            r = leaky_step(states[g], u[:, 0], W, Win_g, b, alpha)
            states[g] = r
            v = wout_groups[g] @ r
            sl = tuple(slice(g[ax] * chunks[ax], (g[ax] + 1) * chunks[ax]) for ax in range(nsp))
            nxt[sl] = v.reshape(chunks)
        out[..., n] = nxt
        current = nxt
    return out


# --------------------------------------------------------------------------- #
# comparators used by the parity tests
# --------------------------------------------------------------------------- #
def max_normalised_err(a, ref):
    """max |a - ref| / max |ref| -- the Wout metric of SURVEY.md section 8(c)."""
    return float(np.nanmax(np.abs(a - ref)) / max(np.nanmax(np.abs(ref)), 1e-300))


def agreement_horizon(a, ref, atol):
    """First forecast step (last axis) at which |a-ref| exceeds atol anywhere;
    equals the trajectory length if it never does."""
    bad = np.abs(a - ref).reshape(-1, a.shape[-1]).max(axis=0) > atol
    idx = np.nonzero(bad)[0]
    return int(idx[0]) if idx.size else a.shape[-1]


# --------------------------------------------------------------------------- cost of forecasts
def nrmse_cost(preds, truths, upper_bound=1e9):
    """NRMSE cost of G candidates over F forecast windows.  Restates xesn/cost.py:227-254 (`nrmse`, drop_time=True:
    error normalised by the temporal standard deviation of the truth per variable, ddof 0, mean over variables and
    time, square root) and xesn/cost.py:201-225 (mean over the windows, NaN/inf/too-large -> cost_upper_bound).
    PARITY UNPINNED: the reference's version is written with xarray, which is not installed here, so this
    restatement could not be run against it.

    preds: (F, G, O, S), truths: (F, O, S).  Returns (cost (G,), nrmse (F, G))."""
    preds, truths = np.asarray(preds, dtype=np.float64), np.asarray(truths, dtype=np.float64)
    w = 1.0 / truths.std(axis=-1, keepdims=True)                      # (F, O, 1)
    err = (preds - truths[:, None]) * w[:, None]                       # (F, G, O, S)
    fg = np.sqrt((err ** 2).mean(axis=(2, 3)))
    cost = fg.mean(axis=0)
    bad = np.isnan(cost) | np.isinf(cost) | (cost > upper_bound)
    return np.where(bad, upper_bound, cost), fg


def psd_1d(x):
    """1-D power spectral density of a field with one spatial axis.  Restates xesn/psd.py:12-65 (`psd`) for
    2-D input (n_x, n_time), with the same scipy calls (fft along space, `binned_statistic` mean over unit-width
    wavenumber bins, annulus factor).  Returns (n_x // 2, n_time); for even n_x the last bin is empty (NaN), as in
    the reference.  PARITY UNPINNED (the reference wraps this in xarray, absent here)."""
    from scipy.fft import fft, fftfreq
    from scipy.stats import binned_statistic
    x = np.asarray(x, dtype=np.float64)
    n_x, n_time = x.shape
    psi = np.abs(fft(x.T)) ** 2                       # (n_time, n_x)
    n_k = n_x // 2 + 1
    k = n_x * fftfreq(n_x)
    k_bins = np.arange(.5, n_k, 1.)
    out = np.zeros((n_x // 2, n_time))
    for n in range(n_time):
        tmp1d, *_ = binned_statistic(k.flatten(), psi[n].flatten(), statistic="mean", bins=k_bins)
        out[:, n] = tmp1d * np.pi * (k_bins[1:] ** 2 - k_bins[:-1] ** 2)
    return out


def psd_nd(x):
    """1-D power spectral density of a time-varying field with one, two or three spatial axes (time last).  Restates
    xesn/psd.py:12-70 (`psd`): FFT over the first axis (one spatial axis) or 2-D FFT over the first two (which must be
    equally long, psd.py:42-49 builds both wavenumber axes from n_x), |.|^2, mean over a third spatial axis (psd.py:59),
    mean over the cells of every unit-width wavenumber annulus (`scipy.stats.binned_statistic`, bins .5, 1.5, ...,
    n_x//2 + .5) times the annulus area factor pi (k2^2 - k1^2).  Returns (n_x // 2, n_time); empty bins are NaN.
    PINNED against the reference's own function on seeded inputs (tests/golden/psd.npz, oracle/make_golden.py)."""
    from scipy.fft import fft, fft2, fftfreq
    from scipy.stats import binned_statistic
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        return psd_1d(x)
    if x.ndim not in (3, 4):
        raise NotImplementedError("psd will only work for up to 3D time varying arrays (i.e., 4 dims total)")
    n_x, n_time = x.shape[0], x.shape[-1]
    psi = np.abs(fft2(x.T)) ** 2                       # (n_time, [n_z,] n_y, n_x)
    n_k = n_x // 2 + 1
    k = n_x * fftfreq(n_x)
    k2d = np.meshgrid(k, k)
    k_tot = np.sqrt(k2d[0] ** 2 + k2d[1] ** 2)
    k_bins = np.arange(.5, n_k, 1.)
    out = np.zeros((n_x // 2, n_time))
    for n in range(n_time):
        arr = psi[n].mean(axis=0) if psi.ndim > 3 else psi[n]
        tmp1d, *_ = binned_statistic(k_tot.flatten(), arr.flatten(), statistic="mean", bins=k_bins)
        out[:, n] = tmp1d * np.pi * (k_bins[1:] ** 2 - k_bins[:-1] ** 2)
    return out


def psd_nrmse_fg(preds, truths, space_shape=None):
    """NRMSE of the PSD per window and candidate.  Restates xesn/cost.py:257-272 (`psd_nrmse` = `nrmse` of the two
    spectra, cost.py:227-254) with xarray's skip-NaN reductions.  preds: (F, G, O, S), truths: (F, O, S) -> (F, G);
    `space_shape`: the spatial shape the O cells unflatten to (C order), default one spatial axis.
    The spectra are pinned (psd_nd); the NRMSE reductions on top are PARITY UNPINNED (xarray absent here)."""
    preds, truths = np.asarray(preds, dtype=np.float64), np.asarray(truths, dtype=np.float64)
    F, G = preds.shape[:2]
    S = preds.shape[-1]
    shape = tuple(space_shape) if space_shape is not None else (preds.shape[2],)
    out = np.zeros((F, G))
    import warnings
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)   # the empty last bin of an even n_x is all-NaN by design
        for f in range(F):
            pt = psd_nd(truths[f].reshape(shape + (S,)))
            w = 1.0 / np.nanstd(pt, axis=-1, keepdims=True)
            for g in range(G):
                err = (psd_nd(preds[f, g].reshape(shape + (S,))) - pt) * w
                out[f, g] = np.sqrt(np.nanmean(err ** 2))
    return out
