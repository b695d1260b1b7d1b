"""Synthetic driving data of the shapes BASELINE.json names (host side, numpy).

* `lorenz96`: dx_i/dt = (x_{i+1} - x_{i-2}) x_{i-1} - x_i + F, the system the
  reference's examples and scaling runs use (docs/lorenz.py:29-49,
  scaling/create_data.py:11-22).  The reference integrates with scipy `odeint`;
  here a fixed-step RK4 is used -- the data only has to have the right shape and
  statistics, it is not a parity quantity.
* `wave_field_2d`: doubly periodic superposition of travelling waves, used for
  the 2-D LazyESN configuration (no reference dataset exists for it).

Both return float64 arrays with time LAST, already normalised by a scalar
mean/std as scaling/create_data.py:18-22 does.
"""
import numpy as np


def _l96_rhs(x, forcing):
    return (np.roll(x, -1) - np.roll(x, 2)) * np.roll(x, 1) - x + forcing


def lorenz96(n_vars, n_steps, dt=0.01, forcing=8.0, discard=2000, normalise=True):
    x = np.zeros(n_vars)
    x[0] = 0.01
    out = np.empty((n_vars, n_steps))
    for n in range(discard + n_steps):
        k1 = _l96_rhs(x, forcing)
        k2 = _l96_rhs(x + 0.5 * dt * k1, forcing)
        k3 = _l96_rhs(x + 0.5 * dt * k2, forcing)
        k4 = _l96_rhs(x + dt * k3, forcing)
        x = x + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        if n >= discard:
            out[:, n - discard] = x
    if normalise:
        out = (out - out.mean()) / out.std()
    return out


def wave_field_2d(nx, ny, n_steps, n_modes=8, seed=0, normalise=True):
    rs = np.random.RandomState(seed)
    kx = rs.randint(-8, 9, size=n_modes)
    ky = rs.randint(-8, 9, size=n_modes)
    amp = rs.uniform(0.5, 1.5, size=n_modes)
    omega = rs.uniform(0.05, 0.5, size=n_modes)
    phase = rs.uniform(0.0, 2.0 * np.pi, size=n_modes)
    x = np.arange(nx)[:, None, None]
    y = np.arange(ny)[None, :, None]
    t = np.arange(n_steps)[None, None, :]
    out = np.zeros((nx, ny, n_steps))
    for m in range(n_modes):
        out += amp[m] * np.sin(2.0 * np.pi * (kx[m] * x / nx + ky[m] * y / ny) - omega[m] * t + phase[m])
    if normalise:
        out = (out - out.mean()) / out.std()
    return out
