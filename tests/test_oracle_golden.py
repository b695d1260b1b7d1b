"""Pin the CPU oracle (oracle/esn_oracle.py) against outputs of the reference's
own source files, committed under tests/golden/ by oracle/make_golden.py."""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import esn_oracle as oc
from tests.conftest import LAZY_CASES, decode_boundary, load_golden


def test_forced_states_match_reference_update():
    z = load_golden("eager_small.npz")
    for n_in in (3, 7):
        W = sp.csr_matrix(z[f"W_{n_in}"])
        states = oc.forced_states(z[f"u_{n_in}"], 60, W, z[f"Win_{n_in}"], z[f"b_{n_in}"], 0.5)
        np.testing.assert_allclose(states, z[f"states_{n_in}"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("n_in", (3, 7))
@pytest.mark.parametrize("n_spinup,batch", list(itertools.product((0, 10), (None, 33, 10_000))))
def test_ridge_train_matches_reference(n_in, n_spinup, batch):
    z = load_golden("eager_small.npz")
    W = sp.csr_matrix(z[f"W_{n_in}"])
    u = z[f"u_{n_in}"]
    Wout = oc.ridge_train(u, u, n_spinup, batch, W, z[f"Win_{n_in}"], z[f"b_{n_in}"], 0.5, 1e-6)
    ref = z[f"Wout_{n_in}_{n_spinup}_{batch}"]
    assert Wout.shape == (n_in, 120)
    assert oc.max_normalised_err(Wout, ref) < 1e-9


def test_separate_targets_and_all_spinup():
    z = load_golden("eager_small.npz")
    # the y != u cases were generated with a fresh 7->3 reservoir of the same seeds:
    # W and b equal the 7-input ones, Win too (same seed, same shape)
    W = sp.csr_matrix(z["W_7"])
    args = (W, z["Win_7"], z["b_7"], 0.5, 1e-6)
    Wout = oc.ridge_train(z["u_7"], z["y_sep"], 5, 100, *args)
    assert oc.max_normalised_err(Wout, z["Wout_sep"]) < 1e-9
    Wz = oc.ridge_train(z["u_7"], z["y_sep"], 500, None, *args)
    assert np.array_equal(Wz, z["Wout_allspin"])
    assert not Wz.any()


@pytest.mark.parametrize("n_in", (3, 7))
@pytest.mark.parametrize("n_spinup", (0, 10))
def test_closed_loop_matches_reference(n_in, n_spinup):
    z = load_golden("eager_small.npz")
    W = sp.csr_matrix(z[f"W_{n_in}"])
    pred = oc.closed_loop(z[f"u_{n_in}"], 25, n_spinup, W, z[f"Win_{n_in}"], z[f"b_{n_in}"], 0.5,
                          z[f"Wout_{n_in}_10_33"])
    ref = z[f"pred_{n_in}_{n_spinup}"]
    assert pred.shape == (n_in, 26)
    assert np.array_equal(pred[:, 0], z[f"u_{n_in}"][:, n_spinup])
    np.testing.assert_allclose(pred, ref, rtol=0, atol=1e-12)


def test_l96_case():
    z = load_golden("eager_l96.npz")
    W = sp.csr_matrix(z["W"])
    T = int(z["n_train"])
    data = z["data"]
    Wout = oc.ridge_train(data[:, :T], data[:, :T], 0, int(z["batch"]), W, z["Win"], z["b"], 0.5, 1e-6)
    assert oc.max_normalised_err(Wout, z["Wout"]) < 1e-7
    pred = oc.closed_loop(data[:, int(z["test_offset"]):], int(z["n_steps"]), int(z["n_spinup"]),
                          W, z["Win"], z["b"], 0.5, z["Wout"])
    np.testing.assert_allclose(pred, z["pred"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", LAZY_CASES)
def test_lazy_train_and_predict(name):
    z = load_golden(f"lazy_{name}.npz")
    field = z["field"]
    nsp = field.ndim - 1
    chunks = tuple(int(c) for c in z["chunks"])
    depth = {a: int(z["overlap"][a]) for a in range(nsp)}
    depth[nsp] = 0
    boundary = decode_boundary(z)
    W = sp.csr_matrix(z["W"])
    common = (W, z["Win"], z["b"], float(z["leak"]))

    blk = next(oc.halo_blocks(field, chunks, depth, boundary))[1]
    assert np.array_equal(oc.inner_mask(blk.shape, depth), z["inner_mask"])

    wout = oc.lazy_train(field, chunks, depth, boundary, int(z["n_spinup_train"]), int(z["batch"]),
                         *common, float(z["beta"]))
    ref = z["wout_groups"]
    assert wout.shape == ref.shape
    assert np.array_equal(np.isnan(wout), np.isnan(ref))
    assert oc.max_normalised_err(wout, ref) < 1e-9

    pred = oc.lazy_predict(field, chunks, depth, boundary, int(z["n_steps"]), int(z["n_spinup"]),
                           *common, ref)
    assert pred.shape == field.shape[:-1] + (int(z["n_steps"]) + 1,)
    assert np.array_equal(np.isnan(pred), np.isnan(z["pred"]))
    np.testing.assert_allclose(pred, z["pred"], rtol=0, atol=1e-11)


def test_wout_block_layout_shapes():
    # shapes pinned by xesn/test/lazy.py:46-88 (4 / 4x6 / 4x6x5 domains, chunks 2 / 2x3 / 2x3x5)
    N = 7
    assert oc.wout_block_layout(np.zeros((2, 2, N))).shape == (2, 2 * N)
    assert oc.wout_block_layout(np.zeros((2, 2, 6, N))).shape == (2 * 6, 2 * N)
    assert oc.wout_block_layout(np.zeros((2, 2, 1, 30, N))).shape == (2 * 6 * 5, 2 * N, 1)
    g = np.arange(2 * 3 * 4 * 5, dtype=float).reshape(2, 3, 4, 5)
    lay = oc.wout_block_layout(g)
    assert np.array_equal(lay[4:8, 10:15], g[1, 2])


# --------------------------------------------------------------------------- cost terms (unpinned restatements): known answers
def test_cost_oracle_known_answers():
    """`nrmse_cost`, `psd_1d`, `psd_nrmse_fg` restate xesn/cost.py:201-272 and xesn/psd.py:12-65, which cannot run here
    (xarray); these are closed-form checks of the restatements themselves."""
    from oracle import esn_oracle as oc
    rs = np.random.RandomState(1)
    truth = rs.normal(size=(2, 6, 50))                                  # (F, O, S)
    # a forecast that is off by exactly one temporal standard deviation everywhere has NRMSE 1; a perfect one 0
    off = truth + truth.std(axis=-1, keepdims=True)
    preds = np.stack([truth, off], axis=1)                              # (F, G=2, O, S)
    cost, fg = oc.nrmse_cost(preds, truth, upper_bound=10.0)
    np.testing.assert_allclose(fg, [[0.0, 1.0], [0.0, 1.0]], atol=1e-12)
    np.testing.assert_allclose(cost, [0.0, 1.0], atol=1e-12)
    # NaN and too-large costs are clamped (cost.py:218-223)
    bad = preds.copy()
    bad[0, 0, 0, 0] = np.nan
    bad[:, 1] *= 1e6
    cost, _ = oc.nrmse_cost(bad, truth, upper_bound=10.0)
    np.testing.assert_array_equal(cost, [10.0, 10.0])
    # a pure cosine of wavenumber 3 on 16 points: all power in bin k=3, |X_3|^2 = (16/2)^2, annulus factor 2 pi k
    x = np.cos(2 * np.pi * 3 * np.arange(16) / 16)[:, None] * np.ones((1, 4))
    p = oc.psd_1d(x)
    assert p.shape == (8, 4) and np.isnan(p[7]).all()                   # even n_x: the reference's last bin is empty
    np.testing.assert_allclose(p[2], 64.0 * 2 * np.pi * 3, rtol=1e-12)
    assert np.abs(np.delete(p[:7], 2, axis=0)).max() < 1e-20 * 1e3 + 1e-9
    # the PSD error of a perfect forecast is 0
    tr = rs.normal(size=(1, 16, 30))
    np.testing.assert_allclose(oc.psd_nrmse_fg(tr[:, None], tr), [[0.0]], atol=1e-12)


@pytest.mark.parametrize("name", ["x12", "x13", "xy8", "xy9", "xyz8"])
def test_psd_matches_reference(name):
    """oracle.psd_nd against the reference's own `psd` (xesn/psd.py:12-70, executed unmodified by oracle/make_golden.py):
    one, two and three spatial axes; the empty last bin of an even 1-D size is NaN on both sides."""
    import warnings
    z = load_golden("psd.npz")
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        got = oc.psd_nd(z[f"in_{name}"])
    ref = z[f"psd_{name}"]
    assert got.shape == ref.shape == (z[f"in_{name}"].shape[0] // 2, z[f"in_{name}"].shape[-1])
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(ref), rtol=1e-13)
