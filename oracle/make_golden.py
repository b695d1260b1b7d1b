"""Generate `tests/golden/*.npz` from the REFERENCE's own source files.

TEST INFRASTRUCTURE ONLY.  Run once in the build container (where
`/root/reference` exists):

    python oracle/make_golden.py

Each fixture stores the seeded inputs together with what the reference's
`xesn/matrix.py`, `xesn/esn.py` and `xesn/lazyesn.py` produced for them with this
container's numpy/scipy.  The eager `predict` loop and the dask halo are the two
pieces that cannot run here (xarray/dask are absent); they are driven below by
12-line loops around the reference's own `_update`, `_spinup`, `_update_nd`,
`_readout` and `_train_nd` block functions, with `np.pad` standing in for dask's
`overlap` (wrap / symmetric / constant, axis by axis).
"""
import itertools
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from xesn_b200.synthetic import lorenz96  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
rmat, resn, rlazy = ref_shim.load()


def dense(W):
    return W.toarray() if hasattr(W, "toarray") else np.asarray(W)


# --------------------------------------------------------------------------- #
def matrices():
    out = {}
    # docstring known answers (xesn/matrix.py:33-72, :204-214)
    m = rmat.RandomMatrix(2, 2, factor=5, distribution="uniform", random_seed=0)
    out["kat_uniform_first"] = m()
    out["kat_uniform_second"] = m()
    out["kat_gauss_eig"] = rmat.RandomMatrix(2, 2, factor=0.9, distribution="gaussian",
                                             normalization="eig", random_seed=0)()
    out["kat_gauss_svd"] = rmat.RandomMatrix(2, 2, factor=1.0, distribution="gaussian",
                                             normalization="svd", random_seed=0)()
    s = rmat.SparseRandomMatrix(100, 100, factor=0.9, distribution="normal", normalization="eig",
                                connectedness=1, random_seed=0)
    out["kat_sparse_eig"] = dense(s())

    idx = 0
    cases = []
    for dist, norm in itertools.product(("uniform", "normal"), ("multiply", "svd", "eig")):
        mk = rmat.RandomMatrix(30, 30, factor=0.7, distribution=dist, normalization=norm, random_seed=11 + idx)
        out[f"dense_{idx}"] = mk()
        cases.append(("dense", 30, 30, 0.7, dist, norm, 11 + idx, "", 0.0))
        idx += 1
    for dist, norm, fmt in itertools.product(("uniform", "normal"), ("multiply", "svd", "eig"), ("csr", "coo")):
        mk = rmat.SparseRandomMatrix(40, 40, factor=0.8, distribution=dist, normalization=norm,
                                     connectedness=4, format=fmt, random_seed=31 + idx)
        out[f"sparse_{idx}"] = dense(mk())
        cases.append(("sparse", 40, 40, 0.8, dist, norm, 31 + idx, fmt, 4.0))
        idx += 1
    # rectangular sparse input matrix by density
    mk = rmat.SparseRandomMatrix(40, 6, factor=0.3, distribution="uniform", normalization="svd",
                                 density=0.5, format="csr", random_seed=5)
    out["sparse_rect"] = dense(mk())
    out["case_table"] = np.array([repr(c) for c in cases])
    np.savez_compressed(os.path.join(OUT, "matrices.npz"), **out)


# --------------------------------------------------------------------------- #
ESN_KW = dict(
    input_kwargs=dict(factor=0.5, distribution="uniform", normalization="svd", random_seed=0),
    adjacency_kwargs=dict(factor=0.9, distribution="uniform", normalization="svd", is_sparse=True,
                          connectedness=5, format="csr", random_seed=1),
    bias_kwargs=dict(factor=0.5, distribution="uniform", random_seed=2),
)


def build_ref_esn(n_in, n_out, n_res, leak=0.5, beta=1e-6, **over):
    kw = {k: dict(v) for k, v in ESN_KW.items()}
    kw.update(over)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        esn = resn.ESN(n_in, n_out, n_res, leak, beta, **kw)
        esn.build()
    return esn


def ref_predict(esn, y, n_steps, n_spinup, Wout):
    # the loop of xesn/esn.py:251-268 around the reference's own _update
    kw = dict(W=esn.W, Win=esn.Win, bias_vector=esn.bias_vector, leak_rate=esn.leak_rate)
    yT = y.T
    r = np.zeros(esn.n_reservoir)
    for n in range(n_spinup):
        r = resn._update(r, yT[n], **kw)
    vT = np.zeros((n_steps + 1, esn.n_output))
    vT[0] = yT[n_spinup]
    for n in range(1, n_steps + 1):
        r = resn._update(r, vT[n - 1], **kw)
        vT[n] = Wout @ r
    return vT.T


def eager_small():
    out = {}
    T = 500
    k = 0
    for n_in in (3, 7):
        esn = build_ref_esn(n_in, n_in, 120)
        u = np.random.RandomState(0).normal(size=(n_in, T))
        out[f"u_{n_in}"] = u
        out[f"W_{n_in}"] = dense(esn.W)
        out[f"Win_{n_in}"] = esn.Win
        out[f"b_{n_in}"] = esn.bias_vector
        # states of a forced run (rT buffer of _train_1d, un-batched)
        kw = dict(W=esn.W, Win=esn.Win, bias_vector=esn.bias_vector, leak_rate=esn.leak_rate)
        r = np.zeros((61, 120))
        for t in range(60):
            r[t + 1] = resn._update(r[t], u[:, t], **kw)
        out[f"states_{n_in}"] = r
        for n_spinup, batch in itertools.product((0, 10), (None, 33, 10_000)):
            Wout = resn._train_1d(u, u, n_spinup, batch, esn.W, esn.Win, esn.bias_vector,
                                  esn.leak_rate, esn.tikhonov_parameter)
            tag = f"{n_in}_{n_spinup}_{batch}"
            out[f"Wout_{tag}"] = Wout
            k += 1
        Wout = out[f"Wout_{n_in}_10_33"]
        for n_spinup in (0, 10):
            out[f"pred_{n_in}_{n_spinup}"] = ref_predict(esn, u, 25, n_spinup, Wout)
    # separate targets (y != u), n_output != n_input
    esn = build_ref_esn(7, 3, 120)
    u = out["u_7"]
    y = np.random.RandomState(1).normal(size=(3, T))
    out["y_sep"] = y
    out["Wout_sep"] = resn._train_1d(u, y, 5, 100, esn.W, esn.Win, esn.bias_vector, esn.leak_rate,
                                     esn.tikhonov_parameter)
    # n_spinup == n_time -> all-zero readout
    out["Wout_allspin"] = resn._train_1d(u, y, T, None, esn.W, esn.Win, esn.bias_vector, esn.leak_rate,
                                         esn.tikhonov_parameter)
    np.savez_compressed(os.path.join(OUT, "eager_small.npz"), **out)


def eager_l96():
    """Lorenz-96 (40 vars), N=500: the shape of BASELINE config 1 scaled to a
    fixture that stays small.  leak 0.5, beta 1e-6 as in scaling/config-eager.yaml."""
    T = 2000
    data = lorenz96(40, T + 400)
    esn = build_ref_esn(40, 40, 500)
    train = data[:, :T]
    Wout = resn._train_1d(train, train, 0, 700, esn.W, esn.Win, esn.bias_vector, esn.leak_rate,
                          esn.tikhonov_parameter)
    test = data[:, T - 100:]
    pred = ref_predict(esn, test, 200, 100, Wout)
    np.savez_compressed(os.path.join(OUT, "eager_l96.npz"), data=data.astype(np.float64), Wout=Wout, pred=pred,
                        W=dense(esn.W), Win=esn.Win, b=esn.bias_vector,
                        n_train=T, batch=700, test_offset=T - 100, n_steps=200, n_spinup=100)


# --------------------------------------------------------------------------- #
def pad(field, depth, boundary):
    out = field
    for ax in range(field.ndim):
        d = depth.get(ax, 0)
        if d == 0:
            continue
        kind = boundary[ax] if isinstance(boundary, dict) else boundary
        w = [(0, 0)] * out.ndim
        w[ax] = (d, d)
        if kind == "periodic":
            out = np.pad(out, w, mode="wrap")
        elif kind == "reflect":
            out = np.pad(out, w, mode="symmetric")
        else:
            out = np.pad(out, w, mode="constant", constant_values=kind)
    return out


def blocks(field, chunks, depth, boundary):
    nsp = field.ndim - 1
    p = pad(field, depth, boundary)
    grid = [field.shape[a] // chunks[a] for a in range(nsp)]
    for g in itertools.product(*[range(n) for n in grid]):
        sl = tuple(slice(g[a] * chunks[a], g[a] * chunks[a] + chunks[a] + 2 * depth.get(a, 0)) for a in range(nsp))
        yield g, p[sl + (slice(None),)]


LAZY_CASES = {
    # name: (shape, chunks, overlap, boundary, n_reservoir, nan_cells)
    "x_periodic": ((12,), (3,), (1,), "periodic", 60, ()),
    "x_const0": ((4,), (2,), (1,), 0.0, 50, ()),
    "x_nan": ((9,), (3,), (2,), float("nan"), 50, ()),
    "xy_periodic": ((6, 8), (3, 4), (1, 1), "periodic", 80, ()),
    "xy_mixed": ((6, 8), (3, 2), (2, 1), {0: "reflect", 1: 0.5}, 64, ()),
    "xy_one_axis": ((6, 4), (2, 4), (1, 0), "periodic", 48, ()),
    "xyz": ((4, 6, 2), (2, 3, 2), (1, 1, 0), {0: "periodic", 1: 0.0, 2: "reflect"}, 56, ()),
    "xy_land": ((6, 6), (3, 3), (1, 1), "periodic", 64, ((1, 1), (4, 2))),
}


def lazy_cases():
    T = 300
    for name, (shape, chunks, ovl, boundary, N, nan_cells) in LAZY_CASES.items():
        nsp = len(shape)
        rs = np.random.RandomState(sum(map(ord, name)))
        field = rs.normal(size=shape + (T,))
        for c in nan_cells:
            field[c] = np.nan
        depth = {a: ovl[a] for a in range(nsp)}
        depth[nsp] = 0
        n_out = int(np.prod(chunks))
        n_in = int(np.prod([chunks[a] + 2 * ovl[a] for a in range(nsp)]))
        esn = build_ref_esn(n_in, n_out, N, leak=0.6, beta=1e-4)
        kw = dict(W=esn.W, Win=esn.Win, bias_vector=esn.bias_vector, leak_rate=esn.leak_rate)
        grid = tuple(shape[a] // chunks[a] for a in range(nsp))

        # --- train: reference _train_nd on every halo'd block
        n_spinup_train, batch = 7, 64
        wout = np.zeros(grid + (n_out, N))
        for g, blk in blocks(field, chunks, depth, boundary):
            w = rlazy._train_nd(blk, overlap=depth, n_spinup=n_spinup_train, batch_size=batch,
                                tikhonov_parameter=esn.tikhonov_parameter,
                                block_info={None: {"chunk-shape": (n_out, N)}}, **kw)
            wout[g] = w

        # --- predict: reference _spinup / _update_nd / _readout per block
        n_spinup, n_steps = 12, 15
        truth = field[..., :n_spinup + 1]
        states = {}
        for g, blk in blocks(truth, chunks, depth, boundary):
            states[g] = rlazy._spinup(blk, n_spinup=n_spinup, block_info={None: {"chunk-shape": (N,)}}, **kw)
        pred = np.zeros(shape + (n_steps + 1,))
        pred[..., 0] = truth[..., n_spinup]
        cur = truth[..., n_spinup]
        for n in range(1, n_steps + 1):
            nxt = np.zeros(shape)
            for g, blk in blocks(cur[..., None], chunks, depth, boundary):
                states[g] = rlazy._update_nd(states[g], blk[..., 0], **kw)
                v = rlazy._readout(wout[g], states[g], block_info={None: {"chunk-shape": tuple(chunks)}})
                sl = tuple(slice(g[a] * chunks[a], (g[a] + 1) * chunks[a]) for a in range(nsp))
                nxt[sl] = v
            pred[..., n] = nxt
            cur = nxt

        bkeys = np.array(sorted(boundary.keys())) if isinstance(boundary, dict) else np.array([])
        bvals = np.array([str(boundary[k]) for k in sorted(boundary.keys())]) if isinstance(boundary, dict) \
            else np.array([str(boundary)])
        np.savez_compressed(
            os.path.join(OUT, f"lazy_{name}.npz"),
            field=field, chunks=np.array(chunks), overlap=np.array(ovl), boundary_keys=bkeys,
            boundary_vals=bvals, n_reservoir=N, leak=0.6, beta=1e-4, n_spinup_train=n_spinup_train,
            batch=batch, n_spinup=n_spinup, n_steps=n_steps, wout_groups=wout, pred=pred,
            W=dense(esn.W), Win=esn.Win, b=esn.bias_vector,
            inner_mask=rlazy._get_inner_mask(next(blocks(field, chunks, depth, boundary))[1].shape, depth),
        )


def psd_cases():
    """Power spectral densities from the reference's own `psd` (xesn/psd.py:12-70, run unmodified over a stub
    DataArray): one, two and three spatial axes, even and odd sizes."""
    psd = ref_shim.load_psd()
    rs = np.random.RandomState(11)
    out = {}
    for name, shape, dims in (("x12", (12, 6), ("x", "time")), ("x13", (13, 5), ("x", "ftime")),
                              ("xy8", (8, 8, 5), ("x", "y", "time")), ("xy9", (9, 9, 4), ("x", "y", "time")),
                              ("xyz8", (8, 8, 3, 4), ("x", "y", "z", "time"))):
        x = rs.normal(size=shape)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = psd(ref_shim.field_array(x, dims))
        out[f"in_{name}"] = x
        out[f"psd_{name}"] = np.asarray(res.data)
        out[f"k_{name}"] = np.asarray(res["k1d"].data)
    np.savez_compressed(os.path.join(OUT, "psd.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-psd" in sys.argv:
        psd_cases()
        sys.exit(0)
    matrices()
    eager_small()
    eager_l96()
    lazy_cases()
    psd_cases()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures written to", OUT, f"({total/1e6:.2f} MB)")
