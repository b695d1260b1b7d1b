"""xesn_b200.matrix reproduces the reference's seed -> matrix conventions
(reference xesn/matrix.py; fixtures from oracle/make_golden.py; known answers from
the reference docstrings matrix.py:33-72,:204-214; properties from
xesn/test/matrix.py:73-181)."""
import ast

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse.linalg

from tests.conftest import load_golden
from xesn_b200.matrix import RandomMatrix, SparseRandomMatrix


def _dense(A):
    return A.toarray() if hasattr(A, "toarray") else A


def test_docstring_known_answers():
    m = RandomMatrix(2, 2, factor=5, distribution="uniform", random_seed=0)
    np.testing.assert_allclose(m(), [[0.48813504, 2.15189366], [1.02763376, 0.44883183]], atol=5e-9)
    np.testing.assert_allclose(m(), [[-0.76345201, 1.45894113], [-0.62412789, 3.91773001]], atol=5e-9)
    A = RandomMatrix(2, 2, factor=0.9, distribution="normal", normalization="eig", random_seed=0)()
    np.testing.assert_allclose(A, [[0.59414168, 0.13477495], [0.32964386, 0.75474407]], atol=5e-9)
    np.testing.assert_allclose(max(abs(scipy.linalg.eigvals(A))), 0.9, rtol=1e-14)
    A = RandomMatrix(2, 2, factor=1.0, distribution="gaussian", normalization="svd", random_seed=0)()
    np.testing.assert_allclose(A, [[0.64082821, 0.14536532], [0.35554665, 0.81405043]], atol=5e-9)
    mk = SparseRandomMatrix(100, 100, factor=0.9, distribution="normal", normalization="eig",
                            connectedness=1, random_seed=0)
    S = mk()
    assert S.nnz == 100 and S.format == "coo" and mk.density == 0.01
    rho = max(abs(scipy.sparse.linalg.eigs(S, k=1, which="LM", return_eigenvectors=False)))
    np.testing.assert_allclose(rho, 0.9, rtol=1e-12)


def test_golden_kats_bitwise():
    z = load_golden("matrices.npz")
    m = RandomMatrix(2, 2, factor=5, distribution="uniform", random_seed=0)
    assert np.array_equal(m(), z["kat_uniform_first"])
    assert np.array_equal(m(), z["kat_uniform_second"])
    A = RandomMatrix(2, 2, factor=0.9, distribution="gaussian", normalization="eig", random_seed=0)()
    np.testing.assert_allclose(A, z["kat_gauss_eig"], rtol=1e-14)
    A = RandomMatrix(2, 2, factor=1.0, distribution="gaussian", normalization="svd", random_seed=0)()
    np.testing.assert_allclose(A, z["kat_gauss_svd"], rtol=1e-14)


def test_golden_case_table():
    z = load_golden("matrices.npz")
    for i, row in enumerate(z["case_table"]):
        kind, nr, nc, factor, dist, norm, seed, fmt, conn = ast.literal_eval(str(row))
        if kind == "dense":
            A = RandomMatrix(nr, nc, factor=factor, distribution=dist, normalization=norm, random_seed=seed)()
            ref = z[f"dense_{i}"]
        else:
            A = SparseRandomMatrix(nr, nc, factor=factor, distribution=dist, normalization=norm,
                                   connectedness=conn, format=fmt, random_seed=seed)()
            assert A.format == fmt
            ref = z[f"sparse_{i}"]
        A = _dense(A)
        # same nonzero pattern, bit-identical for "multiply"; ARPACK/LAPACK scalars may jitter by ulps
        assert np.array_equal(A != 0, ref != 0)
        if norm == "multiply":
            assert np.array_equal(A, ref)
        else:
            np.testing.assert_allclose(A, ref, rtol=1e-12)
    A = SparseRandomMatrix(40, 6, factor=0.3, distribution="uniform", normalization="svd", density=0.5,
                           format="csr", random_seed=5)()
    np.testing.assert_allclose(_dense(A), z["sparse_rect"], rtol=1e-12)


@pytest.mark.parametrize("distribution", ("uniform", "normal"))
@pytest.mark.parametrize("normalization,fn", (
    ("svd", lambda A: np.max(scipy.linalg.svd(_dense(A), compute_uv=False))),
    ("eig", lambda A: np.max(np.abs(scipy.linalg.eigvals(_dense(A))))),
))
@pytest.mark.parametrize("sparse", (False, True))
def test_normalisation_hits_factor(distribution, normalization, fn, sparse):
    # xesn/test/matrix.py:107-118
    if sparse:
        A = SparseRandomMatrix(60, 60, factor=0.37, distribution=distribution, normalization=normalization,
                               density=0.3, random_seed=3)()
    else:
        A = RandomMatrix(60, 60, factor=0.37, distribution=distribution, normalization=normalization,
                         random_seed=3)()
    np.testing.assert_allclose(fn(A), 0.37, rtol=1e-7)


def test_density_bookkeeping():
    # xesn/test/matrix.py:137-181
    for kw, dens in ((dict(density=0.1), 0.1), (dict(sparsity=0.9), 1 - 0.9), (dict(connectedness=5), 0.1)):
        mk = SparseRandomMatrix(50, 50, factor=1.0, distribution="uniform", random_seed=0, **kw)
        np.testing.assert_allclose(mk.density, dens)
        assert mk().nnz == int(round(mk.density * 2500))
    with pytest.raises(TypeError):
        SparseRandomMatrix(5, 5, factor=1.0, distribution="uniform")
    with pytest.raises(TypeError):
        SparseRandomMatrix(5, 5, factor=1.0, distribution="uniform", density=0.1, sparsity=0.9)
    with pytest.raises(TypeError):
        SparseRandomMatrix(5, 5, factor=1.0, distribution="uniform", sparsity=0.9, connectedness=2)
    with pytest.raises(TypeError):
        SparseRandomMatrix(5, 6, factor=1.0, distribution="uniform", connectedness=2)


def test_bad_options():
    with pytest.raises(NotImplementedError):
        RandomMatrix(2, 2, factor=1.0, distribution="cauchy")
    with pytest.raises(NotImplementedError):
        RandomMatrix(2, 2, factor=1.0, distribution="uniform", normalization="frobenius")
    with pytest.raises(NotImplementedError):
        SparseRandomMatrix(2, 2, factor=1.0, distribution="uniform", normalization="nope", density=0.5)
