"""Seed -> matrix conventions of the reference (host side, numpy/scipy).

Drop-in for `xesn.RandomMatrix` / `xesn.SparseRandomMatrix`
(reference xesn/matrix.py:16-193 and :196-316).  These run once per `build()`,
are not part of the timed path, and must reproduce the reference's matrices
exactly, because the random seeds *are* the checkpoint format (xesn/io.py:13-53
re-runs `build()` from the stored kwargs).  So the same numpy `RandomState` and
scipy routines are called, in the same order:

* dense:  `RandomState.uniform(-1, 1)` or `.normal(0, 1)` of shape (rows, cols)
  (matrix.py:133-142);
* sparse: `scipy.sparse.random(rows, cols, density, format, random_state=rs,
  data_rvs=rs.rand | rs.randn)` -- note sparse "uniform" is [0, 1), not [-1, 1)
  (matrix.py:279-291);
* rescaling: `factor * A`, `factor / sigma_max * A` ("svd") or
  `factor / rho * A` ("eig"), with LAPACK svd/eigvals for dense and ARPACK
  svds/eigs (k=1, "LM") for sparse matrices (matrix.py:183-193, :296-316).
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

_GAUSSIAN_NAMES = ("gaussian", "normal")
_DISTRIBUTIONS = ("uniform",) + _GAUSSIAN_NAMES
_NORMALIZATIONS = ("svd", "eig", "multiply")


def _dense_scale(A, how):
    if how == "svd":
        return np.max(np.abs(scipy.linalg.svd(A, compute_uv=False, full_matrices=False)))
    if how == "eig":
        return np.max(np.abs(scipy.linalg.eigvals(A)))
    return 1.0


def _sparse_scale(A, how):
    if how == "svd":
        s = scipy.sparse.linalg.svds(A, k=1, which="LM", return_singular_vectors=False)
        return np.max(np.abs(s))
    if how == "eig":
        s = scipy.sparse.linalg.eigs(A, k=1, which="LM", return_eigenvectors=False)
        return np.max(np.abs(s))
    return 1.0


class RandomMatrix:
    """Dense random matrix factory; every call draws a new matrix from the
    object's `RandomState` (reference docstring known-answers at
    matrix.py:33-72 are reproduced bit for bit, see tests/test_matrix.py)."""

    __slots__ = ("n_rows", "n_cols", "distribution", "normalization", "factor",
                 "random_seed", "random_state")

    def __init__(self, n_rows, n_cols, factor, distribution, normalization="multiply", random_seed=None):
        if distribution not in _DISTRIBUTIONS:
            raise NotImplementedError(
                f"RandomMatrix.__init__: '{distribution}' not recognized, only 'uniform', "
                f"'gaussian'/'normal' are implemented")
        if normalization not in _NORMALIZATIONS:
            raise NotImplementedError(
                f"RandomMatrix.__init__: '{normalization}' not recognized, only 'svd', 'eig', "
                f"and 'multiply' are implemented")
        self.n_rows, self.n_cols = n_rows, n_cols
        self.factor = factor
        self.distribution = distribution
        self.normalization = normalization
        self.random_seed = random_seed
        self.random_state = np.random.RandomState(random_seed)

    def __call__(self):
        return self.normalize(self.create_matrix())

    def create_matrix(self):
        shape = (self.n_rows, self.n_cols)
        if self.distribution == "uniform":
            return self.random_state.uniform(low=-1.0, high=1.0, size=shape)
        return self.random_state.normal(loc=0.0, scale=1.0, size=shape)

    def normalize(self, A):
        return self.factor / _dense_scale(A, self.normalization) * A


class SparseRandomMatrix(RandomMatrix):
    """Sparse random matrix factory.  Exactly one of `density`, `sparsity`
    (= 1 - density) or `connectedness` (= density * n_rows, square only) fixes
    the number of stored entries (matrix.py:252-274)."""

    __slots__ = ("density", "format")

    def __init__(self, n_rows, n_cols, factor, distribution, normalization="multiply",
                 density=None, sparsity=None, connectedness=None, format="coo", random_seed=None):
        super().__init__(n_rows, n_cols, factor, distribution, normalization, random_seed)
        self.format = format
        given = {k: v for k, v in (("density", density), ("sparsity", sparsity),
                                   ("connectedness", connectedness)) if v is not None}
        if not given:
            raise TypeError("SparseRandomMatrix.__init__: specify matrix density by one of the following "
                            "arguments: 'sparsity', 'density', or 'connectedness'")
        if len(given) > 1:
            first = next(k for k in ("density", "sparsity", "connectedness") if k in given)
            others = " and/or ".join(f"'{k}'" for k in ("density", "sparsity", "connectedness") if k != first)
            raise TypeError(f"SparseRandomMatrix.__init__: {first} set to {given[first]}, "
                            f"cannot also specify {others}")
        if density is not None:
            self.density = density
        elif sparsity is not None:
            self.density = 1.0 - sparsity
        else:
            if n_rows != n_cols:
                raise TypeError("SparseRandomMatrix.__init__: matrix connectivity does not make sense for "
                                "non-square matrix, use 'density' or 'sparsity' instead")
            self.density = connectedness / n_rows

    def create_matrix(self):
        rs = self.random_state
        values = rs.rand if self.distribution == "uniform" else rs.randn
        return scipy.sparse.random(self.n_rows, self.n_cols, density=self.density, format=self.format,
                                   random_state=rs, data_rvs=values)

    def normalize(self, A):
        A.data = self.factor / _sparse_scale(A, self.normalization) * A.data
        return A
