"""Functional NumPy restatement of the reference's block-vector algebra.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function works on plain
2-D ndarrays whose ROWS are vectors -- ``S`` is the selected window of `self`
as an ``(m, n)`` array, ``O`` the selected window of `other` as ``(k, n)`` --
and cites the reference lines it restates (paths relative to the reference
checkout, raleigh/algebra/...).
"""
import numpy as np
import scipy.sparse as sp

__all__ = [
    'gram', 'combine', 'add_scaled', 'add_combined', 'add_per_vector',
    'scale_rows', 'row_dots', 'column_dots', 'gather_rows', 'thin_svd',
    'project_out', 'dense_apply', 'row_sqnorms', 'sym_upper_csr',
    'sym_spmm', 'jacobi_apply', 'uniform_fill', 'uniform_fill_cublas',
    'lap3d_csr', 'lap3d_eigenvalues',
]


def _cj(a):
    return a.conj() if a.dtype.kind == 'c' else a


def gram(S, O):
    """Vectors.dot: G[i, j] = <other_i, self_j>, shape (k, m), conj on other.
    dense_numpy.py:78-82."""
    return _cj(O) @ S.T


def combine(S, q):
    """Vectors.multiply: out = q^T . S with q of shape (m, m_out).
    dense_numpy.py:84-93."""
    return q.T @ S


def add_scaled(S, O, s):
    """Vectors.add, scalar s, q None: S + s*O.  dense_numpy.py:97-99."""
    return S + s * O


def add_combined(S, O, s, q):
    """Vectors.add, scalar s, q (k, m): S + s * q^T . O.  dense_numpy.py:100-101."""
    return S + s * (q.T @ O)


def add_per_vector(S, O, s):
    """Vectors.add, array s: S[i] + s[i]*O[i] (q ignored).  dense_numpy.py:103-105."""
    s = np.asarray(s).reshape(-1)[:S.shape[0]]
    return S + s[:, None] * O


def scale_rows(S, s, multiply=False):
    """Vectors.scale: multiply rows by s[i], or divide skipping s[i] == 0.
    dense_numpy.py:44-52."""
    s = np.asarray(s).reshape(-1)[:S.shape[0]]
    out = S.copy()
    if multiply:
        out *= s[:, None]
    else:
        nz = s != 0
        out[nz] = out[nz] / s[nz, None]
    return out


def row_dots(S, O):
    """Vectors.dots(transp=False): w[i] = sum_j conj(O[i,j]) S[i,j].
    dense_numpy.py:68-76."""
    return np.einsum('ij,ij->i', _cj(O), S).astype(S.dtype)


def column_dots(S, O):
    """Vectors.dots(transp=True): w[j] = sum_i conj(O[i,j]) S[i,j], length n.
    dense_numpy.py:55-67."""
    return np.einsum('ij,ij->j', _cj(O), S).astype(S.dtype)


def gather_rows(all_self, ind):
    """Vectors.copy(other, ind): rows of the WHOLE self container picked by
    absolute indices; the caller writes them at other's selection start.
    dense_numpy.py:35-42."""
    return all_self[np.asarray(ind, dtype=np.int64), :]


def thin_svd(S):
    """Vectors.svd: S = v diag(sigma) wt; self <- wt; returns (sigma, conj(v), wt).
    dense_numpy.py:125-128."""
    v, sigma, wt = np.linalg.svd(S, full_matrices=False)
    return sigma, _cj(v), wt


def project_out(S, O):
    """Vectors.orthogonalize: q = conj(O) S^T (k, m); S - q^T O; returns (S_new, q).
    dense_numpy.py:117-123."""
    q = _cj(O) @ S.T
    return S - q.T @ O, q


def dense_apply(A, X, transp=False):
    """Matrix.apply: y = x . A^T (A is (M, N), x has dimension N), or
    y = x . conj(A) when transp (x has dimension M).
    dense_numpy.py:153-175 / dense_cublas.py:732-776."""
    if transp:
        return X @ _cj(A)
    return X @ A.T


def row_sqnorms(A):
    """Matrix.dots: squared 2-norms of the rows.  dense_numpy.py:177-179."""
    return np.einsum('ij,ij->i', _cj(A), A).real.astype(A.dtype) if A.dtype.kind == 'c' \
        else np.einsum('ij,ij->i', A, A).astype(A.dtype)


def sym_upper_csr(A):
    """What SparseSymmetricMatrix keeps: the upper triangle as sorted CSR.
    sparse_mkl.py:18-31 (0-based here; the reference adds 1 for MKL)."""
    u = sp.triu(A, format='csr')
    u.sort_indices()
    return u


def sym_spmm(U, X):
    """SparseSymmetricMatrix.apply: Y = (A_sym . X^T)^T where A_sym mirrors the
    stored upper triangle U (MKL csrmm, descr 'SUNF' / 'HUNF', alpha=1, beta=0;
    X, Y are (m, n) C-order = column-major n x m).  mkl_wrap.py:246-276."""
    strict = sp.triu(U, k=1, format='csr')
    full = (U + strict.conj().T).tocsr()
    return np.ascontiguousarray((full @ X.T).T)


def jacobi_apply(diag, X):
    """Diagonal (Jacobi) preconditioner y = x * (1/diag(A)) handed through
    Operator.apply.  sparse_mkl.py:143-154 (+ user T.apply contract,
    partial_hevp.py:64-73)."""
    return X * (1.0 / diag)[None, :]


def uniform_fill(m, n, dtype=np.float64):
    """fill_random of the NumPy backend: 2*rand(m, n) - 1 from the global host
    RNG.  dense_ndarray.py:34-37."""
    return (2 * np.random.rand(m, n) - 1).astype(dtype)


def uniform_fill_cublas(m, n, dtype=np.float64):
    """fill_random of the cuBLAS backend: rand(m, n).astype(dt)*2 - 1 (differs
    from uniform_fill only by fp32 rounding).  dense_cublas.py:119-131."""
    data = np.random.rand(m, n).astype(dtype)
    data *= 2
    data -= 1
    return data


def lap3d_csr(nx, ny, nz, ax=1.0, ay=1.0, az=1.0, dtype=np.float64):
    """7-point finite-difference Laplacian on the box ax x ay x az with Dirichlet
    boundary, x fastest.  Restates examples/laplace.py:10-27 via Kronecker sums."""
    def lap1d(n, a):
        h = a / (n + 1)
        d = 1.0 / (h * h)
        return sp.diags([-d * np.ones(n - 1), 2 * d * np.ones(n), -d * np.ones(n - 1)],
                        [-1, 0, 1], format='csr')
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    L = sp.kron(Iz, sp.kron(Iy, lap1d(nx, ax))) + sp.kron(Iz, sp.kron(lap1d(ny, ay), Ix)) \
        + sp.kron(lap1d(nz, az), sp.kron(Iy, Ix))
    L = L.tocsr().astype(dtype)
    L.sort_indices()
    return L


def lap3d_eigenvalues(nx, ny, nz, ax=1.0, ay=1.0, az=1.0):
    """Analytic spectrum of lap3d_csr, ascending."""
    def ev(n, a):
        h = a / (n + 1)
        k = np.arange(1, n + 1)
        return (2 - 2 * np.cos(np.pi * k / (n + 1))) / (h * h)
    lx, ly, lz = ev(nx, ax), ev(ny, ay), ev(nz, az)
    return np.sort((lz[:, None, None] + ly[None, :, None] + lx[None, None, :]).ravel())
