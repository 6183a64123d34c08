"""Host (NumPy) Vectors / Matrix / sparse operators with the reference interface.

TEST INFRASTRUCTURE (see oracle/__init__.py): a CPU port of the reference's
dense_numpy.Vectors / Matrix (dense_numpy.py:12-184, dense_ndarray.py:11-151)
and sparse_mkl.SparseSymmetricMatrix / Operator (sparse_mkl.py:16-48, 143-154)
assembled from the functional restatements in algebra_np.py.  It exists so that
(1) parity tests can drive the CUDA backend and a CPU twin through identical
call sequences, and (2) bench.py's cpu_baseline / --impl reference legs can
time the CPU path when the reference's own dense_numpy is not on the box.
"""
import numbers

import numpy as np
import scipy.sparse as sp

from . import algebra_np as K


class Vectors:
    """Row-block of vectors with a (first, count) selection window."""

    def __init__(self, arg, nvec=0, data_type=None, shallow=False):
        if isinstance(arg, Vectors):
            win = arg._window()
            self._buf = win if shallow else win.copy()
        elif isinstance(arg, Matrix):
            if arg.order() != 'C_CONTIGUOUS':
                raise ValueError('Vectors data must be C_CONTIGUOUS')
            self._buf = arg.data() if shallow else arg.data().copy()
        elif isinstance(arg, np.ndarray):
            self._buf = arg
        elif isinstance(arg, numbers.Number):
            dt = np.float64 if data_type is None else data_type
            self._buf = np.zeros((nvec, int(arg)), dtype=dt)
        else:
            raise ValueError('wrong argument %s in constructor' % repr(type(arg)))
        self._sel = (0, self._buf.shape[0])

    # ---- bookkeeping (dense_ndarray.py:18-32, 85-89)
    def dimension(self):
        return self._buf.shape[1]

    def nvec(self):
        return self._sel[1]

    def shape(self):
        return self._buf.shape

    def select(self, nv, first=0):
        assert nv <= self._buf.shape[0] and first >= 0
        self._sel = (first, nv)

    def select_all(self):
        self.select(self._buf.shape[0])

    def selected(self):
        return self._sel

    def first(self):
        return self._sel[0]

    def data_type(self):
        return self._buf.dtype.type

    def is_complex(self):
        return self._buf.dtype.kind == 'c'

    def _window(self):
        f, n = self._sel
        return self._buf[f:f + n, :]

    def all_data(self):
        return self._buf

    def data(self, i=None):
        w = self._window()
        return w if i is None else w[i, :]

    def asarray(self):
        return self.data().T

    # ---- construction helpers (dense_numpy.py:19-33, 113-115)
    def new_vectors(self, arg=0, dim=None):
        if isinstance(arg, np.ndarray):
            return Vectors(arg)
        return Vectors(self.dimension() if dim is None else dim, int(arg), self.data_type())

    def clone(self):
        return Vectors(self)

    def reference(self):
        return Vectors(self, shallow=True)

    def zero(self):
        self._window()[...] = 0

    def fill(self, array_or_value):
        self._window()[...] = array_or_value

    def fill_random(self):
        m, n = self._window().shape
        self._window()[...] = K.uniform_fill(m, n, self._buf.dtype)

    def append(self, other, axis=0):
        # dense_ndarray.py:39-47
        if axis == 0:
            self._buf = np.concatenate((self.data(), other.data()))
            self.select_all()
        else:
            self._buf = np.concatenate((self._buf, other.all_data()), axis=1)

    # ---- algebra
    def copy(self, other, ind=None):
        if ind is None:
            assert other.nvec() == self.nvec()
            other._window()[...] = self._window()
        else:
            j = other.first()
            other._buf[j:j + len(ind), :] = K.gather_rows(self._buf, ind)

    def scale(self, s, multiply=False):
        self._window()[...] = K.scale_rows(self._window(), s, multiply)

    def dots(self, other, transp=False):
        if transp:
            return K.column_dots(self._window(), other._window())
        return K.row_dots(self._window(), other._window())

    def dot(self, other):
        return K.gram(self._window(), other._window())

    def multiply(self, q, output):
        assert output.nvec() == q.shape[1]
        output._window()[...] = K.combine(self._window(), q)

    def add(self, other, s, q=None):
        if np.isscalar(s):
            if q is None:
                self._window()[...] = K.add_scaled(self._window(), other._window(), s)
            else:
                self._window()[...] = K.add_combined(self._window(), other._window(), s, q)
        else:
            self._window()[...] = K.add_per_vector(self._window(), other._window(), s)

    def orthogonalize(self, other):
        new, q = K.project_out(self._window(), other._window())
        self._window()[...] = new
        return self.new_vectors(q)

    def svd(self):
        sigma, vc, wt = K.thin_svd(self._window())
        self._window()[...] = wt
        return sigma, vc


class Matrix:
    """Dense operator holder (dense_ndarray.py:117-151, dense_numpy.py:151-184)."""

    def __init__(self, arg):
        if isinstance(arg, Vectors):
            data = arg.data()
        elif isinstance(arg, np.ndarray):
            data = arg
        else:
            raise ValueError('wrong argument %s in Matrix constructor' % repr(type(arg)))
        if data.flags['C_CONTIGUOUS']:
            self._order = 'C_CONTIGUOUS'
        elif data.flags['F_CONTIGUOUS']:
            self._order = 'F_CONTIGUOUS'
        else:
            raise ValueError('Matrix data must be either C- or F-contiguous')
        self._a = data

    def data(self):
        return self._a

    def shape(self):
        return self._a.shape

    def order(self):
        return self._order

    def data_type(self):
        return self._a.dtype.type

    def is_complex(self):
        return self._a.dtype.kind == 'c'

    def new_vectors(self, dim=None, nv=0):
        return Vectors(self._a.shape[1] if dim is None else dim, nv, self.data_type())

    def apply(self, x, y, transp=False):
        y._window()[...] = K.dense_apply(self._a, x._window(), transp)

    def dots(self):
        return K.row_sqnorms(self._a)


class SparseSymmetricMatrix:
    """sparse_mkl.py:16-48 with SciPy standing in for mkl_?csrmm."""

    def __init__(self, matrix):
        if isinstance(matrix, SparseSymmetricMatrix):
            self._u = matrix.csr()
        else:
            self._u = K.sym_upper_csr(matrix)
        strict = sp.triu(self._u, k=1, format='csr')
        self._full = (self._u + strict.conj().T).tocsr()

    def size(self):
        return self._u.shape[0]

    def data_type(self):
        return self._u.data.dtype

    def csr(self):
        return self._u

    def apply(self, x, y):
        y._window()[...] = (self._full @ x._window().T).T


class Operator:
    """sparse_mkl.py:143-154: hands 2-D ndarrays to a user object with apply(x, y)."""

    def __init__(self, op):
        self._op = op

    def apply(self, x, y):
        self._op.apply(x.data(), y.data())


class Jacobi:
    """User preconditioner T for partial_hevp: y = x / diag(A)
    (contract: partial_hevp.py:64-73)."""

    def __init__(self, A):
        self._idiag = 1.0 / np.asarray(A.diagonal())

    def apply(self, x, y):
        y[...] = x * self._idiag[None, :]
