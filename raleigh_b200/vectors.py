"""B200 implementation of RALEIGH's abstract `Vectors` type and dense `Matrix`.

Drop-in for raleigh/algebra/dense_cublas.py (same class names, constructor
signatures, methods, argument meaning and error behaviour -- the contract is
the docstring at raleigh/core/solver.py:22-96 plus the extras the interfaces
use), so the reference's unmodified core solver, partial_hevp and pca run on
it.  Every method is one call into libraleigh_b200.so (hand-written sm_100a
kernels); where the reference loops over vectors in Python and issues one
cuBLAS call per vector, a single kernel handles the whole block.

Layout: vector-major, vector j / component r at ``base + (j*ld + r)*itemsize``
with ld padded to a 128-byte multiple; a selection window (first, nv) is a
contiguous sub-block, so windows cost nothing.

Semantics follow the NumPy backend (dense_numpy.py), which is the oracle every
other backend is diffed against in the reference's own tests.
"""
import ctypes
import numbers
import weakref

import numpy

from . import _lib
from ._lib import lib, check
from . import device as dev

MINMAX_ON_DEVICE_BYTES = 64 << 20     # host arrays at least this big get their min/max from the device copy
_recent_uploads = {}                   # id(host array) -> (weakref(Matrix), shape, dtype) of the latest big upload


def _np_type(t):
    return numpy.dtype(t).type


TC_MIN_VECTORS = 8                # below this the FMA-pipe / GEMV kernels are used

# lra.update (lra.py:287-290) grows the left factor by `data()` -> numpy.concatenate(axis=1) -> `new_vectors(ndarray)`:
# the whole factor (1.2 GB at the end of config 5) crosses PCIe twice per chunk.  While compat's hooked update runs,
# data() of a block at least this big returns a DeviceData handle instead of a host copy; compat's numpy proxy
# concatenates two handles on the device and new_vectors() adopts the result.  Any other use of a handle turns it
# into the host array it stands for (`__array__`), so the reference code stays correct whatever it does with it.
LAZY_DATA_MIN_BYTES = None
_SAMPLE_LOCAL = [False]            # True while SampleVectors builds its local block from the matrix


class DeviceData:
    """Stand-in for the ndarray `Vectors.data()` would return: a private device copy of the selected block."""

    def __init__(self, vec, owned=False):
        self._vec = vec if owned else Vectors(vec)        # private copy of the selected block (keeps the sharding)
        self._host = None

    # -- what the concatenation hook uses
    def concatenate(self, other):
        out = Vectors(self._vec)
        out.append(other._vec, axis=1)
        return DeviceData(out, owned=True)

    def take(self):
        """The block as Vectors (new_vectors(ndarray) semantics: a fresh container).  A device copy -- 0.3 ms per GB --
        so that the handle keeps standing for the same data whatever happens to the new container."""
        return Vectors(self._vec)

    # -- ndarray behaviour on demand
    def _array(self):
        if self._host is None:
            saved = LAZY_DATA_MIN_BYTES
            globals()['LAZY_DATA_MIN_BYTES'] = None
            try:
                self._host = self._vec.data()
            finally:
                globals()['LAZY_DATA_MIN_BYTES'] = saved
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._array()
        return a if dtype is None else a.astype(dtype)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self._array(), name)

    def __getitem__(self, key):
        return self._array()[key]

    def __setitem__(self, key, value):
        self._array()[key] = value

    def __len__(self):
        return len(self._array())

    def __iter__(self):
        return iter(self._array())


def _device_data_binop(name):
    def op(self, *args):
        return getattr(self._array(), name)(*args)
    op.__name__ = name
    return op


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__', '__neg__',
           '__abs__', '__matmul__', '__rmatmul__', '__pow__', '__lt__', '__le__', '__gt__', '__ge__', '__eq__', '__ne__'):
    setattr(DeviceData, _n, _device_data_binop(_n))


def block_gemm(code, a_ptr, lda, M, N, x_ptr, ldx, y_ptr, ldy, k, transp, alpha=1.0, beta=0.0):
    """Y = alpha X A^T (transp = 0) or alpha X A (transp = 1) + beta Y with A an (M, N) row-major block:
    tcgen05 tensor cores (3xTF32, low parts split in shared memory) for fp32 whenever the operands
    are TMA addressable, the FMA-pipe kernel otherwise."""
    if (code == _lib.RL_F32 and k >= TC_MIN_VECTORS and
            lib.rl_dense_apply_tc_supported(a_ptr, lda, x_ptr, ldx)):
        wsb = lib.rl_dense_apply_tc_ws_bytes(M, N, k, transp)
        ws = dev.Buffer(wsb)
        check(lib.rl_dense_apply_tc(a_ptr, 0, lda, M, N, x_ptr, ldx, y_ptr, ldy, k, transp, alpha, beta, ws.ptr, wsb,
                                    dev.stream()))
        return
    check(lib.rl_dense_apply(code, a_ptr, lda, M, N, x_ptr, ldx, y_ptr, ldy, k, transp, alpha, beta, dev.stream()))


class Vectors:
    """Block of vectors resident in HBM.  dense_cublas.py:17-632."""

    _rl_device_block = True           # marks blocks the device-resident solver driver (jcg.py) can run on
    HOST_RNG_MAX_ELEMENTS = 1 << 24   # above this fill_random() switches to the device RNG
    MIN_INC = 16      # capacity growth policy of the reference (dense_cublas.py:424-425)
    MAX_INC = 1024

    # ------------------------------------------------------------------ ctor
    def __new__(cls, arg=None, nvec=0, data_type=None, shallow=False):
        if cls is Vectors and isinstance(arg, Matrix) and arg._mshard is not None and not _SAMPLE_LOCAL[0]:
            return SampleVectors(arg, shallow)      # not a Vectors instance: __init__ below is not run on it
        return object.__new__(cls)

    def __init__(self, arg, nvec=0, data_type=None, shallow=False):
        self._buf = None
        self._off = 0                       # first vector of this object inside the buffer
        if isinstance(arg, DeviceData):
            arg = numpy.asarray(arg)
        if isinstance(arg, Vectors):
            first, nv = arg.selected()
            self._set_type(arg.data_type())
            self._n, self._ng, self._shard = arg._n, arg._ng, arg._shard
            self._ld = arg._ld
            if shallow:
                # NumPy semantics (dense_ndarray.py:55-56): a view of the selected slice
                self._buf = arg._buf
                self._off = arg._off + first
                self._cap = nv
            else:
                self._alloc(nv)
                if nv > 0:
                    check(lib.rl_copy(self._code, self._ptr(0), self._ld, arg._ptr(first), arg._ld,
                                      nv, self._n, dev.stream()))
            self._nvec = nv
        elif isinstance(arg, Matrix):
            if arg.order() != 'C_CONTIGUOUS':
                raise ValueError('Vectors data must be C_CONTIGUOUS')
            m, n = arg._local_shape()
            self._set_type(arg.data_type())
            self._n, self._ng, self._shard = n, n, None     # rows of a (row-sharded) matrix: local vectors
            self._ld = arg._ld
            if shallow:
                self._buf = arg._buf        # alias the matrix memory (dense_cublas.py:369-376)
                self._off = arg._base
                self._cap = m
            else:
                self._alloc(m)
                if m > 0:
                    check(lib.rl_copy(self._code, self._ptr(0), self._ld, arg._aptr(), arg._ld, m, n,
                                      dev.stream()))
            self._nvec = m
        elif isinstance(arg, numpy.ndarray):
            if arg.ndim != 2:
                raise ValueError('Vectors data must be a 2D array')
            m, n = arg.shape
            self._set_type(arg.dtype.type)
            self._resolve_layout(n)
            self._ld = dev.padded_ld(self._n, self._w)
            self._alloc(m)
            self._nvec = m
            if m > 0:
                dev.upload_2d(self._ptr(0), self._ld * self._w, numpy.ascontiguousarray(self._local_part(arg)))
        elif isinstance(arg, numbers.Number):
            self._set_type(numpy.float64 if data_type is None else data_type)
            self._resolve_layout(int(arg))
            assert nvec >= 0
            self._ld = dev.padded_ld(self._n, self._w)
            self._alloc(int(nvec), zero=True)
            self._nvec = int(nvec)
        else:
            raise ValueError('wrong argument %s in constructor' % repr(type(arg)))
        self._sel = (0, self._nvec)
        self._min_inc = Vectors.MIN_INC

    def _resolve_layout(self, n_logical):
        """Logical (global) dimension -> local length.  A dimension registered with
        the active ShardContext (dist.py) is row-sharded: this process holds rows
        [row0, row0 + nloc) and reductions are all-reduced."""
        from . import dist
        ctx = dist.current()
        part = ctx.lookup(n_logical) if ctx is not None else None
        self._ng = int(n_logical)
        if part is None:
            self._n, self._shard = int(n_logical), None
        else:
            self._n, self._shard = part[1], (ctx, part[0])

    def _local_part(self, a):
        """Columns of a host array of logical width that this process owns."""
        if self._shard is None:
            return a
        if a.shape[1] == self._ng:
            row0 = self._shard[1]
            return a[:, row0:row0 + self._n]
        if a.shape[1] == self._n:
            return a                      # already the local part
        raise ValueError('array width %d matches neither the global (%d) nor the local (%d) dimension'
                         % (a.shape[1], self._ng, self._n))

    def _reduce_host(self, a):
        """Sum a small host result over the ranks that share this sharded block."""
        if self._shard is None:
            return a
        return self._shard[0].allreduce_host(a)

    def _reduce_device(self, buf, count, np_dtype):
        """All-reduce `count` elements of a device Buffer in place (NCCL over NVLink)."""
        if self._shard is None:
            return
        import torch
        tdt = torch.float32 if numpy.dtype(np_dtype) == numpy.float32 else torch.float64
        view = buf.tensor[:count * numpy.dtype(np_dtype).itemsize].view(tdt)
        self._shard[0].allreduce_(view)

    def is_sharded(self):
        return self._shard is not None

    def local_dimension(self):
        return self._n

    def _set_type(self, t):
        t = _np_type(t)
        if t not in (numpy.float32, numpy.float64):
            raise ValueError('data type %s not supported' % repr(t))
        self._dtype = t
        self._code = _lib.dtype_code(t)
        self._w = 4 if t is numpy.float32 else 8

    def _alloc(self, nvec, zero=False):
        dev.require_cuda()
        self._cap = nvec
        self._off = 0
        if nvec > 0:
            self._buf = dev.Buffer(nvec * self._ld * self._w, zero=zero)
        else:
            self._buf = None

    def _ptr(self, j):
        """Device address of vector j (absolute index inside this object)."""
        return self._buf.ptr + (self._off + int(j)) * self._ld * self._w

    def _touch(self):
        """Record a write to the underlying buffer (a Matrix sharing it keeps a
        derived low-part copy for the tensor-core path and must refresh it)."""
        if self._buf is not None:
            self._buf.version += 1

    def _wptr(self):
        """Device address of the selected window (0 if nothing is allocated)."""
        if self._buf is None:
            return 0
        return self._ptr(self._sel[0])

    # ------------------------------------------- methods required by the solver
    def new_vectors(self, arg=0, dim=None):
        if isinstance(arg, numbers.Number):
            return Vectors(self.dimension() if dim is None else dim, int(arg), self.data_type())
        if isinstance(arg, DeviceData):
            return arg.take()
        return Vectors(arg)

    @staticmethod
    def _local(n, nvec, data_type):
        """Unsharded block whatever the active ShardContext says about `n`."""
        from . import dist
        saved, dist._current = dist._current, None
        try:
            return Vectors(n, nvec, data_type)
        finally:
            dist._current = saved

    def clone(self):
        return Vectors(self)

    def dimension(self):
        return self._ng

    def nvec(self):
        return self._sel[1]

    def select(self, nv, first=0):
        assert nv <= self._nvec and first >= 0
        self._sel = (int(first), int(nv))       # callers pass NumPy integers (lra.py:364)

    def selected(self):
        return self._sel

    def data_type(self):
        return self._dtype

    def fill_random(self):
        """Host NumPy RNG + H2D, exactly as the reference's GPU backend does
        (dense_cublas.py:119-131), so that seeded runs reproduce across backends."""
        m, n = self.nvec(), self._n
        if m < 1:
            return
        if self._shard is not None or m * n > Vectors.HOST_RNG_MAX_ELEMENTS:
            # Row-sharded or very large blocks (config 4: 16.8M x 120 would be 16 GB of host
            # RNG + PCIe): counter-based device RNG, seeded from the host stream so that
            # numpy.random.seed() still makes runs reproducible; same seed on every rank,
            # rows keyed globally => independent of the partition.
            seed = int(numpy.random.randint(0, 2 ** 31 - 1))
            self.fill_random_device(seed, row0=self._shard[1] if self._shard is not None else 0)
            return
        data = numpy.random.rand(m, n).astype(self._dtype)
        data *= 2
        data -= 1
        dev.upload_2d(self._wptr(), self._ld * self._w, data)
        self._touch()

    def fill_random_device(self, seed, vector0=0, row0=0):
        """Counter-based device fill (no host traffic): element (j, r) depends only
        on (seed, vector0 + j, row0 + r) -- partition independent."""
        m = self.nvec()
        if m < 1:
            return
        first = self._sel[0]
        self._touch()
        check(lib.rl_fill_uniform(self._code, self._wptr(), self._ld, m, self._n, int(seed),
                                  int(vector0) + first, int(row0), dev.stream()))

    def append(self, other, axis=0):
        if other.nvec() < 1:
            return
        if axis == 1:
            if (self._shard is None) != (other._shard is None):
                raise ValueError('append(axis=1) of a row-sharded and a replicated block')
            m, n = self.nvec(), self._n
            l, n_other = other.nvec(), other._n
            if m != l:
                raise ValueError('Cannot append %d vectors to %d vectors' % (l, m))
            if self.data_type() != other.data_type():
                raise ValueError('Cannot append %s vectors to %s vectors'
                                 % (repr(other.data_type()), repr(self.data_type())))
            n_new = n + n_other
            ld_new = dev.padded_ld(n_new, self._w)
            buf = dev.Buffer(m * ld_new * self._w)
            st = dev.stream()
            check(lib.rl_copy(self._code, buf.ptr, ld_new, self._ptr(self._sel[0]), self._ld, m, n, st))
            check(lib.rl_copy(self._code, buf.ptr + n * self._w, ld_new, other._ptr(other._sel[0]), other._ld, m,
                              n_other, st))
            ng_new = n_new
            if self._shard is not None:
                # every process appends ITS rows: the logical row order of the result is process-major
                # (rows of process 0 of both blocks, then process 1, ...), a contiguous partition again
                ctx = self._shard[0]
                ng_new = self._ng + other._ng
                row0 = self._shard[1] + other._shard[1]
                ctx.register(ng_new, row0, n_new)
                self._shard = (ctx, row0)
            self._buf, self._off, self._ld, self._n, self._ng, self._cap = buf, 0, ld_new, n_new, ng_new, m
            self._nvec = m
            self._sel = (0, m)
            self._touch()
            return
        i, m = self.selected()
        j, l = other.selected()
        if other._n != self._n or other.data_type() != self._dtype:
            raise ValueError('Cannot append incompatible vectors')
        nvec = i + m + l
        st = dev.stream()
        if nvec > self._cap or self._buf is None:
            cap = ((nvec - 1) // self._min_inc + 1) * self._min_inc
            buf = dev.Buffer(cap * self._ld * self._w)
            if i + m > 0:
                check(lib.rl_copy(self._code, buf.ptr, self._ld, self._ptr(0), self._ld, i + m, self._n, st))
            self._buf, self._off, self._cap = buf, 0, cap
            if self._min_inc < Vectors.MAX_INC:
                self._min_inc *= 2
        check(lib.rl_copy(self._code, self._ptr(i + m), self._ld, other._ptr(j), other._ld, l, self._n, st))
        self._touch()
        self._nvec = nvec
        self.select_all()

    def copy(self, other, ind=None):
        i, m = self.selected()
        j, l = other.selected()
        other._touch()
        if ind is None:
            assert m == l
            if m < 1:
                return
            check(lib.rl_copy(self._code, other._ptr(j), other._ld, self._ptr(i), self._ld, m, self._n,
                              dev.stream()))
        else:
            cnt = len(ind)
            if cnt < 1:
                return
            idx = numpy.ascontiguousarray(ind, dtype=numpy.int64)
            # absolute source indices, written at other's selection start (dense_numpy.py:42)
            check(lib.rl_gather(self._code, other._ptr(j), other._ld, self._ptr(0), self._ld,
                                idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cnt, self._n, dev.stream()))

    def _coeffs(self, s, count):
        s = numpy.asarray(s)
        if s.dtype.kind == 'c':
            s = s.real
        s = numpy.ascontiguousarray(s.reshape(-1)[:count], dtype=self._dtype)
        if s.shape[0] < count:
            raise ValueError('coefficient array too short')
        return s

    def scale(self, s, multiply=False):
        f, m = self.selected()
        if m < 1:
            return
        s = self._coeffs(s, m)
        self._touch()
        check(lib.rl_scale_h(self._code, self._wptr(), self._ld, m, self._n, dev.host_ptr(s),
                             1 if multiply else 0, dev.stream()))

    def dots(self, other, transp=False):
        m, n = self.nvec(), self._n
        if transp:
            w = numpy.zeros((n,), dtype=self._dtype)
            if n < 1 or m < 1:
                return w
            out = dev.Buffer(n * self._w)
            check(lib.rl_dots_t(self._code, self._wptr(), self._ld, other._wptr(), other._ld, m, n, out.ptr,
                                dev.stream()))
            check(lib.rl_d2h(dev.host_ptr(w), out.ptr, n * self._w, dev.stream()))
            if self._shard is not None:
                # one entry per COMPONENT: a row-sharded block owns a slice of the result; callers
                # (truncated_svd.py:183-197, lra.py:321) index it with the global dimension
                ctx = self._shard[0]
                w = ctx.allgather_columns(w.reshape(1, -1), ctx.allgather_counts(n)).reshape(-1)
            return w
        w = numpy.zeros((m,), dtype=self._dtype)
        if m < 1:
            return w
        if self._shard is not None:
            wsb = lib.rl_dots_ws_bytes(self._code, m, n)
            ws = dev.Buffer(wsb) if wsb else None
            out = dev.Buffer(m * self._w)
            check(lib.rl_dots(self._code, self._wptr(), self._ld, other._wptr(), other._ld, m, n, out.ptr,
                              ws.ptr if ws else 0, wsb, dev.stream()))
            self._reduce_device(out, m, self._dtype)
            check(lib.rl_d2h(dev.host_ptr(w), out.ptr, m * self._w, dev.stream()))
            return w
        check(lib.rl_dots_h(self._code, self._wptr(), self._ld, other._wptr(), other._ld, m, n,
                            dev.host_ptr(w), dev.stream()))
        return w

    def dot(self, other):
        """G[i, j] = <other_i, self_j>, shape (other.nvec, self.nvec) (dense_numpy.py:78-82)."""
        m, k = self.nvec(), other.nvec()
        g = numpy.zeros((k, m), dtype=self._dtype)
        if m < 1 or k < 1:
            return g
        if max(m, k) > Vectors._GEMM_THRESHOLD:
            return self._reduce_host(self._dot_via_gemm(other))
        if self._shard is not None:
            wsb = lib.rl_gram_ws_bytes(self._code, m, k, self._n)
            ws = dev.Buffer(wsb) if wsb else None
            out = dev.Buffer(k * m * self._w)
            check(lib.rl_gram(self._code, self._wptr(), self._ld, m, other._wptr(), other._ld, k, self._n,
                              out.ptr, ws.ptr if ws else 0, wsb, dev.stream()))
            self._reduce_device(out, k * m, self._dtype)
            check(lib.rl_d2h(dev.host_ptr(g), out.ptr, k * m * self._w, dev.stream()))
            return g
        check(lib.rl_gram_h(self._code, self._wptr(), self._ld, m, other._wptr(), other._ld, k, self._n,
                            dev.host_ptr(g), dev.stream()))
        return g

    _GEMM_THRESHOLD = 2048   # blocks with more vectors than this are data matrices: use the GEMM

    def _dot_via_gemm(self, other):
        # q (k, m) = other (k, n) . self^T  == Matrix(self).apply(other)
        m, k = self.nvec(), other.nvec()
        q = Vectors._local(m, k, self._dtype)
        block_gemm(self._code, self._wptr(), self._ld, m, self._n, other._wptr(), other._ld, q._wptr(), q._ld, k, 0)
        return q.data()

    def _q_host(self, q, rows, cols):
        q = numpy.asarray(q)
        if q.ndim != 2 or q.shape[0] != rows or q.shape[1] != cols:
            raise ValueError('coefficient matrix has shape %s, expected (%d, %d)' % (repr(q.shape), rows, cols))
        if q.dtype.type is not self._dtype:
            q = q.real.astype(self._dtype) if q.dtype.kind == 'c' else q.astype(self._dtype)
        rs, cs = q.strides[0] // q.itemsize, q.strides[1] // q.itemsize
        if q.strides[0] % q.itemsize or q.strides[1] % q.itemsize or rs < 0 or cs < 0:
            q = numpy.ascontiguousarray(q)
            rs, cs = q.strides[0] // q.itemsize, q.strides[1] // q.itemsize
        return q, rs, cs

    def multiply(self, q, output):
        """output <- q^T . self  with q of shape (self.nvec, output.nvec) (dense_numpy.py:84-93)."""
        k = self.nvec()
        m = q.shape[1]
        assert output.nvec() == m
        if m < 1:
            return
        qh, rs, cs = self._q_host(q, k, m)
        output._touch()
        if k > Vectors._GEMM_THRESHOLD:
            # self is a data matrix viewed as vectors (lra.py:236): out (m, n) = q^T (m, k) . S (k, n)
            qt = Vectors(numpy.ascontiguousarray(qh.T))
            block_gemm(self._code, self._wptr(), self._ld, k, self._n, qt._wptr(), qt._ld, output._wptr(),
                       output._ld, m, 1)
            return
        check(lib.rl_update_h(self._code, output._wptr(), output._ld, m, self._wptr(), self._ld, k,
                              dev.host_ptr(qh), rs, cs, 1.0, 0.0, self._n, dev.stream()))

    def add(self, other, s, q=None):
        """Three modes (dense_numpy.py:95-105): scalar s (axpy), scalar s with q
        (self += s q^T other), array s (per-vector axpy)."""
        m = self.nvec()
        if m < 1:
            return
        self._touch()
        if numpy.isscalar(s):
            if q is None:
                check(lib.rl_axpy(self._code, self._wptr(), self._ld, other._wptr(), other._ld, m, self._n,
                                  float(numpy.real(s)), dev.stream()))
            else:
                k = other.nvec()
                if k < 1:
                    return
                qh, rs, cs = self._q_host(q, k, m)
                check(lib.rl_update_h(self._code, self._wptr(), self._ld, m, other._wptr(), other._ld, k,
                                      dev.host_ptr(qh), rs, cs, float(numpy.real(s)), 1.0, self._n,
                                      dev.stream()))
        else:
            sv = self._coeffs(s, m)
            check(lib.rl_axpy_diag_h(self._code, self._wptr(), self._ld, other._wptr(), other._ld, m, self._n,
                                     dev.host_ptr(sv), dev.stream()))

    # ------------------------------------------------------------ other methods
    def shape(self):
        return (self._nvec, self._ng)

    def first(self):
        return self._sel[0]

    def select_all(self):
        self.select(self._nvec)

    def reference(self):
        return Vectors(self, shallow=True)

    def is_complex(self):
        return False

    def conjugate(self):
        return

    def data_size(self):
        return self._w

    def data_ptr(self):
        return self._wptr()

    def all_data_ptr(self):
        return self._ptr(0) if self._buf is not None else 0

    def leading_dimension(self):
        return self._ld

    def zero(self):
        m = self.nvec()
        if m < 1:
            return
        self._touch()
        check(lib.rl_memset(self._wptr(), 0, m * self._ld * self._w, dev.stream()))

    def fill(self, data):
        if isinstance(data, numbers.Number):
            data = numpy.full((self.nvec(), self._n), data, dtype=self._dtype)
        m, n = data.shape
        if m != self.nvec() or (n != self._ng and n != self._n):
            raise ValueError('mismatching dimensions in fill()')
        data = self._local_part(data)
        if m < 1:
            return
        if data.dtype.type is not self._dtype:
            raise ValueError('mismatching data types in fill()')
        self._touch()
        dev.upload_2d(self._wptr(), self._ld * self._w, numpy.ascontiguousarray(data))

    def data(self):
        m = self.nvec()
        if m < 1:
            return numpy.ndarray((m, self._ng), dtype=self._dtype)
        if LAZY_DATA_MIN_BYTES is not None and m * self._n * self._w >= LAZY_DATA_MIN_BYTES:
            return DeviceData(self)                    # inside compat's hooked lra.update only
        if self._shard is None:
            return dev.download_2d(self._wptr(), self._ld * self._w, m, self._n, self._dtype)
        ctx = self._shard[0]
        counts = ctx.allgather_counts(self._n)
        if ctx.on_device:
            return ctx.allgather_columns_device(self._wptr(), self._ld, m, self._n, self._dtype, counts)
        local = dev.download_2d(self._wptr(), self._ld * self._w, m, self._n, self._dtype)
        return ctx.allgather_columns(local, counts)

    def local_data(self):
        """The rows of the selected vectors this process owns, (nvec, nloc)."""
        m = self.nvec()
        if m < 1:
            return numpy.ndarray((m, self._n), dtype=self._dtype)
        return dev.download_2d(self._wptr(), self._ld * self._w, m, self._n, self._dtype)

    def asarray(self):
        return self.data().T

    def orthogonalize(self, other):
        """q = <other, self> (k, m); self -= q^T other; returns q as Vectors
        (k vectors of dimension m).  dense_cublas.py:513-535."""
        m, k, n = self.nvec(), other.nvec(), self._n
        q = Vectors._local(m, k, self._dtype)
        if m < 1 or k < 1:
            return q
        st = dev.stream()
        self._touch()
        if max(m, k) > Vectors._GEMM_THRESHOLD:
            # self is a data chunk: both products are real GEMMs
            block_gemm(self._code, self._wptr(), self._ld, m, n, other._wptr(), other._ld, q._wptr(), q._ld, k, 0)
        else:
            wsb = lib.rl_gram_ws_bytes(self._code, m, k, n)
            ws = dev.Buffer(wsb) if wsb else None
            g = dev.Buffer(k * m * self._w)
            check(lib.rl_gram(self._code, self._wptr(), self._ld, m, other._wptr(), other._ld, k, n, g.ptr,
                              ws.ptr if ws else 0, wsb, st))
            self._reduce_device(g, k * m, self._dtype)
            check(lib.rl_copy(self._code, q._wptr(), q._ld, g.ptr, m, k, m, st))
        check(lib.rl_update(self._code, self._wptr(), self._ld, m, other._wptr(), other._ld, k, q._wptr(),
                            q._ld, 1, -1.0, 1.0, n, st))
        return q

    def svd(self):
        """Thin SVD of the selected block S (m, n): S = v diag(sigma) S_new with
        orthonormal rows S_new overwriting self; returns (sigma, conj(v))
        (dense_numpy.py:125-128; reference GPU: cusolverDn?gesvd, dense_cublas.py:537-591).
        See svd.py for the on-device algorithm."""
        from .svd import block_svd
        return block_svd(self)


# ---- input pipeline for chunked runs ------------------------------------------------
# lra.icompute (lra.py:381-422) hands `matrix[first:next, :]` to AMatrix chunk after chunk; each construction is a
# blocking upload of ~1 GB.  While compat's hooked icompute runs (it is the caller that promises sequential, read-only
# access to the big array) the constructor of chunk i starts a background upload of the rows that FOLLOW it, on a
# side stream, and the constructor of chunk i+1 adopts that buffer if it is asked for exactly those rows.
CHUNK_PREFETCH = False
CHUNK_PREFETCH_MIN_BYTES = 64 << 20


class _ChunkPrefetch:
    def __init__(self):
        self._key = None          # (address, rows, cols, itemsize, ld_bytes)
        self._buf = None
        self._thread = None
        self._stream = None
        self._error = None

    def _join(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def drop(self):
        self._join()
        self._key = self._buf = self._error = None

    def claim(self, a, ld_bytes):
        """Device buffer holding `a` if the pending prefetch is for exactly this block, else None."""
        if self._key is None:
            return None
        key = (a.ctypes.data, a.shape[0], a.shape[1], a.itemsize, ld_bytes)
        self._join()
        buf, ok = self._buf, (self._key == key and self._error is None)
        self._key = self._buf = self._error = None
        return buf if ok else None

    def start_next(self, a, ld_bytes):
        """`a` is rows [r0, r1) of a bigger C-contiguous array: upload rows [r1, r1 + (r1 - r0)) (or what is left)."""
        base = a.base
        if base is None or not isinstance(base, numpy.ndarray) or base.ndim != 2 or not base.flags['C_CONTIGUOUS'] \
                or base.shape[1] != a.shape[1] or base.dtype != a.dtype or a.nbytes < CHUNK_PREFETCH_MIN_BYTES:
            return
        row_bytes = a.shape[1] * a.itemsize
        off = a.ctypes.data - base.ctypes.data
        if off < 0 or off % row_bytes:
            return
        r1 = off // row_bytes + a.shape[0]
        rows = min(a.shape[0], base.shape[0] - r1)
        if rows < 1:
            return
        nxt = base[r1:r1 + rows]
        import threading
        import torch
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        buf = dev.Buffer(rows * ld_bytes)
        # the allocator may hand out memory that work already queued on the main stream still reads
        self._stream.wait_stream(torch.cuda.current_stream())
        stream_handle = self._stream.cuda_stream
        self._key = (nxt.ctypes.data, rows, nxt.shape[1], nxt.itemsize, ld_bytes)
        self._buf = buf
        self._error = None
        device_index = torch.cuda.current_device()

        def work():
            try:
                torch.cuda.set_device(device_index)
                row = nxt.shape[1] * nxt.itemsize
                check(lib.rl_h2d_2d(buf.ptr, ld_bytes, dev.host_ptr(nxt), row, row, rows, stream_handle))
                check(lib.rl_sync_stream(stream_handle))
            except Exception as e:      # the claim falls back to a normal upload
                self._error = e

        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()


_chunk_prefetch = _ChunkPrefetch()


class SampleVectors:
    """The rows of a SAMPLE-PARTITIONED data matrix viewed as vectors (AMatrix.as_vectors(), dense_matrix.py:40-43,
    used by lra.update, lra.py:182-260): every process holds all components of SOME of the vectors.  The object
    answers for the whole set -- nvec() is the global count, per-vector results are concatenated over the
    processes (process-major order, the order of the logical matrix), sums over the vectors are all-reduced --
    so that the reference's update code, which runs replicated on every process, sees one consistent matrix.
    Only what that code uses is provided; selections are not."""
    _rl_device_block = True

    def __init__(self, matrix, shallow=True):
        ctx, row0 = matrix._mshard
        _SAMPLE_LOCAL[0] = True
        try:
            self._loc = Vectors(matrix, shallow=shallow)         # this process's rows, an ordinary local block
        finally:
            _SAMPLE_LOCAL[0] = False
        self._ctx, self._row0 = ctx, int(row0)
        self._counts = ctx.allgather_counts(self._loc.nvec())
        self._nvg = int(sum(self._counts))
        assert self._row0 == sum(self._counts[:ctx.rank])

    # -- bookkeeping
    def nvec(self):
        return self._nvg

    def dimension(self):
        return self._loc.dimension()

    def data_type(self):
        return self._loc.data_type()

    def is_complex(self):
        return False

    def selected(self):
        return (0, self._nvg)

    def select(self, nv, first=0):
        if first != 0 or nv != self._nvg:
            raise NotImplementedError('selection of sample-partitioned vectors')

    def select_all(self):
        pass

    def is_sharded(self):
        return False

    def local(self):
        return self._loc

    def new_vectors(self, arg=0, dim=None):
        return self._loc.new_vectors(arg, dim)

    def _rows(self, a, axis):
        """This process's part of a host array indexed by the (global) vectors along `axis`."""
        a = numpy.asarray(a)
        n = self._loc.nvec()
        if a.shape[axis] == self._nvg:
            return a[self._row0:self._row0 + n] if axis == 0 else a[:, self._row0:self._row0 + n]
        if a.shape[axis] == n:
            return a
        raise ValueError('array extent %d matches neither the global (%d) nor the local (%d) number of vectors'
                         % (a.shape[axis], self._nvg, n))

    # -- algebra
    def dots(self, other, transp=False):
        if transp:
            w = self._loc.dots(other.local() if isinstance(other, SampleVectors) else other, transp=True)
            return self._ctx.allreduce_host(w)
        o = other.local() if isinstance(other, SampleVectors) else other
        w = self._loc.dots(o)
        return self._ctx.allgather_columns(w[None, :], self._counts)[0]

    def multiply(self, q, output):
        """output = q^T self with q (nvec, k): the sum runs over the vectors, i.e. over the processes."""
        q = numpy.asarray(q)
        if q.ndim == 1:
            q = q.reshape(-1, 1)
        self._loc.multiply(numpy.ascontiguousarray(self._rows(q, 0)), output)
        _allreduce_block(self._ctx, output)

    def add(self, other, s, q=None):
        """self += s q^T other, q (other.nvec, nvec): row-local once q is cut to this process's vectors."""
        if q is None:
            o = other.local() if isinstance(other, SampleVectors) else other
            return self._loc.add(o, s)
        self._loc.add(other, s, numpy.ascontiguousarray(self._rows(q, 1)))

    def scale(self, s, multiply=False):
        self._loc.scale(numpy.ascontiguousarray(self._rows(numpy.asarray(s).reshape(-1), 0)), multiply)

    def orthogonalize(self, other):
        """q = <other, self>, self -= q^T other; q comes back as k vectors whose dimension is the (global) number of
        vectors of self -- a row-sharded block like every other block over the samples."""
        q = self._loc.orthogonalize(other)
        q._ng = self._nvg
        q._shard = (self._ctx, self._row0)
        return q

    def data(self):
        local = self._loc.data()
        return numpy.ascontiguousarray(self._ctx.allgather_columns(numpy.ascontiguousarray(local.T), self._counts).T)


def _allreduce_block(ctx, v):
    """Sum the selected (replicated-dimension) block of `v` over the processes, in place."""
    import torch
    m = v.nvec()
    if m < 1 or ctx.world == 1:
        return
    host = v.data()
    v.fill(ctx.allreduce_host(host))


class Matrix:
    """Dense operator holder.  dense_cublas.py:635-776."""

    def __init__(self, arg):
        self._mshard = None                # (ctx, row0) when this process holds a row slab
        if isinstance(arg, DeviceData):
            arg = numpy.asarray(arg)
        if isinstance(arg, Vectors):
            f, m = arg.selected()
            if arg.is_sharded():
                raise NotImplementedError('Matrix over row-sharded vectors')
            self._shape = (m, arg.dimension())
            self._dtype = arg.data_type()
            self._order = 'C_CONTIGUOUS'
            ref = Vectors(arg, shallow=True)
            self._buf, self._ld, self._base = ref._buf, ref._ld, ref._off
        elif isinstance(arg, numpy.ndarray):
            if arg.ndim != 2:
                raise ValueError('Matrix data must be a 2D array')
            self._shape = arg.shape
            self._dtype = _np_type(arg.dtype.type)
            if arg.flags['C_CONTIGUOUS']:
                self._order = 'C_CONTIGUOUS'
                stored = arg
            elif arg.flags['F_CONTIGUOUS']:
                self._order = 'F_CONTIGUOUS'
                stored = arg.T          # device keeps the (N, M) C-ordered transpose
            else:
                raise ValueError('Matrix data must be either C- or F-contiguous')
            if self._dtype not in (numpy.float32, numpy.float64):
                raise ValueError('data type %s not supported' % repr(self._dtype))
            rows, cols = stored.shape
            w = stored.itemsize
            self._ld = dev.padded_ld(cols, w)
            self._base = 0
            got = _chunk_prefetch.claim(stored, self._ld * w) if CHUNK_PREFETCH else None
            if got is not None:
                self._buf = got                               # uploaded in the background during the previous chunk
            else:
                self._buf = dev.Buffer(max(rows, 1) * self._ld * w)
                dev.upload_2d(self._buf.ptr, self._ld * w, stored)
            if CHUNK_PREFETCH:
                _chunk_prefetch.start_next(stored, self._ld * w)
            if arg.nbytes >= MINMAX_ON_DEVICE_BYTES:
                # AMatrix scans this same host array with numpy.amin / amax right after building
                # the Matrix (dense_matrix.py:32-34): let compat's proxy answer from the device copy
                _recent_uploads.clear()
                _recent_uploads[id(arg)] = (weakref.ref(self), arg.shape, arg.dtype)
            from . import dist
            ctx = dist.current()
            if ctx is not None and ctx.shard_matrices and ctx.world > 1:
                # sample-partitioned data matrix: `arg` holds this
                # rank's rows; the logical shape is the concatenation over ranks
                if self._order != 'C_CONTIGUOUS':
                    raise ValueError('a row-sharded Matrix must be C_CONTIGUOUS')
                counts = ctx.allgather_counts(rows)
                row0, total = sum(counts[:ctx.rank]), sum(counts)
                if total == cols:
                    raise ValueError('square global matrix: sharded and replicated dimensions coincide')
                ctx.register(total, row0, rows)
                self._mshard = (ctx, row0)
                self._shape = (total, cols)
        else:
            raise ValueError('wrong argument %s in Matrix constructor' % repr(type(arg)))
        self._code = _lib.dtype_code(self._dtype)
        self._w = 4 if self._dtype is numpy.float32 else 8
        self._lo = None               # low part of the 3xTF32 split (tensor-core path)
        self._lo_version = -1

    TC_MIN_VECTORS = TC_MIN_VECTORS
    MATERIALISED_LO = False           # True: keep a - tf32(a) in HBM (r1 scheme, twice the traffic; measurements only)

    def minmax(self):
        """(min, max) over the stored block in one HBM-bound pass on the device."""
        rows, cols = self._stored_shape()
        lo, hi = numpy.zeros(1, self._dtype), numpy.zeros(1, self._dtype)
        check(lib.rl_minmax_h(self._code, self._aptr(), self._ld, rows, cols, dev.host_ptr(lo), dev.host_ptr(hi),
                              dev.stream()))
        return lo[0], hi[0]

    def _local_shape(self):
        """(rows, cols) of the part of the logical matrix held by this process."""
        m, n = self._shape
        if self._mshard is not None:
            m = self._mshard[0].lookup(m)[1]
        return (m, n)

    def _stored_shape(self):
        m, n = self._local_shape()
        return (m, n) if self._order == 'C_CONTIGUOUS' else (n, m)

    def _lo_ptr(self):
        """Device copy of a - tf32(a), refreshed whenever the matrix memory has
        been written since it was built (vectors aliasing the matrix bump the
        buffer version, e.g. lra.update centring a chunk in place)."""
        rows, cols = self._stored_shape()
        if self._lo is None:
            self._lo = dev.Buffer(max(rows, 1) * self._ld * 4)
        if self._lo_version != self._buf.version:
            check(lib.rl_split_tf32(self._aptr(), self._ld, self._lo.ptr, self._ld, rows, cols, dev.stream()))
            self._lo_version = self._buf.version
        return self._lo.ptr

    def _aptr(self):
        return self._buf.ptr + self._base * self._ld * self._w

    def data_ptr(self):
        return self._aptr()

    def order(self):
        return self._order

    def shape(self):
        return self._shape

    def data_type(self):
        return self._dtype

    def data_size(self):
        return self._w

    def is_complex(self):
        return False

    def fill(self, data):
        stored = data if self._order == 'C_CONTIGUOUS' else data.T
        self._buf.version += 1
        dev.upload_2d(self._aptr(), self._ld * self._w, numpy.ascontiguousarray(stored, dtype=self._dtype))

    def dots(self):
        """Squared 2-norms of the rows, one per row of the LOGICAL matrix (dense_numpy.py:177-179)."""
        v = Vectors(self, shallow=True)     # row slab of a sharded matrix: SampleVectors, results concatenated
        return v.dots(v)

    def new_vectors(self, dim=None, nv=0):
        if dim is None:
            dim = self.shape()[1]
        return Vectors(dim, nv, self.data_type())

    def apply(self, x, y, transp=False):
        """y = x . A^T, or y = x . A when transp (dense_cublas.py:732-776)."""
        if x.data_type() != self._dtype or y.data_type() != self._dtype:
            raise ValueError('Matrix and vectors data types differ')
        m, n = self._shape
        if transp:
            if n != y.dimension() or m != x.dimension():
                raise ValueError('Matrix and vectors dimensions incompatible')
        else:
            if m != y.dimension() or n != x.dimension():
                raise ValueError('Matrix and vectors dimensions incompatible')
        k = x.nvec()
        if k != y.nvec():
            raise ValueError('Numbers of input and output vectors differ')
        if k < 1:
            return
        m, n = self._local_shape()
        if self._mshard is not None:
            xs, ys = (x, y) if transp else (y, x)      # xs lives on the sharded (row) dimension
            if not xs.is_sharded() or xs.local_dimension() != m or ys.is_sharded():
                raise ValueError('vectors are not laid out like the row-sharded matrix')
        if self._order == 'C_CONTIGUOUS':
            M, N, t = m, n, 1 if transp else 0
        else:   # stored transposed: A = B^T with B (n, m) row-major
            M, N, t = n, m, 0 if transp else 1
        y._touch()
        self._apply_local(x, y, k, M, N, t)
        if self._mshard is not None and transp:
            # partial products of the row slabs -> replicated result (NCCL all-reduce over NVLink)
            import torch
            tdt = torch.float32 if self._dtype is numpy.float32 else torch.float64
            first = (y._off + y._sel[0]) * y._ld * y._w
            view = y._buf.tensor[first:first + k * y._ld * y._w].view(tdt)
            self._mshard[0].allreduce_(view)

    def _apply_local(self, x, y, k, M, N, t):
        if Matrix.MATERIALISED_LO and self._dtype is numpy.float32 and k >= TC_MIN_VECTORS and \
                lib.rl_dense_apply_tc_supported(self._aptr(), self._ld, x._wptr(), x._ld):
            wsb = lib.rl_dense_apply_tc_ws_bytes(M, N, k, t)      # A/B variant: lo copy of the matrix in HBM
            ws = dev.Buffer(wsb)
            check(lib.rl_dense_apply_tc(self._aptr(), self._lo_ptr(), self._ld, M, N, x._wptr(), x._ld,
                                        y._wptr(), y._ld, k, t, 1.0, 0.0, ws.ptr, wsb, dev.stream()))
            return
        block_gemm(self._code, self._aptr(), self._ld, M, N, x._wptr(), x._ld, y._wptr(), y._ld, k, t)
