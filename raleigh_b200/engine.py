"""Device engine of the block Jacobi-CG driver (jcg.py): every operation is one
call into libraleigh_b200.so (include/raleigh_b200.h, section "device-resident
Rayleigh-Ritz"); all small matrices are fp64 views into ONE device workspace
that lives for the whole solve, so that Gram results flow into the pivoted
Cholesky / Rayleigh-Ritz kernels and the resulting coefficients into the block
updates without touching the host.  The only D2H traffic is three packets of a
few hundred bytes per iteration (fetch_ritz, fetch_chol, fetch_estimates).

Row-sharded blocks (dist.py): Gram matrices and dot products are summed over
the ranks with an NCCL all-reduce of the device result; everything small is
then computed redundantly (and identically) on every rank.
"""
import ctypes

import numpy

from . import _lib
from ._lib import lib, check
from . import device as dev


class DSmall:
    """Row-major fp64 matrix view in device memory."""
    __slots__ = ('ptr', 'ld', 'rows', 'cols')

    def __init__(self, ptr, ld, rows, cols):
        self.ptr, self.ld, self.rows, self.cols = ptr, ld, rows, cols

    def sub(self, r0, c0, nr, nc):
        return DSmall(self.ptr + (int(r0) * self.ld + int(c0)) * 8, self.ld, nr, nc)


class _Arena:
    def __init__(self, nbytes):
        self.buf = dev.Buffer(nbytes, zero=True)
        self.off = 0

    def take(self, nbytes):
        nbytes = (int(nbytes) + 255) & ~255
        if self.off + nbytes > self.buf.nbytes:
            raise MemoryError('device workspace of the Rayleigh-Ritz engine exhausted')
        p = self.buf.ptr + self.off
        self.off += nbytes
        return p

    def matrix(self, rows, cols):
        return DSmall(self.take(rows * cols * 8), cols, rows, cols)


class DeviceEngine:
    def begin(self, vector, m):
        self.m = m
        self._template = vector
        self._code = vector._code
        # Jacobi stopping tolerance of the Rayleigh-Ritz eigenproblems: working precision for fp64 data; for
        # fp32 data the Gram matrices themselves carry 1e-7 relative noise
        self._eig_tol = 1e-9 if vector._code == _lib.RL_F32 else 0.0
        M = 2 * m
        self.M = M
        rr_ws = lib.rl_rr_solve_ws_bytes(M)
        total = (6 * M * M + 6 * m * m + 3 * M * m + 16 * M) * 8 + rr_ws + (64 << 10)
        ar = self._arena = _Arena(total)
        self.GB, self.GA = ar.matrix(M, M), ar.matrix(M, M)
        self._A0 = ar.matrix(M, M)
        self.XAX, self.XBX = ar.matrix(m, m), ar.matrix(m, m)
        self.ZAY, self.ZBY, self.Beta, self.T1 = ar.matrix(m, m), ar.matrix(m, m), ar.matrix(m, m), ar.matrix(m, m)
        self.CX, self.CZ = ar.matrix(M, m), ar.matrix(M, M)
        # one packet per host fetch: [stats(8) | lmd (M) | s2 (M)], [info(4 ints)+ind (M ints)], [dX (M) | dlmd (M)]
        ritz = ar.take((8 + 2 * M) * 8)
        self._ritz_ptr = ritz
        self.stats = DSmall(ritz, 8, 1, 8)
        self.v_lmd = DSmall(ritz + 8 * 8, M, 1, M)
        self.v_s2 = DSmall(ritz + (8 + M) * 8, M, 1, M)
        self.v_t2 = ar.matrix(1, M)
        self.v_y2 = ar.matrix(1, M)          # squared norms of the search directions (conjugation)
        self._chol_ptr = ar.take((8 + M) * 4)
        self._est_ptr = ar.take(2 * M * 8)
        self._lmdx = ar.take(M * 8)
        self._lmdz = ar.take(M * 8)
        self._rr_ws = ar.take(rr_ws)
        self._rr_ws_bytes = rr_ws
        self._eig_info = ar.take(64)
        self._h_ritz = numpy.zeros(8 + 2 * M, dtype=numpy.float64)
        self._h_chol = numpy.zeros(8 + M, dtype=numpy.int32)
        self._h_est = numpy.zeros(2 * M, dtype=numpy.float64)
        self.Gc = self.TC = self.QC = None
        self._ccap = 0
        self._cbuf = None

    # ---- storage ----------------------------------------------------------------------
    def new_block(self):
        return self._template.new_vectors(self.m)

    def reserve_constraints(self, cap):
        """Device Gram matrix of the locked vectors, grown geometrically."""
        if cap <= self._ccap:
            return
        new = max(int(cap), 2 * self._ccap, 4 * self.m)
        M = self.M
        buf = dev.Buffer((new * new + 2 * new * M) * 8 + 1024, zero=True)
        Gc = DSmall(buf.ptr, new, new, new)
        if self.Gc is not None and self._ccap > 0:
            check(lib.rl_small_copy(self.Gc.ptr, self.Gc.ld, Gc.ptr, Gc.ld, self._ccap, self._ccap, dev.stream()))
        self.Gc = Gc
        self.TC = DSmall(buf.ptr + new * new * 8, M, new, M)
        self.QC = DSmall(buf.ptr + (new * new + new * M) * 8, M, new, M)
        self._cbuf, self._ccap = buf, new          # the old buffer is released after the copy (same stream)

    # ---- block <-> small ----------------------------------------------------------------
    def _reduce(self, S, view, rows, cols):
        """Sum a device result over the ranks of a row-sharded block."""
        if S._shard is None:
            return
        import torch
        ctx = S._shard[0]
        if view.ld == cols:
            t = self._as_tensor(view.ptr, rows * cols)
            ctx.allreduce_(t)
            return
        tmp = torch.empty(rows * cols, dtype=torch.float64, device='cuda')
        check(lib.rl_small_copy(view.ptr, view.ld, tmp.data_ptr(), cols, rows, cols, dev.stream()))
        ctx.allreduce_(tmp)
        check(lib.rl_small_copy(tmp.data_ptr(), cols, view.ptr, view.ld, rows, cols, dev.stream()))

    def _as_tensor(self, ptr, count):
        """torch view of `count` doubles of the engine's own workspace."""
        import torch
        for buf in (self._arena.buf, self._cbuf):
            if buf is not None and buf.ptr <= ptr < buf.ptr + buf.nbytes:
                off = ptr - buf.ptr
                return buf.tensor[off:off + count * 8].view(torch.float64)
        raise ValueError('pointer outside the engine workspace')

    def gram(self, S, O, out):
        m, k = S.nvec(), O.nvec()
        if m < 1 or k < 1:
            return
        check(lib.rl_gram_dev(self._code, S._wptr(), S._ld, m, O._wptr(), O._ld, k, S._n, out.ptr, out.ld,
                              dev.stream()))
        self._reduce(S, out, k, m)

    def dots(self, S, O, vec):
        m = S.nvec()
        if m < 1:
            return
        check(lib.rl_dots_dev(self._code, S._wptr(), S._ld, O._wptr(), O._ld, m, S._n, vec.ptr, dev.stream()))
        self._reduce(S, vec, 1, m)

    def update(self, out, X, q, alpha, beta):
        m, k = out.nvec(), X.nvec()
        if m < 1:
            return
        out._touch()
        check(lib.rl_update_dev(self._code, out._wptr(), out._ld, m, X._wptr(), X._ld, k, q.ptr, q.ld,
                                float(alpha), float(beta), out._n, dev.stream()))

    def residual(self, W, AX, X, vec):
        m = X.nvec()
        if m < 1:
            return
        W._touch()
        check(lib.rl_residual_dev(self._code, W._wptr(), W._ld, AX._wptr(), AX._ld, X._wptr(), X._ld, m, X._n,
                                  vec.ptr, dev.stream()))

    def scale_rsqrt(self, Y, vec):
        m = Y.nvec()
        if m < 1:
            return
        Y._touch()
        check(lib.rl_scale_rsqrt_dev(self._code, Y._wptr(), Y._ld, m, Y._n, vec.ptr, dev.stream()))

    def gather(self, src, idx, dst):
        """dst[t] = src[selection start + idx[t]]."""
        cnt = len(idx)
        if cnt < 1:
            return
        dst._touch()
        ind = numpy.ascontiguousarray(idx, dtype=numpy.int64)
        check(lib.rl_gather(self._code, dst._wptr(), dst._ld, src._wptr(), src._ld,
                            ind.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cnt, src._n, dev.stream()))

    # ---- small matrices -------------------------------------------------------------------
    def copy_small(self, src, dst):
        check(lib.rl_small_copy(src.ptr, src.ld, dst.ptr, dst.ld, src.rows, src.cols, dev.stream()))

    def mirror_upper(self, G, nx, ny):
        check(lib.rl_small_mirror(G.ptr, G.ld, nx, ny, dev.stream()))

    def ritz_check(self, nx):
        check(lib.rl_rr_ritz_check(self.XAX.ptr, self.XBX.ptr, self.XAX.ld, nx, self._lmdx, self.v_lmd.ptr,
                                   self.stats.ptr, dev.stream()))

    def fetch_ritz(self, nx):
        h, M = self._h_ritz, self.M
        check(lib.rl_d2h(dev.host_ptr(h), self._ritz_ptr, (8 + 2 * M) * 8, dev.stream()))
        return h[8:8 + nx].copy(), h[8 + M:8 + M + nx].copy(), float(h[0]), float(h[1])

    def constraint_coeffs(self, nc, nv):
        """QC = 2 TC - Gc TC: the approximate inverse 2I - Gc of the locked vectors' Gram matrix
        applied to the projections (solver.py:757, 774, 1270)."""
        st = dev.stream()
        check(lib.rl_small_copy(self.TC.ptr, self.TC.ld, self.QC.ptr, self.QC.ld, nc, nv, st))
        check(lib.rl_small_gemm(0, 0, nc, nv, nc, -1.0, self.Gc.ptr, self.Gc.ld, self.TC.ptr, self.TC.ld, 2.0,
                                self.QC.ptr, self.QC.ld, st))

    def conjugation(self, nz, ny):
        check(lib.rl_rr_conjugation(self.ZAY.ptr, self.ZBY.ptr, self.Beta.ptr, self.ZAY.ld, nz, ny, self.v_lmd.ptr,
                                    self._lmdz, self.v_y2.ptr, self.v_t2.ptr, dev.stream()))

    def piv_chol(self, G, n, k, eps):
        check(lib.rl_rr_piv_chol(G.ptr, self._A0.ptr, G.ld, n, k, float(eps), self._chol_ptr + 32, self._chol_ptr,
                                 dev.stream()))

    def fetch_chol(self, n):
        h = self._h_chol
        check(lib.rl_d2h(dev.host_ptr(h), self._chol_ptr, (8 + n) * 4, dev.stream()))
        if h[1] != 0:
            raise RuntimeError('Gram matrix of the iterates is not positive definite')
        return int(h[0]), h[8:8 + n].astype(numpy.int64)

    def ritz_initial(self, n):
        """Generalised n x n problem GA q = lambda GB q of the initial block (solver.py:822)."""
        self.piv_chol(self.GB, n, n, 0.0)
        self._rr(n, 0, n, 0, n, 0)

    def rayleigh_ritz(self, nx, ny, leftX, rightX, leftXn, rightXn):
        self._rr(nx, ny, leftX, rightX, leftXn, rightXn)

    def _rr(self, nx, ny, leftX, rightX, leftXn, rightXn):
        check(lib.rl_rr_solve(self.GA.ptr, self.GB.ptr, self.GA.ld, nx, ny, leftX, rightX, leftXn, rightXn,
                              self.CX.ptr, self.CX.ld, self.CZ.ptr, self.CZ.ld, self._lmdx, self._lmdz,
                              self._est_ptr, self.M, self._eig_tol, self._rr_ws, self._rr_ws_bytes, self._eig_info,
                              dev.stream()))

    def fetch_estimates(self, nx):
        h, M = self._h_est, self.M
        check(lib.rl_d2h(dev.host_ptr(h), self._est_ptr, 2 * M * 8, dev.stream()))
        return h[:nx].copy(), h[M:M + nx].copy()
