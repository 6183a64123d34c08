"""NumPy engine for the device-resident block-CG driver (TEST INFRASTRUCTURE ONLY).

raleigh_b200/jcg.py drives an `engine` object; the product engine
(raleigh_b200/engine.py) is ctypes over the CUDA library.  This module is the
CPU twin used by tests/ to check (1) the driver's control flow against the
reference's own solver on the CPU and (2) every small-matrix kernel of
csrc/rr.cu / csrc/jacobi.cu against a NumPy statement of the SAME algorithm
(right-looking pivoted Cholesky with the drop rule of solver.py:1749-1826,
shifted one-sided Jacobi, ...).  Never imported by raleigh_b200.

Blocks are oracle.host_backend.Vectors; small matrices are float64 ndarrays
wrapped in `Small` so that `.sub(r0, c0, nr, nc)` is a writable view.
"""
import numpy as np
import scipy.linalg as sla



class Small:
    def __init__(self, a):
        self.a = a

    def sub(self, r0, c0, nr, nc):
        return Small(self.a[r0:r0 + nr, c0:c0 + nc])


# --------------------------------------------------------------------------- kernels
def piv_chol(A, n, k, eps):
    """In-place factorisation of the leading n x n block of the symmetric matrix A:
    unpivoted Cholesky on the first k rows/columns, max-diagonal pivoting on the rest,
    U^T U = P^T A P with U upper triangular.  Pivots <= eps end the factorisation
    (everything from there on is dropped); the condition estimate lmin/lmax <= eps of
    the leading factor is tested every 64 pivoted columns and at the end, and after a
    pivot drop the largest well-conditioned leading block is found by bisection --
    the rule of solver.py:1749-1826, restated right-looking (the trailing matrix is
    updated after every column).  Returns (ind, dropped, status)."""
    A0 = A[:n, :n].copy()
    ind = np.arange(n)
    dropped = 0
    drop_case = 0
    last_check = -1
    blk = 64
    l = k
    status = 0
    for i in range(n):
        if i >= k:
            d = np.diag(A)[i:n]
            j = i + int(np.argmax(d))
            if j != i:
                A[[i, j], :n] = A[[j, i], :n]
                A[:n, [i, j]] = A[:n, [j, i]]
                ind[[i, j]] = ind[[j, i]]
        piv = A[i, i]
        if i >= k and piv <= eps:
            A[i:n, :n] = 0.0
            drop_case = 1
            dropped = n - i
            break
        if piv <= 0.0:
            status = 1          # leading block not positive definite (the reference's LAPACK call raises)
            A[i:n, :n] = 0.0
            dropped = n - i
            drop_case = 2
            break
        r = np.sqrt(piv)
        A[i, i] = r
        A[i, i + 1:n] /= r
        A[i + 1:n, i] = 0.0
        row = A[i, i + 1:n]
        A[i + 1:n, i + 1:n] -= np.outer(row, row)
        if i >= k and (i - l == blk - 1 or i == n - 1):
            last_check = i
            if _cond_inverse(A, A0, ind, i + 1) <= eps:
                A[i:n, :n] = 0.0
                drop_case = 2
                dropped = n - i
                break
            if i - l == blk - 1:
                l += blk
    if last_check < n - 1 and drop_case == 1:
        i = last_check
        j = n - dropped - 1
        while i < j:
            mid = i + (j - i + 1) // 2
            if _cond_inverse(A, A0, ind, mid + 1) <= eps:
                if j > mid:
                    j = mid
                    continue
                A[j:n, :n] = 0.0
                dropped = n - j
                break
            i = mid
    return ind, dropped, status


def _cond_inverse(U, A0, ind, p):
    """lmin / lmax estimate of the leading p x p block: lmax = 1-norm of the (permuted)
    matrix itself (= U^T U, solver.py:1828-1830), lmin from three steps of inverse
    iteration started from the vector of ones (solver.py:1831-1845)."""
    G = A0[np.ix_(ind[:p], ind[:p])]
    lmax = np.max(np.sum(np.abs(G), axis=0))
    T = np.triu(U[:p, :p])
    x = np.ones(p)
    s = float(x @ x)
    rq = 0.0
    for _ in range(3):
        y = sla.solve_triangular(T, x, trans=1)
        t = float(y @ y)
        rq = s / t
        x = sla.solve_triangular(T, y)
        s = float(x @ x)
    return rq / lmax


def jacobi_layout(n):
    """(cluster size C, columns per block W) chosen by rl_syevj_cluster for order n."""
    npairs = (n + 1) // 2
    length = (n + 63) // 64 * 64
    c = 1
    while True:
        w = (npairs + c - 1) // c
        if (w <= 16 or (c == 8 and w <= 32)) and 4 * w * length * 8 <= 200 * 1024:
            return c, max(w, 1)
        if c == 8:
            raise ValueError('order %d too large for the cluster kernel' % n)
        c *= 2


def _rotate(B, p, q, tol2):
    bp, bq = B[:, p].copy(), B[:, q].copy()
    alpha, beta, gamma = bp @ bp, bq @ bq, bp @ bq
    if not gamma * gamma > tol2 * alpha * beta:
        return False
    delta = beta - alpha
    h = delta * delta + 4.0 * gamma * gamma
    r = 1.0 / np.sqrt(h)                       # cos 2theta = |delta| r, |theta| <= pi/4 (no division, no tangent:
    c2 = 0.5 + 0.5 * abs(delta) * r            # the form csrc/jacobi.cu uses)
    c = np.sqrt(c2)
    s = (gamma if delta >= 0 else -gamma) * r / c
    B[:, p], B[:, q] = c * bp - s * bq, s * bp + c * bq
    return True


def jacobi_eigh(G, max_sweeps=48):
    """Symmetric eigen-decomposition by ONE-SIDED Jacobi on the shifted matrix
    B = G + sigma I (sigma from Gershgorin discs, so that B is positive definite):
    column rotations orthogonalise B V; then B V = Q diag(lambda + sigma).  Same
    algorithm and the same two-level rotation order as csrc/jacobi.cu (blocks of W
    columns, two per CTA of a cluster of C): per sweep first the pairs inside every
    block, then 2C-1 outer rounds of block-against-block pairs with the blocks moving
    round-robin between the CTAs.  Returns (w ascending, Q columns, sweeps)."""
    n = G.shape[0]
    if n == 0:
        return np.zeros(0), np.zeros((0, 0)), 0
    G = 0.5 * (G + G.T)
    d = np.diag(G)
    radius = np.sum(np.abs(G), axis=1) - np.abs(d)
    norm = float(np.max(np.abs(d) + radius))
    low = float(np.min(d - radius))
    sigma = 0.0
    if low <= 1e-3 * norm:
        sigma = -low + 1e-2 * norm
    if not norm > 0.0:
        sigma = 1.0
    C, W = jacobi_layout(n)
    ncols = 2 * C * W
    B = np.zeros((n, ncols))
    B[:, :n] = G + sigma * np.eye(n)
    tol = np.sqrt(n) * np.finfo(np.float64).eps
    tol2 = tol * tol
    # block b of CTA c: slot (c, side); blocks[c][side] = list of column ids
    blocks = [[list(range((2 * c + s) * W, (2 * c + s + 1) * W)) for s in range(2)] for c in range(C)]
    P = (W + 1) & ~1
    sweeps = 0
    for sweep in range(max_sweeps):
        rotated = False
        for t in range(P - 1):
            for c in range(C):
                for side in range(2):
                    cols = blocks[c][side]
                    for i in range(P // 2):
                        if i == 0:
                            a, b = P - 1, t
                        else:
                            a, b = (t + i) % (P - 1), (t - i + P - 1) % (P - 1)
                        if a > b:
                            a, b = b, a
                        if b >= W:
                            continue
                        rotated |= _rotate(B, cols[a], cols[b], tol2)
        for o in range(2 * C - 1 if C > 1 else 1):
            for r in range(W):
                for c in range(C):
                    A_, B_ = blocks[c]
                    for w in range(W):
                        rotated |= _rotate(B, A_[w], B_[(w + r) % W], tol2)
            if C > 1:
                new = [[None, None] for _ in range(C)]
                for c in range(C):
                    if c == 0:
                        new[0][0] = blocks[0][0]
                        new[1][0] = blocks[0][1]
                    else:
                        if c == C - 1:
                            new[c][1] = blocks[c][0]
                        else:
                            new[c + 1][0] = blocks[c][0]
                        new[c - 1][1] = blocks[c][1]
                blocks = new
        sweeps = sweep + 1
        if not rotated:
            break
    norms = np.sqrt(np.sum(B * B, axis=0))
    real = norms > 0
    w = norms[real] - sigma
    Q = B[:, real] / norms[real][None, :]
    order = np.argsort(w, kind='stable')
    return w[order], Q[:, order], sweeps


def transform(GA, U):
    """G = U^-T GA U^-1 (solver.py:1685-1688) by two forward substitutions."""
    B = sla.solve_triangular(U.T, GA.T, lower=True)
    return sla.solve_triangular(U.T, B.T, lower=True)


class NumpyEngine:
    """Same surface as raleigh_b200.engine.DeviceEngine."""

    def __init__(self, eigh='lapack'):
        self._eigh_kind = eigh
        self.sweeps = []

    def _eigh(self, G):
        if self._eigh_kind == 'lapack':
            return sla.eigh(G)
        w, Q, sweeps = jacobi_eigh(np.array(G))
        self.sweeps.append((G.shape[0], sweeps))
        return w, Q

    # ---- set-up
    def begin(self, vector, m):
        self.m = m
        self._template = vector
        z = lambda r, c: Small(np.zeros((r, c)))     # noqa: E731
        M = 2 * m
        self.GB, self.GA = z(M, M), z(M, M)
        self.XAX, self.XBX = z(m, m), z(m, m)
        self.ZAY, self.ZBY, self.Beta, self.T1 = z(m, m), z(m, m), z(m, m), z(m, m)
        self.CX, self.CZ = z(M, m), z(M, M)
        self.v_s2, self.v_t2, self.v_lmd = z(1, M), z(1, M), z(1, M)
        self.v_y2 = z(1, M)
        self.lmdx, self.lmdz = np.zeros(M), np.zeros(M)
        self.stats = np.zeros(4)
        self.Gc = self.TC = self.QC = None
        self._ccap = 0
        self._chol = None
        self._est = None

    def new_block(self):
        return self._template.new_vectors(self.m)        # works for any host backend with the reference interface

    def reserve_constraints(self, cap):
        if cap <= self._ccap:
            return
        new = max(cap, 2 * self._ccap, 64)
        Gc = np.zeros((new, new))
        if self.Gc is not None:
            Gc[:self._ccap, :self._ccap] = self.Gc.a
        self.Gc = Small(Gc)
        self.TC, self.QC = Small(np.zeros((new, 2 * self.m))), Small(np.zeros((new, 2 * self.m)))
        self._ccap = new

    # ---- block <-> small
    def gram(self, S, O, out):
        out.a[...] = O.data().astype(np.float64) @ S.data().astype(np.float64).T

    def dots(self, S, O, vec):
        k = S.nvec()
        vec.a[0, :k] = np.sum(S.data().astype(np.float64) * O.data().astype(np.float64), axis=1)

    def update(self, out, X, q, alpha, beta):
        dt = out.data_type()
        new = alpha * (q.a.astype(dt).T @ X.data())
        if beta == 0.0:
            out.data()[...] = new
        else:
            out.data()[...] = beta * out.data() + new

    def residual(self, W, AX, X, vec):
        k = X.nvec()
        W.data()[...] = AX.data() - vec.a[0, :k].astype(X.data_type())[:, None] * X.data()

    def scale_rsqrt(self, Y, vec):
        k = Y.nvec()
        s = np.sqrt(np.abs(vec.a[0, :k]))
        nz = s != 0
        Y.data()[nz, :] = (Y.data()[nz, :] / s[nz].astype(Y.data_type())[:, None])

    def gather(self, src, idx, dst):
        f = src.selected()[0]
        dst.data()[...] = src.all_data()[f + np.asarray(idx, dtype=np.int64), :]

    # ---- small matrices
    def copy_small(self, src, dst):
        dst.a[...] = src.a

    def mirror_upper(self, G, nx, ny):
        G.a[nx:nx + ny, :nx] = G.a[:nx, nx:nx + ny].T

    def ritz_check(self, nx):
        a, b = self.XAX.a[:nx, :nx], self.XBX.a[:nx, :nx]
        lmd = np.diag(a) / np.diag(b)
        self.v_lmd.a[0, :nx] = lmd
        lx = self.lmdx[:nx]
        self.stats[0] = np.max(np.abs(lmd - lx)) / np.max(np.abs(lx)) if nx else 0.0
        self.stats[1] = np.max(np.abs(b - np.eye(nx))) if nx else 0.0

    def fetch_ritz(self, nx):
        return (self.v_lmd.a[0, :nx].copy(), self.v_s2.a[0, :nx].copy(), float(self.stats[0]),
                float(self.stats[1]))

    def constraint_coeffs(self, nc, nv):
        T = self.TC.a[:nc, :nv]
        self.QC.a[:nc, :nv] = 2.0 * T - self.Gc.a[:nc, :nc] @ T

    def conjugation(self, nz, ny):
        lmd = self.v_lmd.a[0, :ny]
        num = self.ZAY.a[:nz, :ny] - self.ZBY.a[:nz, :ny] * lmd[None, :]
        den = self.lmdz[:nz, None] - lmd[None, :]
        sy = np.sqrt(np.abs(self.v_y2.a[0, :ny]))
        sz = np.sqrt(np.abs(self.v_t2.a[0, :nz]))
        with np.errstate(divide='ignore', invalid='ignore'):
            ratio = sy[None, :] / sz[:, None]
            keep = ~(np.abs(num) >= 100.0 * ratio * np.abs(den))
            self.Beta.a[:nz, :ny] = np.where(keep, num / np.where(keep, den, 1.0), 0.0)

    def piv_chol(self, G, n, k, eps):
        ind, dropped, status = piv_chol(G.a, n, k, eps)
        self._chol = (dropped, ind, status)

    def fetch_chol(self, n):
        dropped, ind, status = self._chol
        if status:
            raise RuntimeError('Gram matrix of the iterates is not positive definite')
        return dropped, ind

    def ritz_initial(self, n):
        """Generalised problem GA q = lambda GB q in the initial space (solver.py:822)."""
        U = self.GB.a[:n, :n].copy()
        ind, dropped, status = piv_chol(U, n, n, 0.0)
        if status:
            raise RuntimeError('initial vectors are linearly dependent')
        G = transform(self.GA.a[:n, :n], np.triu(U))
        w, Q = self._eigh(G)
        self.CX.a[:n, :n] = sla.solve_triangular(np.triu(U), Q)
        self.lmdx[:n] = w

    def rayleigh_ritz(self, nx, ny, leftX, rightX, leftXn, rightXn):
        nxy = nx + ny
        U = np.triu(self.GB.a[:nxy, :nxy])
        G = transform(self.GA.a[:nxy, :nxy], U)
        lmdy, Qy = self._eigh(G[nx:, nx:])
        G[:, nx:] = G[:, nx:] @ Qy
        G[nx:, :nx] = G[:nx, nx:].T
        G[nx:, nx:] = Qy.T @ G[nx:, nx:]
        w, Q = self._eigh(G)
        sel = np.r_[0:leftX, nxy - rightX:nxy]
        lx = w[sel]
        ly = w[leftX:nxy - rightX]
        QYX = Q[nx:, sel]
        self._est = (np.sqrt(np.sum(QYX * QYX, axis=0)),
                     np.sum((ly[:, None] - lx[None, :]) * QYX * QYX, axis=0))
        Q = Q.copy()
        Q[nx:, :] = Qy @ Q[nx:, :]
        Q = sla.solve_triangular(U, Q)
        seln = np.r_[0:leftXn, nxy - rightXn:nxy]
        nxn = leftXn + rightXn
        nz = nxy - nxn
        self.CX.a[:nxy, :nxn] = Q[:, seln]
        self.CZ.a[:nxy, :nz] = Q[:, leftXn:nxy - rightXn]
        self.lmdx[:nxn] = w[seln]
        self.lmdz[:nz] = w[leftXn:nxy - rightXn]

    def fetch_estimates(self, nx):
        return self._est
