"""On-device thin SVD of a block of vectors (Vectors.svd()).

Reference: ``v, sigma, wt = numpy.linalg.svd(S, full_matrices=False); S <- wt;
return sigma, conj(v)`` (dense_numpy.py:125-128); the reference's GPU backend
calls cusolverDn?gesvd (dense_cublas.py:537-591).  What callers need
(solver.py:885, partial_svd.py:104,183, lra.py:476-481, tests_algebra.py:330-341):

  (i)  S_old = v . diag(sigma) . S_new   to working precision,
  (ii) S_new has orthonormal rows, sigma descending.

Algorithm (no host LAPACK; every decomposition runs in libraleigh_b200.so):
  1. G = S S^T accumulated in fp64 whatever the data type (rl_gram_acc64), so
     the squared condition number is resolved in fp64;
  2. cyclic-Jacobi eigendecomposition of G on the device (rl_syevj);
  3. S1 = diag(lambda^-1/2) V^T S (rl_update), B = V diag(lambda^1/2);
  4. one re-orthonormalisation sweep S2 = diag(mu^-1/2) W^T S1 from the Gram
     matrix of S1 (CholQR2/SVQB style) when S1 S1^T is not yet the identity;
     B <- B W diag(mu^1/2);
  5. B = U_b Sigma_b W_b^T from the device eigendecomposition of B^T B;
     S_new = W_b^T S2 keeps its rows orthonormal exactly, and
     v = B W_b Sigma_b^-1 reproduces S_old = v Sigma_b S_new identically.
Only products and scalings of the small (m, m) factors are done on the host.
"""
import numpy

from ._lib import lib, check
from . import device as dev


def _gram64(v):
    """(m, m) fp64 Gram matrix of the selected block, left on the device."""
    m, n = v.nvec(), v.local_dimension()
    wsb = lib.rl_gram_acc64_ws_bytes(v._code, m, m, n)
    ws = dev.Buffer(wsb) if wsb else None
    g = dev.Buffer(m * m * 8)
    check(lib.rl_gram_acc64(v._code, v._wptr(), v._ld, m, v._wptr(), v._ld, m, n, g.ptr,
                            ws.ptr if ws else 0, wsb, dev.stream()))
    v._reduce_device(g, m * m, numpy.float64)      # row-sharded block: sum the partial Gram matrices
    return g


def device_eigh(g_buf, p):
    """Eigendecomposition of the symmetric fp64 (p, p) matrix in g_buf (device,
    overwritten).  Returns host (w ascending, V with eigenvectors as columns)."""
    import ctypes
    wsb = lib.rl_syevj_ws_bytes(p)
    ws = dev.Buffer(wsb)
    w_d = dev.Buffer(p * 8)
    sweeps = ctypes.c_int(0)
    check(lib.rl_syevj(g_buf.ptr, p, w_d.ptr, ws.ptr, wsb, ctypes.byref(sweeps), dev.stream()))
    w = numpy.empty((p,), dtype=numpy.float64)
    V = numpy.empty((p, p), dtype=numpy.float64)
    check(lib.rl_d2h(dev.host_ptr(w), w_d.ptr, p * 8, dev.stream()))
    check(lib.rl_d2h(dev.host_ptr(V), g_buf.ptr, p * p * 8, dev.stream()))
    return w, V


def _host_sym_to_device(a):
    a = numpy.ascontiguousarray(a, dtype=numpy.float64)
    buf = dev.Buffer(a.size * 8)
    check(lib.rl_h2d(buf.ptr, dev.host_ptr(a), a.size * 8, dev.stream()))
    check(lib.rl_sync_stream(dev.stream()))
    return buf


def _apply_left(v, coeff, tmp):
    """block <- coeff^T . block  (coeff is (m, m) host fp64) through `tmp`."""
    m, n = v.nvec(), v.local_dimension()
    q = numpy.ascontiguousarray(coeff, dtype=v.data_type())
    v._touch()
    check(lib.rl_update_h(v._code, tmp._wptr(), tmp._ld, m, v._wptr(), v._ld, m, dev.host_ptr(q), m, 1,
                          1.0, 0.0, n, dev.stream()))
    check(lib.rl_copy(v._code, v._wptr(), v._ld, tmp._wptr(), tmp._ld, m, n, dev.stream()))


def block_svd(v):
    m, n = v.nvec(), v.dimension()
    dt = v.data_type()
    if m < 1:
        return numpy.zeros((0,), dtype=dt), numpy.zeros((0, 0), dtype=dt)
    eps = float(numpy.finfo(dt).eps)
    tmp = v.new_vectors(m)

    # pass 1.  Directions whose singular value is zero at the data precision
    # (sigma <= sigma_max * m * eps) are NULL: they get weight 0 instead of 1/tiny,
    # so rank-deficient blocks (lra.py:283-285 appends zero vectors) stay finite.
    w, V = device_eigh(_gram64(v), m)
    wmax = max(float(w[-1]), 0.0)
    live = w > wmax * (m * eps) ** 2
    root = numpy.where(live, numpy.sqrt(numpy.where(live, w, 1.0)), 0.0)
    iroot = numpy.where(live, 1.0 / numpy.where(live, root, 1.0), 0.0)
    _apply_left(v, V * iroot[None, :], tmp)          # S1 = diag(1/root) V^T S
    B = V * root[None, :]                            # S = B S1

    # pass 2: re-orthonormalise when S1 S1^T differs from I beyond working precision
    g2 = _gram64(v)
    G2 = numpy.empty((m, m), dtype=numpy.float64)
    check(lib.rl_d2h(dev.host_ptr(G2), g2.ptr, m * m * 8, dev.stream()))
    target = numpy.diag(live.astype(numpy.float64))
    if numpy.amax(abs(G2 - target)) > 4 * eps:
        mu, W = device_eigh(g2, m)
        live2 = mu > 1e-6
        rmu = numpy.where(live2, numpy.sqrt(numpy.where(live2, mu, 1.0)), 0.0)
        irmu = numpy.where(live2, 1.0 / numpy.where(live2, rmu, 1.0), 0.0)
        _apply_left(v, W * irmu[None, :], tmp)       # S2 = diag(1/rmu) W^T S1
        B = (B @ W) * rmu[None, :]

    # small SVD of B through the device eigensolver: B^T B = Wb diag(s^2) Wb^T
    s2, Wb = device_eigh(_host_sym_to_device(B.T @ B), m)
    order = numpy.argsort(-s2, kind='stable')
    s2 = numpy.maximum(s2[order], 0.0)
    Wb = Wb[:, order]
    sigma = numpy.sqrt(s2)
    _apply_left(v, Wb, tmp)                          # S_new = Wb^T S2
    safe = numpy.where(sigma > 0, sigma, 1.0)
    U = (B @ Wb) / safe[None, :]
    return sigma.astype(dt), U.astype(dt)
