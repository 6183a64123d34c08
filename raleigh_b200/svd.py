"""On-device thin SVD of a block of vectors (Vectors.svd()).

Reference: ``v, sigma, wt = numpy.linalg.svd(S, full_matrices=False); S <- wt;
return sigma, conj(v)`` (dense_numpy.py:125-128); the reference's GPU backend
calls cusolverDn?gesvd (dense_cublas.py:537-591).  What callers need
(solver.py:885, partial_svd.py:104,183, lra.py:476-481, tests_algebra.py:330-341):

  (i)  S_old = v . diag(sigma) . S_new   to working precision,
  (ii) S_new has ORTHONORMAL rows -- all m of them, also when S is rank
       deficient or so ill-conditioned that some directions drown in rounding
       noise (the callers feed S_new to Cholesky / generalised eigensolvers) --
  (iii) sigma descending.

Algorithm (no host LAPACK; every factorisation runs in libraleigh_b200.so):
  1. G = S S^T in fp64 whatever the data type;
  2. G = Q diag(w) Q^T by one-sided Jacobi on the Cholesky factor of G when that
     exists (relative accuracy for every eigenvalue), by the shifted symmetric
     solver otherwise.  Directions with w below the noise floor of the
     computation, max((m eps_data)^2, m eps_64) w_max, are NULL;
  3. S1 = diag(w^-1/2) Q^T S on the live directions; NULL rows are replaced by
     random vectors orthogonalised against the live rows and each other, so
     S1 is a full orthonormal set (numpy.linalg.svd returns such a completion);
  4. one re-orthonormalisation pass S2 = diag(mu^-1/2) W^T S1 (CholQR2 style)
     when S1 S1^T still differs from I beyond working precision;
  5. the accumulated m x m factor B (S = B S2) is decomposed by Jacobi on B
     itself: B^T B = Wb Sigma^2 Wb^T, S_new = Wb^T S2 stays orthonormal exactly
     and v = B Wb Sigma^-1 reproduces S_old = v Sigma S_new.
Only scalars and the final (sigma, v) cross to the host.
"""
import numpy

from ._lib import lib, check
from . import device as dev
from . import psvd


def _max_dev_from_identity(work, g_ptr):
    n = work.n
    check(lib.rl_rr_ritz_check(g_ptr, g_ptr, n, n, work.w, work.sigma, work.ger, dev.stream()))
    return float(psvd._fetch(work.ger, 2)[1])


def block_svd(v):
    m = v.nvec()
    dt = v.data_type()
    if m < 1:
        return numpy.zeros((0,), dtype=dt), numpy.zeros((0, 0), dtype=dt)
    st = dev.stream
    eps = float(numpy.finfo(dt).eps)
    eps64 = float(numpy.finfo(numpy.float64).eps)
    work = psvd._Work(m, numpy.float64)            # full-precision Jacobi: the factors must be orthogonal to eps
    tmp = v.new_vectors(m)
    extra = dev.Buffer(3 * m * m * 8, zero=True)
    B, T1, T2 = extra.ptr, extra.ptr + m * m * 8, extra.ptr + 2 * m * m * 8

    # 1-2. Gram matrix and its eigen-decomposition
    psvd._gram64(v, work, exact=True)
    factored = psvd._factor(work)
    psvd._eigh_gram(work, factored)
    w = psvd._fetch(work.w, m)
    wmax = max(float(w[-1]), 0.0)
    floor = max((m * eps) ** 2, m * eps64) * wmax
    live = (w > floor) if not factored else (w > (m * eps) ** 2 * wmax)
    if wmax == 0.0:
        live[:] = False
    root = numpy.where(live, numpy.sqrt(numpy.where(live, w, 1.0)), 0.0)
    iroot = numpy.where(live, 1.0 / numpy.where(live, root, 1.0), 0.0)

    # 3. S1 = diag(iroot) Q^T S;  B = Q diag(root)  (S = B S1 on the live directions)
    scale = dev.Buffer(2 * m * 8)
    _upload(scale.ptr, numpy.concatenate((iroot, root)))
    check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr, T1, m, st()))              # T1 = Q diag(iroot)
    check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr + m * 8, B, m, st()))       # B  = Q diag(root)
    psvd._rotate(v, T1, m, tmp)
    nnull = int(m - numpy.count_nonzero(live))
    if nnull > 0:
        _complete(v, numpy.nonzero(~live)[0], numpy.nonzero(live)[0])

    # 4. re-orthonormalise when needed
    psvd._gram64(v, work, exact=True)
    if _max_dev_from_identity(work, work.G) > 4 * eps:
        if psvd._factor(work):
            psvd._eigh_gram(work, True)
            mu = psvd._fetch(work.w, m)
            rmu = numpy.sqrt(numpy.maximum(mu, 0.0))
            irmu = numpy.where(rmu > 0, 1.0 / numpy.where(rmu > 0, rmu, 1.0), 0.0)
            _upload(scale.ptr, numpy.concatenate((irmu, rmu)))
            check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr, T1, m, st()))      # W diag(1/rmu)
            psvd._rotate(v, T1, m, tmp)                                                  # S2 = diag(1/rmu) W^T S1
            check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr + m * 8, T1, m, st()))   # W diag(rmu)
            check(lib.rl_small_gemm(0, 0, m, m, m, 1.0, B, m, T1, m, 0.0, T2, m, st()))  # B <- B W diag(rmu)
            B, T2 = T2, B

    # 5. B = Ub Sigma Wb^T from Jacobi on B itself (eigenvectors of B^T B = rows of B rotated)
    check(lib.rl_small_eigh_factor(B, m, m, 0.0, work.w, work.Q, m, work.ews, work.ews_bytes, work.info, st()))
    check(lib.rl_psvd_coeffs(work.Q, m, work.w, m, work.q, work.cs, m, work.sigma, st()))  # q = Wb (descending), cs = Wb / sigma
    psvd._rotate(v, work.q, m, tmp)                                                      # S_new = Wb^T S2
    check(lib.rl_small_gemm(0, 0, m, m, m, 1.0, B, m, work.cs, m, 0.0, T1, m, st()))     # v = B Wb Sigma^-1
    sigma = psvd._fetch(work.sigma, m)
    U = psvd._fetch(T1, m * m).reshape(m, m)
    dead = sigma <= 0.0
    if dead.any():
        # left vectors of exactly zero singular values: any orthonormal completion (numpy returns one)
        basis = [U[:, j] for j in numpy.nonzero(~dead)[0]]
        for j in numpy.nonzero(dead)[0]:
            best = None
            for e in range(m):                       # Gram-Schmidt on the unit vectors, keep the best conditioned
                c = numpy.zeros(m)
                c[e] = 1.0
                for _ in range(2):
                    for b in basis:
                        c -= (b @ c) * b
                nrm = numpy.sqrt(c @ c)
                if best is None or nrm > best[0]:
                    best = (nrm, c / max(nrm, 1e-300))
                if nrm > 0.5:
                    break
            U[:, j] = best[1]
            basis.append(best[1])
    return sigma.astype(dt), U.astype(dt)


def _upload(ptr, host):
    a = numpy.ascontiguousarray(host, dtype=numpy.float64)
    check(lib.rl_h2d(ptr, dev.host_ptr(a), a.size * 8, dev.stream()))
    check(lib.rl_sync_stream(dev.stream()))


def _complete(v, null_rows, live_rows):
    """Replace the NULL rows of the selected block by an orthonormal set orthogonal to the live rows."""
    f, m = v.selected()
    k = len(null_rows)
    R = v.new_vectors(k)
    seed = 977 + 31 * k + m          # fixed: numpy.linalg.svd does not advance the host RNG stream either
    R.fill_random_device(seed, row0=v._shard[1] if v._shard is not None else 0)
    if len(live_rows) > 0:
        L = v.new_vectors(len(live_rows))
        v.copy(L, [f + int(i) for i in live_rows])
        for _ in range(2):
            R.orthogonalize(L)
    R.svd()                                   # full rank with probability one: recursion depth 1
    for t, i in enumerate(null_rows):
        R.select(1, t)
        v.select(1, f + int(i))
        R.copy(v)
    R.select(k)
    v.select(m, f)
