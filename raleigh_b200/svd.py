"""On-device thin SVD of a block of vectors (Vectors.svd()).

Reference: ``v, sigma, wt = numpy.linalg.svd(S, full_matrices=False); S <- wt;
return sigma, conj(v)`` (dense_numpy.py:125-128); the reference's GPU backend
calls cusolverDn?gesvd (dense_cublas.py:537-591).  What callers need
(solver.py:885, partial_svd.py:104,183, lra.py:476-481, tests_algebra.py:330-341):

  (i)  S_old = v . diag(sigma) . S_new   to working precision,
  (ii) S_new has ORTHONORMAL rows -- all m of them, also when S is rank
       deficient or ill-conditioned (the callers feed S_new to Cholesky and to
       generalised eigensolvers; numpy.linalg.svd always returns a full set) --
  (iii) sigma descending.

Algorithm (no host LAPACK; every factorisation runs in libraleigh_b200.so):
  1. G = S S^T in fp64 (exact products of the data, fp64 accumulation) and its
     eigen-decomposition G = Q diag(w) Q^T by one-sided Jacobi, on the Cholesky
     factor of G when that exists (relative accuracy for every eigenvalue);
  2. the Gram route resolves singular values down to ~sqrt(m eps_64) sigma_max.
     Directions above that are LIVE: S1 = diag(w^-1/2) Q^T S.  The others are
     kept unscaled, R = Q_null^T S, projected off the live rows and -- DEFLATION --
     factorised by the same procedure one level down (R is then ~1e-7 of the
     size of S, so its own Gram matrix resolves the next seven decades).  What
     is below the noise floor of the DATA, m eps_data sigma_max, is rank
     deficiency: those rows are replaced by random vectors orthogonalised
     against everything else;
  3. one re-orthonormalisation pass of the whole set (CholQR2 style) when
     S1 S1^T still differs from I beyond working precision;
  4. the accumulated m x m factor B (S = B S1) is decomposed by Jacobi on B
     itself: B^T B = Wb Sigma^2 Wb^T, S_new = Wb^T S1 stays orthonormal exactly
     and v = B Wb Sigma^-1 reproduces S_old = v Sigma S_new.
Only scalars, m x m coefficient matrices of the (rare) deflation branch and the
final (sigma, v) cross to the host.
"""
import numpy

from ._lib import lib, check
from . import device as dev
from . import psvd


def _max_dev_from_identity(work, g_ptr):
    n = work.n
    check(lib.rl_rr_ritz_check(g_ptr, g_ptr, n, n, work.w, work.sigma, work.ger, dev.stream()))
    return float(psvd._fetch(work.ger, 2)[1])


def _upload(ptr, host):
    a = numpy.ascontiguousarray(host, dtype=numpy.float64)
    check(lib.rl_h2d(ptr, dev.host_ptr(a), a.size * 8, dev.stream()))
    check(lib.rl_sync_stream(dev.stream()))


def _complete_columns(U, dead):
    """Replace the `dead` columns of the (m, m) host matrix U by an orthonormal completion of the others
    (modified Gram-Schmidt on unit vectors; no LAPACK)."""
    m = U.shape[0]
    basis = [U[:, j].copy() for j in numpy.nonzero(~dead)[0]]
    for j in numpy.nonzero(dead)[0]:
        best = None
        for e in range(m):
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
    return U


def _orthonormalise_columns(U):
    """Two passes of modified Gram-Schmidt over the columns, left to right (host, m x m, no LAPACK).  A column
    that (numerically) lies in the span of its predecessors -- a direction of zero singular value -- is replaced
    by the best-conditioned unit vector, projected the same way."""
    U = numpy.array(U, dtype=numpy.float64)
    m = U.shape[1]
    for j in range(m):
        c = U[:, j]
        n0 = numpy.sqrt(c @ c)
        for _ in range(2):
            if j:
                c = c - U[:, :j] @ (U[:, :j].T @ c)
        nrm = numpy.sqrt(c @ c)
        if not (nrm > 1e-6 * max(n0, 1e-300)) or not numpy.isfinite(nrm):
            best = None
            for e in range(m):
                t = numpy.zeros(m)
                t[e] = 1.0
                for _ in range(2):
                    if j:
                        t = t - U[:, :j] @ (U[:, :j].T @ t)
                tn = numpy.sqrt(t @ t)
                if best is None or tn > best[0]:
                    best = (tn, t)
                if tn > 0.5:
                    break
            nrm, c = best
        U[:, j] = c / nrm
    return U


def _random_orthonormal(like, k, against, salt):
    """k random vectors orthogonal to the (orthonormal) blocks in `against` and to each other."""
    R = like.new_vectors(k)
    R.fill_random_device(977 + 31 * k + salt, row0=like._shard[1] if like._shard is not None else 0)
    for blk in against:
        if blk is not None and blk.nvec() > 0:
            for _ in range(2):
                R.orthogonalize(blk)
    block_svd(R, floor=0.0, against=against)      # full rank with probability one
    return R


def block_svd(v, floor=None, against=()):
    """SVD of the selected block of `v` in place.  `floor`: absolute singular-value level below which a
    direction is rank deficiency (default: m eps_data sigma_max); `against`: orthonormal blocks the random
    completion of such directions must also be orthogonal to (used by the deflation levels)."""
    m = v.nvec()
    dt = v.data_type()
    if m < 1:
        return numpy.zeros((0,), dtype=dt), numpy.zeros((0, 0), dtype=dt)
    st = dev.stream
    f0 = v.selected()[0]
    eps = float(numpy.finfo(dt).eps)
    eps64 = float(numpy.finfo(numpy.float64).eps)
    work = psvd._Work(m, numpy.float64)            # full-precision Jacobi: the factors must be orthogonal to eps
    tmp = v.new_vectors(m)
    extra = dev.Buffer(3 * m * m * 8 + 256, zero=True)
    B, T1, T2 = extra.ptr, extra.ptr + m * m * 8, extra.ptr + 2 * m * m * 8

    # 1. Gram matrix and its eigen-decomposition
    psvd._gram64(v, work, exact=True)
    factored = psvd._factor(work)
    psvd._eigh_gram(work, factored)
    w = psvd._fetch(work.w, m)
    wmax = max(float(w[-1]), 0.0)
    smax = numpy.sqrt(wmax)
    if floor is None:
        floor = m * eps * smax
    resolvable = 64.0 * m * eps64 * wmax                   # what an fp64 Gram matrix can tell from rounding noise
    live = w > max(resolvable, floor * floor)
    if wmax == 0.0:
        live[:] = False
    nnull = int(m - numpy.count_nonzero(live))
    root = numpy.where(live, numpy.sqrt(numpy.where(live, w, 1.0)), 0.0)
    coef = numpy.where(live, 1.0 / numpy.where(live, root, 1.0), 1.0)      # null directions: kept UNSCALED

    # 2. S1 = diag(coef) Q^T S
    scale = dev.Buffer(2 * m * 8 + 256)
    _upload(scale.ptr, numpy.concatenate((coef, root)))
    check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr, T1, m, st()))              # T1 = Q diag(coef)
    psvd._rotate(v, T1, m, tmp)
    if nnull == 0:
        check(lib.rl_small_scale_cols(work.Q, m, m, m, scale.ptr + m * 8, B, m, st()))   # B = Q diag(root)
    else:
        Q = psvd._fetch(work.Q, m * m).reshape(m, m)
        inull, ilive = numpy.nonzero(~live)[0], numpy.nonzero(live)[0]
        L = None
        if len(ilive) > 0:
            L = v.new_vectors(len(ilive))
            v.copy(L, [f0 + int(i) for i in ilive])
        R = v.new_vectors(nnull)
        v.copy(R, [f0 + int(i) for i in inull])
        C = numpy.zeros((nnull, len(ilive)))
        if L is not None:
            for _ in range(2):
                C += R.orthogonalize(L).data().astype(numpy.float64).T
        rn = numpy.sqrt(numpy.abs(R.dots(R).astype(numpy.float64)))
        M = numpy.zeros((m, m))
        M[ilive, ilive] = root[ilive]
        M[numpy.ix_(inull, ilive)] = C
        others = tuple(against) + ((L,) if L is not None else ())
        if rn.size and float(numpy.max(rn)) > floor:
            sig_r, v_r = block_svd(R, floor=floor, against=others)                       # deflation: one level down
            M[numpy.ix_(inull, inull)] = v_r.astype(numpy.float64) * sig_r.astype(numpy.float64)[None, :]
        else:
            R = _random_orthonormal(v, nnull, others, salt=m)                            # rank deficiency
        for t, i in enumerate(inull):
            R.select(1, t)
            v.select(1, f0 + int(i))
            R.copy(v)
        v.select(m, f0)
        _upload(B, Q @ M)

    # 3. re-orthonormalise when needed
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

    # 4. B = Ub Sigma Wb^T from Jacobi on B itself (eigenvectors of B^T B = rows of B rotated)
    check(lib.rl_small_eigh_factor(B, m, m, 0.0, work.w, work.Q, m, work.ews, work.ews_bytes, work.info, st()))
    check(lib.rl_psvd_coeffs(work.Q, m, work.w, m, work.q, work.cs, m, work.sigma, st()))  # q = Wb (descending), cs = Wb / sigma
    sigma = psvd._fetch(work.sigma, m)
    dead = sigma <= 0.0
    if dead.any():
        # singular B (rank deficient block): the null columns of Wb come back as zeros; any orthonormal completion
        Wb = _complete_columns(psvd._fetch(work.q, m * m).reshape(m, m), dead)
        _upload(work.q, Wb)
    psvd._rotate(v, work.q, m, tmp)                                                      # S_new = Wb^T S1
    check(lib.rl_small_gemm(0, 0, m, m, m, 1.0, B, m, work.cs, m, 0.0, T1, m, st()))     # v = B Wb Sigma^-1
    U = psvd._fetch(T1, m * m).reshape(m, m)
    if dead.any() or sigma[-1] < 1e-3 * sigma[0]:
        # v_j = B Wb_j / sigma_j carries an error ~ eps |B| / sigma_j: re-orthonormalise the columns in the order of
        # decreasing sigma (the accurate ones first).  A change delta of v_j moves the reconstruction by sigma_j delta,
        # i.e. by ~eps |S|: (i) is untouched, and v becomes orthonormal like numpy.linalg.svd's.
        U = _orthonormalise_columns(U)
    return sigma.astype(dt), U.astype(dt)
