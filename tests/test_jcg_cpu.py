"""CPU checks of the device-resident block-CG driver's HOST logic (raleigh_b200/jcg.py, jcg_host.py):
the driver runs here on the NumPy engine (oracle/jcg_engine_np.py, test infrastructure) and on host
vectors, against the reference's own UNMODIFIED solver on the same inputs and seeds.  Same algorithm =>
identical iteration counts and eigenvalues to rounding.  Also pins the NumPy statements of the two
sequential kernels (pivoted Cholesky with the reference's drop rule, one-sided Jacobi) that the CUDA
kernels are tested against on the GPU."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import algebra_np as K
from oracle import jcg_engine_np as E
from oracle.host_backend import Vectors, SparseSymmetricMatrix, Operator, Jacobi
from tests_common import spd_c3_like


@pytest.fixture(scope='module')
def rs(ref_root):
    import sys
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    from tools import refenv
    refenv.load_reference()
    import raleigh.core.solver as solver
    return solver


def _run(rs, device, L, dtype, which, tol, block, jac=False, eigh='lapack', crit='k eigenvector error'):
    from raleigh_b200 import jcg
    np.random.seed(1)
    opt = rs.Options()
    opt.block_size = block
    opt.max_iter = 1000
    opt.convergence_criteria = rs.DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance(crit, tol)
    n = L.shape[0]
    v = Vectors(n, data_type=dtype)
    solver = rs.Solver(rs.Problem(v, SparseSymmetricMatrix(L.astype(dtype))))
    if jac:
        solver.set_preconditioner(Operator(Jacobi(L)))
    orig = rs.Solver._solve
    if device:
        eng = E.NumpyEngine(eigh)
        rs.Solver._solve = lambda self, ev, o, w, e, i: jcg.solve(self, ev, o, w, e, i, eng)
    try:
        status = solver.solve(v, opt, which=which)
    finally:
        rs.Solver._solve = orig
    return status, solver.iteration, np.array(solver.eigenvalues), solver


CASES = [
    ('lap12 left, block 8', 12, np.float64, (6, 0), 1e-6, 8, False),
    ('lap12 both ends', 12, np.float64, (4, 3), 1e-5, 12, False),
    ('lap12 largest', 12, np.float64, 6, 1e-6, 8, False),
    ('lap10 right only', 10, np.float64, (0, 5), 1e-6, 8, False),
    ('lap12 fp32', 12, np.float32, (6, 0), 1e-3, 8, False),
]


@pytest.mark.parametrize('name,N,dtype,which,tol,block,jac', CASES, ids=[c[0] for c in CASES])
def test_driver_reproduces_the_reference_iteration(rs, name, N, dtype, which, tol, block, jac):
    L = K.lap3d_csr(N, N, N)
    s0, it0, lmd0, _ = _run(rs, False, L, dtype, which, tol, block, jac)
    s1, it1, lmd1, sol = _run(rs, True, L, dtype, which, tol, block, jac)
    assert s0 == s1 == 0
    assert it1 == it0, (it0, it1)
    assert len(lmd0) == len(lmd1)
    rel = np.max(np.abs(np.sort(lmd1) - np.sort(lmd0)) / np.abs(np.sort(lmd0)))
    assert rel < (1e-12 if dtype is np.float64 else 1e-5)
    # the reporting arrays of the Solver object are filled like the reference's
    assert len(sol.residual_norms) == len(lmd1) and len(sol.convergence_status) == len(lmd1)
    assert sol.eigenvector_errors.kinematic.shape[0] == len(lmd1)


def test_driver_with_jacobi_preconditioner_and_locked_vectors(rs):
    A = spd_c3_like(3000)
    s0, it0, lmd0, _ = _run(rs, False, A, np.float64, (5, 0), 1e-6, 8, jac=True)
    s1, it1, lmd1, _ = _run(rs, True, A, np.float64, (5, 0), 1e-6, 8, jac=True)
    assert s0 == s1 == 0 and it0 == it1
    assert np.max(np.abs(lmd1 - lmd0) / lmd0) < 1e-12


def test_driver_with_the_jacobi_eigensolver_statement(rs):
    """The Rayleigh-Ritz eigenproblems solved by the NumPy statement of csrc/jacobi.cu (shifted one-sided
    Jacobi, two-level rotation order) instead of LAPACK: same iteration count."""
    L = K.lap3d_csr(10, 10, 10)
    s0, it0, lmd0, _ = _run(rs, False, L, np.float64, (5, 0), 1e-6, 8)
    s1, it1, lmd1, _ = _run(rs, True, L, np.float64, (5, 0), 1e-6, 8, eigh='jacobi')
    assert s0 == s1 == 0 and abs(it1 - it0) <= 1
    assert np.max(np.abs(lmd1 - lmd0) / lmd0) < 1e-11


def test_driver_runs_the_reference_pca_flows(rs):
    """pca(npc=...), pca(tol=...) through lra / partial_svd with the driver in place of Solver._solve
    (stopping criteria objects, 'largest' selection, hundreds of locked vectors, fp32)."""
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    from raleigh_b200 import jcg
    np.random.seed(1)
    A, sigma, u, v = generate(600, 400, 200, pca=True)
    out = {}
    for device in (False, True):
        orig = rs.Solver._solve
        if device:
            eng = E.NumpyEngine()
            rs.Solver._solve = lambda self, ev, o, w, e, i: jcg.solve(self, ev, o, w, e, i, eng)
        try:
            np.random.seed(1)
            m1, t1, c1 = pca(A, npc=40, opt=rs.Options())
            np.random.seed(1)
            m2, t2, c2 = pca(A, tol=0.1, opt=rs.Options())
        finally:
            rs.Solver._solve = orig
        out[device] = (c1.shape[0], pca_error(A, m1, t1, c1), np.linalg.norm(t1, axis=0), c2.shape[0],
                       pca_error(A, m2, t2, c2))
    a, b = out[False], out[True]
    assert a[0] == b[0] and abs(a[3] - b[3]) <= 2
    assert abs(a[1][0] - b[1][0]) < 1e-4 and abs(a[1][1] - b[1][1]) < 1e-4
    assert np.max(np.abs(a[2][:20] - b[2][:20]) / a[2][:20]) < 1e-5
    assert b[4][1] <= 0.1


@pytest.mark.parametrize('n,k,rank_y,noise', [(12, 0, 12, 0.0), (40, 20, 20, 0.0), (40, 20, 10, 0.0), (40, 20, 10, 1e-7),
                                              (140, 70, 50, 1e-6), (200, 60, 90, 1e-5), (96, 16, 79, 3e-5)])
def test_pivoted_cholesky_statement_against_the_reference(rs, n, k, rank_y, noise):
    """oracle.jcg_engine_np.piv_chol (right-looking, what csrc/rr.cu implements) against the reference's
    blocked left-looking _piv_chol (solver.py:1749-1826): same permutation, same drop count, same factor."""
    rng = np.random.RandomState(n + rank_y)
    N = 4 * n
    X = np.linalg.qr(rng.randn(N, max(k, 1)))[0][:, :k]
    Y = rng.randn(N, rank_y) @ rng.randn(rank_y, n - k) + noise * rng.randn(N, n - k)
    if k:
        Y -= X @ (X.T @ Y)
    Y /= np.linalg.norm(Y, axis=0)
    V = np.concatenate((X, Y), axis=1)
    A = V.T @ V
    Ar, Am = A.copy(), A.copy()
    ind_r, dropped_r = rs._piv_chol(Ar, k, 1e-8)
    ind_m, dropped_m, status = E.piv_chol(Am, n, k, 1e-8)
    assert status == 0
    assert abs(dropped_m - dropped_r) <= 1, (dropped_m, dropped_r)
    kept = n - max(dropped_m, dropped_r)
    if dropped_m == dropped_r:
        assert list(ind_m[:kept]) == list(ind_r[:kept])
        assert np.max(np.abs(np.triu(Am)[:kept, :kept] - np.triu(Ar)[:kept, :kept])) < 1e-7
    Uk = np.triu(Am)[:n - dropped_m, :n - dropped_m]
    P = A[np.ix_(ind_m[:n - dropped_m], ind_m[:n - dropped_m])]
    assert np.max(np.abs(Uk.T @ Uk - P)) < 1e-10


@pytest.mark.parametrize('n', [1, 2, 3, 7, 16, 33, 64, 100])
def test_jacobi_statement(n):
    rng = np.random.RandomState(n)
    G = rng.randn(n, n)
    G = G + G.T
    w, Q, sweeps = E.jacobi_eigh(G)
    wr = np.linalg.eigvalsh(G)
    scale = max(1.0, np.max(np.abs(wr)))
    assert np.max(np.abs(w - wr)) < 1e-12 * scale * max(n, 4)
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) < 1e-13 * max(n, 4)
    assert np.max(np.abs(G @ Q - Q * w[None, :])) < 1e-12 * scale * max(n, 4)
    assert sweeps < 20


def test_history_shift_matches_the_reference_loops(rs):
    """jcg_host.History.shift (slice moves) against the per-element loops of solver.py:1543-1587."""
    from raleigh_b200.jcg_host import History
    rng = np.random.RandomState(2)
    m = 12
    for l, nl, sl, sr in ((6, 6, 2, 1), (6, 6, 0, 3), (6, 4, -4, 2), (6, 9, 1, -5), (5, 5, 0, 0), (7, 7, 3, 2)):
        h = History(m, 1e-16)
        for a in (h.cnv, h.iterations):
            a[:] = rng.randint(1, 9, size=a.shape)
        for a in (h.lmd, h.res, h.dX, h.acf, h.err_lmd, h.err_X, h.dlmd):
            a[...] = rng.rand(*a.shape)
        ref = {k: getattr(h, k).copy() for k in ('cnv', 'lmd', 'res', 'acf', 'err_lmd', 'dlmd', 'err_X', 'dX', 'iterations')}
        h.shift(l, nl, sl, sr)
        cnv, lmd, res, acf, err_lmd, dlmd, err_X, dX, iterations = (ref[k] for k in
                                                                   ('cnv', 'lmd', 'res', 'acf', 'err_lmd', 'dlmd', 'err_X', 'dX', 'iterations'))

        def move(i, j):
            cnv[i] = cnv[j]; lmd[i] = lmd[j]; res[i] = res[j]; acf[:, i] = acf[:, j]
            err_lmd[:, i] = err_lmd[:, j]; dlmd[i, :] = dlmd[j, :]; err_X[:, i] = err_X[:, j]
            dX[i] = dX[j]; iterations[i] = iterations[j]

        def reset(i):
            rs._reset_cnv_data(i, cnv, res, acf, err_lmd, dlmd, err_X, dX, iterations)
        if sl > 0:
            for i in range(l - sl):
                move(i, i + sl)
        if sl >= 0:
            for i in range(l - sl, nl):
                reset(i)
        else:
            for i in range(l):
                reset(i)
        if sr > 0:
            for i in range(m - 1, l + sr - 1, -1):
                move(i, i - sr)
        if sr >= 0:
            for i in range(l + sr - 1, nl - 1, -1):
                reset(i)
        else:
            for i in range(l, m):
                reset(i)
        for k in ref:
            assert np.array_equal(getattr(h, k), ref[k]), (k, l, nl, sl, sr)


def _mass_matrix(n, seed=0):
    """SPD 'mass matrix' with a varying diagonal and weak nearest-neighbour coupling."""
    import scipy.sparse as sp
    rng = np.random.RandomState(seed)
    d = 1.0 + rng.rand(n)
    o = 0.1 * rng.rand(n - 1)
    return (sp.diags(d) + sp.diags(o, 1) + sp.diags(o, -1)).tocsr()


def _run_gen(rs, device, L, M, which, tol, block, jac=False, dtype=np.float64, product=False):
    from raleigh_b200 import jcg
    np.random.seed(1)
    opt = rs.Options()
    opt.block_size = block
    opt.max_iter = 1000
    opt.convergence_criteria = rs.DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance('k eigenvector error', tol)
    n = L.shape[0]
    v = Vectors(n, data_type=dtype)
    # NB: Problem(v, A, B) is the generalised problem; ANY fourth argument -- also the string 'gen' that
    # partial_hevp.py:216 passes -- selects the product form A B x = lambda x (solver.py:241-250)
    if product:
        problem = rs.Problem(v, SparseSymmetricMatrix(L.astype(dtype)), SparseSymmetricMatrix(M.astype(dtype)), 'pro')
        assert problem.type() == 'p'
    else:
        problem = rs.Problem(v, SparseSymmetricMatrix(L.astype(dtype)), SparseSymmetricMatrix(M.astype(dtype)))
        assert problem.type() == 'g'
    solver = rs.Solver(problem)
    if jac:
        solver.set_preconditioner(Operator(Jacobi(L)))
    orig = rs.Solver._solve
    if device:
        eng = E.NumpyEngine('lapack')
        rs.Solver._solve = lambda self, ev, o, w, e, i: jcg.solve(self, ev, o, w, e, i, eng)
    try:
        status = solver.solve(v, opt, which=which)
    finally:
        rs.Solver._solve = orig
    return status, solver.iteration, np.array(solver.eigenvalues), v.data().copy(), solver


GEN_CASES = [
    ('gen left, block 8', 10, (6, 0), 1e-6, 8, False),
    ('gen both ends', 10, (3, 3), 1e-5, 12, False),
    ('gen largest', 10, 5, 1e-6, 8, False),
    ('gen left, Jacobi preconditioner', 12, (6, 0), 1e-6, 8, True),
]


@pytest.mark.parametrize('name,N,which,tol,block,jac', GEN_CASES, ids=[c[0] for c in GEN_CASES])
def test_driver_generalised_problem_reproduces_the_reference_iteration(rs, name, N, which, tol, block, jac):
    """A x = lambda B x (solver.py 'gen' branches :684-689, 747-752, 949, 963, 1213-1222, 1382-1391, 1626-1641):
    B-images of X, Y, Z and of the locked vectors carried along, B-Gram matrices, residuals A X - B X lambda."""
    L = K.lap3d_csr(N, N, N)
    M = _mass_matrix(L.shape[0])
    s0, it0, lmd0, x0, _ = _run_gen(rs, False, L, M, which, tol, block, jac)
    s1, it1, lmd1, x1, sol = _run_gen(rs, True, L, M, which, tol, block, jac)
    assert s0 == s1 == 0
    assert it1 == it0, (it0, it1)
    assert len(lmd0) == len(lmd1)
    o0, o1 = np.argsort(lmd0), np.argsort(lmd1)
    assert np.max(np.abs(lmd1[o1] - lmd0[o0]) / np.abs(lmd0[o0])) < 1e-11
    # B-orthonormal eigenvectors with small residuals, and the image block the reference exposes
    Lf, Mf = L.toarray(), M.toarray()
    x = x1.T
    assert np.max(np.abs(x.T @ Mf @ x - np.eye(x.shape[1]))) < 1e-6
    res = Lf @ x - (Mf @ x) * lmd1[None, :]
    assert np.max(np.linalg.norm(res, axis=0)) < 1e-3 * np.max(np.abs(lmd1))
    assert np.allclose(sol.eigenvectors_im.data(), (Mf @ x).T, rtol=1e-9, atol=1e-9)
    # exact eigenvalues of the pencil
    exact = sla.eigh(Lf, Mf, eigvals_only=True)
    lo = np.sort(lmd1)
    nearest = exact[np.argmin(np.abs(exact[None, :] - lo[:, None]), axis=1)]
    assert np.max(np.abs(lo - nearest) / np.abs(nearest)) < 1e-8


PRO_CASES = [
    ('pro left, block 8', 10, (6, 0), 1e-6, 8),
    ('pro both ends', 10, (3, 3), 1e-5, 12),
    ('pro largest', 8, 4, 1e-6, 8),
]


@pytest.mark.parametrize('name,N,which,tol,block', PRO_CASES, ids=[c[0] for c in PRO_CASES])
def test_driver_product_problem_reproduces_the_reference_iteration(rs, name, N, which, tol, block):
    """A B x = lambda x (solver.py 'pro' branches: A applied to the B-images, residuals A B X - X lambda measured in
    the B-norm, no preconditioning step, image block updated alongside the search directions)."""
    L = K.lap3d_csr(N, N, N)
    M = _mass_matrix(L.shape[0])
    s0, it0, lmd0, x0, _ = _run_gen(rs, False, L, M, which, tol, block, product=True)
    s1, it1, lmd1, x1, sol = _run_gen(rs, True, L, M, which, tol, block, product=True)
    assert s0 == s1 == 0
    assert it1 == it0, (it0, it1)
    assert len(lmd0) == len(lmd1)
    o0, o1 = np.argsort(lmd0), np.argsort(lmd1)
    assert np.max(np.abs(lmd1[o1] - lmd0[o0]) / np.abs(lmd0[o0])) < 1e-11
    Lf, Mf = L.toarray(), M.toarray()
    x = x1.T
    assert np.max(np.abs(x.T @ Mf @ x - np.eye(x.shape[1]))) < 1e-6
    res = Lf @ (Mf @ x) - x * lmd1[None, :]
    assert np.max(np.linalg.norm(res, axis=0)) < 1e-3 * np.max(np.abs(lmd1))
    exact = np.sort(np.linalg.eigvals(Lf @ Mf).real)
    lo = np.sort(lmd1)
    nearest = exact[np.argmin(np.abs(exact[None, :] - lo[:, None]), axis=1)]
    assert np.max(np.abs(lo - nearest) / np.abs(nearest)) < 1e-8
