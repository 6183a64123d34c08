"""End-to-end parity on the GPU: the reference's UNMODIFIED core solver,
partial_hevp and pca running on the raleigh_b200 backend, against the golden
results the reference produced on its own NumPy algebra
(tests/golden/make_golden.py) -- eigenvalues within 1e-10 (fp64) / 1e-5 (fp32),
residuals below the solver tolerance, iteration counts side by side."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

from oracle import algebra_np as K
from tests_common import spd_c3_like

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ref(gpu_backend, ref_root):
    gpu_backend.install(ref_root)
    import raleigh.core.solver as rs
    return rs


def _solve(rs, Vectors, op, n, dtype, which, tol, block, T=None, crit='k eigenvector error', max_iter=1000):
    np.random.seed(1)
    opt = rs.Options()
    opt.block_size = block
    opt.max_iter = max_iter
    opt.convergence_criteria = rs.DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance(crit, tol)
    v = Vectors(n, data_type=dtype)
    solver = rs.Solver(rs.Problem(v, op))
    if T is not None:
        solver.set_preconditioner(T)
    status = solver.solve(v, opt, which=which)
    return status, solver.iteration, np.array(solver.eigenvalues), v


def test_core_solver_doctest(gpu_backend, ref):
    """examples/core_solver.py:65-71: 'after 58 iterations, 6 converged eigenvalues are: [1. ... 6.]'"""
    a = np.arange(1, 101).astype(np.float64)
    st, it, lmd, v = _solve(ref, gpu_backend.Vectors, gpu_backend.Matrix(np.diag(a)), 100, np.float64, (6, 0),
                            1e-8, -1, crit='eigenvector error', max_iter=-1)
    assert v.nvec() == 6
    assert np.allclose(lmd, [1, 2, 3, 4, 5, 6], atol=1e-10)
    assert abs(it - 58) <= 3, it          # reference: 58; identical host RNG stream


def test_laplacian_fp64(gpu_backend, ref):
    g = np.load(os.path.join(GOLDEN, 'solver.npz'))
    L = K.lap3d_csr(12, 12, 12)
    op = gpu_backend.SparseSymmetricMatrix(L)
    st, it, lmd, v = _solve(ref, gpu_backend.Vectors, op, L.shape[0], np.float64, (6, 0), 1e-6, 8)
    assert st == 0
    assert np.max(np.abs(lmd - g['lap_lmd']) / g['lap_lmd']) < 1e-10
    assert abs(it - int(g['lap_iter'])) <= max(3, int(g['lap_iter']) // 10), (it, int(g['lap_iter']))
    x = v.data()
    res = np.linalg.norm(L @ x.T - x.T * lmd[None, :], axis=0)
    assert np.max(res / lmd) < 1e-5


def test_laplacian_fp32(gpu_backend, ref):
    g = np.load(os.path.join(GOLDEN, 'solver.npz'))
    L = K.lap3d_csr(12, 12, 12).astype(np.float32)
    op = gpu_backend.SparseSymmetricMatrix(L)
    st, it, lmd, v = _solve(ref, gpu_backend.Vectors, op, L.shape[0], np.float32, (6, 0), 1e-3, 8)
    assert st == 0
    exact = K.lap3d_eigenvalues(12, 12, 12)[:6]
    assert np.max(np.abs(np.sort(lmd) - exact) / exact) < 1e-5
    assert abs(it - int(g['lap32_iter'])) <= 8, (it, int(g['lap32_iter']))


def test_partial_hevp_with_jacobi(gpu_backend, ref):
    """partial_hevp(A, T=...) preconditioned branch (partial_hevp.py:202-224) verbatim."""
    g = np.load(os.path.join(GOLDEN, 'solver.npz'))
    from raleigh.interfaces.partial_hevp import partial_hevp
    A = spd_c3_like(3000)
    np.random.seed(1)
    opt = ref.Options()
    opt.block_size = 8
    opt.max_iter = 1000
    T = gpu_backend.DiagonalPreconditioner(A)
    lmd, x, status = partial_hevp(A, T=T, which=5, tol=1e-6, verb=-1, opt=opt)
    assert status == 0
    assert np.max(np.abs(lmd - g['spd_lmd']) / g['spd_lmd']) < 1e-10
    res = np.linalg.norm(A @ x - x * lmd[None, :], axis=0)
    assert np.max(res) < 10 * np.max(g['spd_resnorm']) + 1e-5


def test_pca_small(gpu_backend, ref):
    g = np.load(os.path.join(GOLDEN, 'pca.npz'))
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    np.random.seed(1)
    A, sigma, u, v = generate(600, 400, 200, pca=True)
    assert np.allclose(sigma[:64], g['small_sigma'])
    mean, trans, comps = pca(A, npc=40, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert comps.shape[0] == int(g['small_npc40_ncomp'])
    assert abs(em - g['small_npc40_err'][0]) < 2e-3 and abs(ef - g['small_npc40_err'][1]) < 2e-3
    sv = np.linalg.norm(trans, axis=0)
    lead = slice(0, 20)
    assert np.max(np.abs(sv[lead] - g['small_npc40_sv'][lead]) / g['small_npc40_sv'][lead]) < 1e-4
    assert np.max(np.abs(mean - g['small_mean'])) < 1e-5
    assert np.max(np.abs(comps @ comps.T - np.eye(comps.shape[0]))) < 1e-3
    mean, trans, comps = pca(A, tol=0.1, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert ef <= 0.1 and abs(comps.shape[0] - int(g['small_tol_ncomp'])) <= 3
    mean, trans, comps = pca(A, batch_size=200, tol=0.1, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert ef <= 0.1 + 1e-3


def test_pca_doctest(gpu_backend, ref):
    """interfaces/pca.py:92-133 known answers, arch='gpu!'."""
    g = np.load(os.path.join(GOLDEN, 'pca.npz'))
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    np.random.seed(1)
    A, sigma, u, v = generate(3000, 2000, 1000, pca=True)
    mean, trans, comps = pca(A, npc=300, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert comps.shape[0] == 300
    assert '%.0e %.0e' % (em, ef) == '5e-02 1e-01'
    assert abs(em - g['doc_npc300_err'][0]) < 1e-3 and abs(ef - g['doc_npc300_err'][1]) < 1e-3
    mean, trans, comps = pca(A, tol=0.05, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert '%.0e %.0e' % (em, ef) == '2e-02 4e-02'
    assert abs(comps.shape[0] - int(g['doc_tol_ncomp'])) <= 5


@pytest.mark.skipif(os.environ.get('RALEIGH_B200_LONG_TESTS', '0') != '1',
                    reason='written after the round-1 GPU budget was spent: first run is manual (RALEIGH_B200_LONG_TESTS=1)')
@pytest.mark.parametrize('block', [8, 16])
def test_laplacian_fp64_large_enough_for_the_tma_gram(gpu_backend, ref, block):
    """n = 24^3 = 13,824 >= 8192: every Vectors.dot of the solve goes through the TMA-fed Gram kernel
    (all tile shapes as windows shrink) and the SpMM runs with the L2 prefetch; eigenvalues against
    the analytic spectrum of the 7-point Laplacian, residuals below the solver tolerance."""
    L = K.lap3d_csr(24, 24, 24)
    n = L.shape[0]
    op = gpu_backend.SparseSymmetricMatrix(L)
    st, it, lmd, v = _solve(ref, gpu_backend.Vectors, op, n, np.float64, (6, 0), 1e-6, block)
    assert st == 0
    exact = K.lap3d_eigenvalues(24, 24, 24)[:6]
    assert np.max(np.abs(np.sort(lmd) - exact) / exact) < 1e-10, (it, lmd)
    x = v.data()
    res = np.linalg.norm(L @ x.T - x.T * lmd[None, :], axis=0)
    assert np.max(res / lmd) < 1e-4          # CPU oracle on the same problem: 7.7e-6 (block 8), 1.2e-5 (block 16)
