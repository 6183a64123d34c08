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
    assert np.max(np.abs(sv[lead] - g['small_npc40_sv'][lead]) / g['small_npc40_sv'][lead]) < 1e-5
    assert np.max(np.abs(mean - g['small_mean'])) < 1e-5
    assert np.max(np.abs(comps @ comps.T - np.eye(comps.shape[0]))) < 1e-3
    mean, trans, comps = pca(A, tol=0.1, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert ef <= 0.1 and abs(comps.shape[0] - int(g['small_tol_ncomp'])) <= 3
    mean, trans, comps = pca(A, batch_size=200, tol=0.1, arch='gpu!', opt=ref.Options())
    em, ef = pca_error(A, mean, trans, comps)
    assert ef <= 0.1 + 1e-3


def test_incremental_pca_left_factor_stays_on_the_device(gpu_backend, ref):
    """lra.update (lra.py:287-290) grows the left factor through data() -> numpy.concatenate -> new_vectors();
    compat keeps those blocks on the device (vectors.DeviceData).  Same result as the host round trip, bit for bit;
    with the threshold at zero every data() inside update() goes through the stand-in, which then has to behave
    like the ndarray it replaces."""
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    from raleigh_b200 import compat, vectors
    np.random.seed(1)
    A, sigma, u, v = generate(900, 400, 200, pca=True)
    results = []
    taken = []
    orig_take = vectors.DeviceData.take

    def counting_take(self):
        taken.append(1)
        return orig_take(self)

    vectors.DeviceData.take = counting_take
    saved = compat.LAZY_UPDATE_BYTES
    try:
        for threshold in (1 << 60, 0, 64 << 10):
            compat.LAZY_UPDATE_BYTES = threshold
            del taken[:]
            np.random.seed(7)
            mean, trans, comps = pca(A, batch_size=300, tol=0.1, arch='gpu!', opt=ref.Options())
            results.append((mean, trans, comps, len(taken)))
    finally:
        compat.LAZY_UPDATE_BYTES = saved
        vectors.DeviceData.take = orig_take
    assert results[0][3] == 0 and results[1][3] >= 2 and results[2][3] >= 2     # two updates, one adoption each
    for mean, trans, comps, _ in results[1:]:
        assert np.array_equal(mean, results[0][0])
        assert np.array_equal(trans, results[0][1])
        assert np.array_equal(comps, results[0][2])
    em, ef = pca_error(A, *results[1][:3])
    assert ef <= 0.1 + 1e-3


def test_incremental_pca_chunk_prefetch(gpu_backend, ref):
    """lra.icompute with the next chunk uploaded in the background (vectors._ChunkPrefetch, active inside compat's
    hooked icompute): every chunk after the first is adopted from the prefetch, results identical to blocking uploads."""
    from raleigh.interfaces.pca import pca
    from raleigh.examples.pca.generate_matrix import generate
    from raleigh_b200 import vectors
    np.random.seed(1)
    A, sigma, u, v = generate(1000, 400, 200, pca=True)
    pf = vectors._chunk_prefetch
    saved_min = vectors.CHUNK_PREFETCH_MIN_BYTES
    orig_claim, orig_start = pf.claim, pf.start_next
    hits = []

    def counting_claim(a, ld_bytes):
        got = orig_claim(a, ld_bytes)
        hits.append(got is not None)
        return got

    try:
        vectors.CHUNK_PREFETCH_MIN_BYTES = 0
        pf.claim = counting_claim
        np.random.seed(7)
        r1 = pca(A, batch_size=300, tol=0.1, arch='gpu!', opt=ref.Options())
        assert hits == [False, True, True, True], hits          # 300 + 300 + 300 + 100 rows
        pf.start_next = lambda a, ld_bytes: None
        del hits[:]
        np.random.seed(7)
        r0 = pca(A, batch_size=300, tol=0.1, arch='gpu!', opt=ref.Options())
        assert not any(hits)
    finally:
        vectors.CHUNK_PREFETCH_MIN_BYTES = saved_min
        pf.claim, pf.start_next = orig_claim, orig_start
        pf.drop()
    for a, b in zip(r0, r1):
        assert np.array_equal(a, b)
    assert not vectors.CHUNK_PREFETCH


def test_incremental_pca_from_npy_memory_map(gpu_backend, ref, tmp_path):
    """examples/pca/incremental_pca.py:43 streams `numpy.load(path, mmap_mode='r')` through pca(batch_size=...):
    chunks of the memory map are uploaded (and prefetched) like chunks of an array; same result."""
    from raleigh.interfaces.pca import pca
    from raleigh.examples.pca.generate_matrix import generate
    from raleigh_b200 import vectors, io
    np.random.seed(1)
    A, sigma, u, v = generate(900, 400, 200, pca=True)
    path = str(tmp_path / 'data.npy')
    np.save(path, A)
    data = io.open_npy(path)
    pf = vectors._chunk_prefetch
    saved_min = vectors.CHUNK_PREFETCH_MIN_BYTES
    orig_claim = pf.claim
    hits = []

    def counting_claim(a, ld_bytes):
        got = orig_claim(a, ld_bytes)
        hits.append(got is not None)
        return got

    try:
        vectors.CHUNK_PREFETCH_MIN_BYTES = 0
        pf.claim = counting_claim
        np.random.seed(7)
        r1 = pca(data, batch_size=300, tol=0.1, arch='gpu!', opt=ref.Options())
        assert hits == [False, True, True], hits
        np.random.seed(7)
        r0 = pca(A, batch_size=300, tol=0.1, arch='gpu!', opt=ref.Options())
    finally:
        vectors.CHUNK_PREFETCH_MIN_BYTES = saved_min
        pf.claim = orig_claim
        pf.drop()
    for a, b in zip(r0, r1):
        assert np.array_equal(a, b)


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


# ---- device-resident driver (raleigh_b200/jcg.py) against the reference's own main loop ----------
def _both_paths(gpu_backend, fn):
    out = {}
    for name, on in (('verbatim', False), ('device', True)):
        gpu_backend.use_device_solver(on)
        try:
            out[name] = fn()
        finally:
            gpu_backend.use_device_solver(True)
    return out['verbatim'], out['device']


SOLVER_CASES = {
    # name: (grid, which, tol, block, jacobi, expected iterations of the reference on its NumPy algebra)
    'c1_32cube_block16': (32, (10, 0), 1e-6, 16, False, 155),          # BASELINE config 1
    'lap12_block8': (12, (6, 0), 1e-6, 8, False, None),
    'lap16_both_ends': (16, (4, 3), 1e-6, 12, False, None),
    'lap24_tma_gram': (24, (6, 0), 1e-6, 16, False, None),             # n >= 8192: TMA-fed Gram kernels
    'lap20_largest': (20, 5, 1e-6, 8, False, None),
}


@pytest.mark.parametrize('case', sorted(SOLVER_CASES))
def test_device_solver_matches_verbatim_solver(gpu_backend, ref, case):
    N, which, tol, block, jac, expected = SOLVER_CASES[case]
    L = K.lap3d_csr(N, N, N)
    op = gpu_backend.SparseSymmetricMatrix(L)
    (st0, it0, lmd0, v0), (st1, it1, lmd1, v1) = _both_paths(
        gpu_backend, lambda: _solve(ref, gpu_backend.Vectors, op, L.shape[0], np.float64, which, tol, block))
    assert st0 == 0 and st1 == 0
    assert len(lmd0) == len(lmd1)
    assert np.max(np.abs(np.sort(lmd1) - np.sort(lmd0)) / np.abs(np.sort(lmd0))) < 1e-10, (it0, it1)
    assert abs(it1 - it0) <= max(2, it0 // 20), (it0, it1)      # same algorithm, different rounding
    if expected is not None:
        assert abs(it1 - expected) <= max(2, expected // 20), (it1, expected)
    x = v1.data()
    lam = np.array(lmd1)
    res = np.linalg.norm(L @ x.T - x.T * lam[None, :], axis=0)
    assert np.max(res / np.abs(lam)) < 1e-4


@pytest.mark.parametrize('N,which,block,jac', [(12, (6, 0), 8, False), (16, (4, 3), 12, False), (22, (6, 0), 16, True)])
def test_device_solver_generalised_problem(gpu_backend, ref, N, which, block, jac):
    """A x = lambda B x with a sparse SPD mass matrix (solver.py 'gen' branches): the device-resident driver against
    the reference's own loop on the same backend, and against the pencil's eigenvalues."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rs = ref
    L = K.lap3d_csr(N, N, N)
    n = L.shape[0]
    rng = np.random.RandomState(0)
    o = 0.1 * rng.rand(n - 1)
    M = (sp.diags(1.0 + rng.rand(n)) + sp.diags(o, 1) + sp.diags(o, -1)).tocsr()
    opA, opB = gpu_backend.SparseSymmetricMatrix(L), gpu_backend.SparseSymmetricMatrix(M)
    T = gpu_backend.Operator(gpu_backend.DiagonalPreconditioner(L)) if jac else None

    def run():
        np.random.seed(1)
        opt = rs.Options()
        opt.block_size = block
        opt.max_iter = 1000
        opt.convergence_criteria = rs.DefaultConvergenceCriteria()
        opt.convergence_criteria.set_error_tolerance('k eigenvector error', 1e-6)
        v = gpu_backend.Vectors(n, data_type=np.float64)
        problem = rs.Problem(v, opA, opB)
        assert problem.type() == 'g'
        solver = rs.Solver(problem)
        if T is not None:
            solver.set_preconditioner(T)
        status = solver.solve(v, opt, which=which)
        return status, solver.iteration, np.array(solver.eigenvalues), v, solver

    (st0, it0, lmd0, v0, _), (st1, it1, lmd1, v1, sol) = _both_paths(gpu_backend, run)
    assert st0 == 0 and st1 == 0 and len(lmd0) == len(lmd1)
    assert np.max(np.abs(np.sort(lmd1) - np.sort(lmd0)) / np.abs(np.sort(lmd0))) < 1e-10, (it0, it1)
    assert abs(it1 - it0) <= max(2, it0 // 20), (it0, it1)
    x = v1.data().T
    assert np.max(np.abs(x.T @ (M @ x) - np.eye(x.shape[1]))) < 1e-6
    res = np.linalg.norm(L @ x - (M @ x) * lmd1[None, :], axis=0)
    assert np.max(res / np.abs(lmd1)) < 1e-3
    assert np.allclose(sol.eigenvectors_im.data(), (M @ x).T, rtol=1e-9, atol=1e-9)
    if N <= 16:
        exact = spla.eigsh(L.tocsc(), k=max(which[0], 1) + 2, M=M.tocsc(), sigma=0, which='LM', return_eigenvectors=False)
        left = np.sort(lmd1)[:which[0]]
        assert np.max(np.abs(left - np.sort(exact)[:which[0]]) / left) < 1e-8


def test_device_solver_partial_hevp_with_mass_matrix(gpu_backend, ref):
    """partial_hevp(A, B, T=...) (partial_hevp.py:202-224): `Problem(v, A, B, 'gen')` -- whose fourth argument selects
    the PRODUCT form A B x = lambda x (solver.py:241-250) -- through the device-resident driver and through the
    reference's own loop: same eigenvalues, which are eigenvalues of A B."""
    import scipy.sparse as sp
    from raleigh.interfaces.partial_hevp import partial_hevp
    L = K.lap3d_csr(12, 12, 12)
    n = L.shape[0]
    rng = np.random.RandomState(0)
    o = 0.1 * rng.rand(n - 1)
    M = (sp.diags(1.0 + rng.rand(n)) + sp.diags(o, 1) + sp.diags(o, -1)).tocsr()
    T = gpu_backend.DiagonalPreconditioner(L)

    def run():
        np.random.seed(1)
        opt = ref.Options()
        opt.block_size = 8
        opt.max_iter = 1000
        return partial_hevp(L, B=M, T=T, which=5, tol=1e-6, verb=-1, opt=opt)

    (lmd0, x0, st0), (lmd1, x1, st1) = _both_paths(gpu_backend, run)
    assert st0 == 0 and st1 == 0 and len(lmd0) >= 5 and len(lmd1) >= 5      # every converged pair is returned
    k = min(len(lmd0), len(lmd1))
    assert np.max(np.abs(lmd1[:k] - lmd0[:k]) / np.abs(lmd0[:k])) < 1e-10
    res = L @ (M @ x1) - x1 * lmd1[None, :]
    assert np.max(np.linalg.norm(res, axis=0) / np.abs(lmd1)) < 1e-3


def test_device_solver_jacobi_preconditioned_c3_like(gpu_backend, ref):
    """Small twin of BASELINE config 3 (n >= 8192 so that the TMA Gram runs): partial_hevp with the
    Jacobi preconditioner, block 32, device-resident driver against the verbatim path."""
    from raleigh.interfaces.partial_hevp import partial_hevp
    A = spd_c3_like(20000)

    def run():
        np.random.seed(1)
        opt = ref.Options()
        opt.block_size = 32
        opt.max_iter = 1000
        T = gpu_backend.DiagonalPreconditioner(A)
        lmd, x, status = partial_hevp(A, T=T, which=10, tol=1e-6, verb=-1, opt=opt)
        return lmd, x, status

    (l0, x0, s0), (l1, x1, s1) = _both_paths(gpu_backend, run)
    assert s0 == 0 and s1 == 0
    assert np.max(np.abs(l1 - l0) / l0) < 1e-10
    res = np.linalg.norm(A @ x1 - x1 * l1[None, :], axis=0)
    assert np.max(res / l1) < 1e-4


def test_device_solver_pca_matches_verbatim(gpu_backend, ref):
    """pca(npc=...) in fp32: singular values of the device-resident driver against the verbatim path
    to 1e-5 on the converged leading components (north_star tolerance)."""
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    np.random.seed(1)
    A, sigma, u, v = generate(3000, 2000, 1000, pca=True)

    def run():
        np.random.seed(1)
        mean, trans, comps = pca(A, npc=300, arch='gpu!', opt=ref.Options())
        return mean, trans, comps

    (m0, t0, c0), (m1, t1, c1) = _both_paths(gpu_backend, run)
    assert c1.shape == c0.shape == (300, 2000)
    sv0, sv1 = np.linalg.norm(t0, axis=0), np.linalg.norm(t1, axis=0)
    lead = slice(0, 100)
    assert np.max(np.abs(sv1[lead] - sv0[lead]) / sv0[lead]) < 1e-5
    e0, e1 = pca_error(A, m0, t0, c0), pca_error(A, m1, t1, c1)
    assert abs(e0[0] - e1[0]) < 2e-3 and abs(e0[1] - e1[1]) < 2e-3
    assert np.max(np.abs(c1 @ c1.T - np.eye(300))) < 1e-3


@pytest.mark.parametrize('dtype,nsv,cond', [(np.float32, 60, 1e2), (np.float32, 200, 1e3), (np.float64, 100, 1e5),
                                            (np.float32, 1000, 2e2)])
def test_device_finalize_svd_against_reference(gpu_backend, ref, dtype, nsv, cond):
    """psvd.finalize_svd (one device eigen-decomposition) against the reference's
    PartialSVD._finalize_svd (host cholesky / svd / inv, partial_svd.py:163-235) on the same inputs."""
    from raleigh.interfaces.partial_svd import PartialSVD
    from raleigh_b200 import psvd
    rng = np.random.RandomState(nsv)
    m, n = 3 * nsv + 17, 2 * nsv + 5
    sig = np.logspace(0, -np.log10(cond), nsv)
    ua, _ = np.linalg.qr(rng.randn(m, nsv))
    va, _ = np.linalg.qr(rng.randn(n, nsv))
    # v: slightly rotated right singular vectors (what the solver delivers at svtol), Av = A v
    rot, _ = np.linalg.qr(np.eye(nsv) + 1e-3 * rng.randn(nsv, nsv))
    v0 = (va @ rot).T.astype(dtype)                      # (nsv, n)
    A = (ua * sig) @ va.T
    av0 = (A @ v0.T.astype(np.float64)).T.astype(dtype)  # (nsv, m)
    V, AV = gpu_backend.Vectors(v0.copy()), gpu_backend.Vectors(av0.copy())
    u, sigma, v = psvd.finalize_svd(V, AV, 1e-3)
    Vr, AVr = gpu_backend.Vectors(v0.copy()), gpu_backend.Vectors(av0.copy())
    gpu_backend.use_device_solver(False)
    try:
        ur, sigr, vr = PartialSVD._finalize_svd(Vr, AVr, 1e-3)
    finally:
        gpu_backend.use_device_solver(True)
    tol = 1e-5 if dtype is np.float32 else 1e-10
    assert sigma.dtype == dtype
    assert np.max(np.abs(sigma - sigr) / sigr) < tol * 20            # relative, down to the smallest one
    assert np.max(np.abs(sigma.astype(np.float64) - sig) / sig) < (2e-4 if dtype is np.float32 else 1e-6)
    U, Vv = u.data().astype(np.float64), v.data().astype(np.float64)
    ortho = 2e-5 if dtype is np.float32 else 1e-11
    Ur = ur.data().astype(np.float64)
    ref_ortho = np.max(np.abs(Ur @ Ur.T - np.eye(nsv)))
    assert np.max(np.abs(U @ U.T - np.eye(nsv))) < max(ortho * 10, 10 * ref_ortho)      # eps * cond for both routes
    assert np.max(np.abs(Vv @ Vv.T - np.eye(nsv))) < ortho * 10
    # A v' = u diag(sigma)
    lhs = A @ Vv.T
    rhs = U.T * sigma.astype(np.float64)[None, :]
    assert np.max(np.abs(lhs - rhs)) < (5e-6 if dtype is np.float32 else 1e-10)


def test_device_pca_wide_matrix_uses_finalize(gpu_backend, ref):
    """m < n: pca() sets ortho = svtol and goes through _finalize_svd (pca.py:147-148) -- the config-2 route."""
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    np.random.seed(1)
    A, sigma, u, v = generate(500, 900, 300, pca=True)

    def run():
        np.random.seed(1)
        return pca(A, npc=60, arch='gpu!', opt=ref.Options())

    (m0, t0, c0), (m1, t1, c1) = _both_paths(gpu_backend, run)
    assert c0.shape == c1.shape
    sv0, sv1 = np.linalg.norm(t0, axis=0), np.linalg.norm(t1, axis=0)
    lead = slice(0, 30)
    # two fp32 computations of the same quantities: north_star's 1e-5 plus their own rounding
    assert np.max(np.abs(sv1[lead] - sv0[lead]) / sv0[lead]) < 2e-5
    # and both sit on the exact singular values of the centred matrix to the solver's own tolerance
    # (svtol = 1e-3 on the vectors: measured 6e-6 on the first value, 6e-4 on the tenth, for either path)
    exact = np.linalg.svd(A.astype(np.float64) - A.astype(np.float64).mean(axis=0), compute_uv=False)
    assert np.max(np.abs(sv1[:10] - exact[:10]) / exact[:10]) < 2e-3
    assert np.max(np.abs(sv0[:10] - exact[:10]) / exact[:10]) < 2e-3
    e0, e1 = pca_error(A, m0, t0, c0), pca_error(A, m1, t1, c1)
    assert abs(e0[0] - e1[0]) < 2e-3 and abs(e0[1] - e1[1]) < 2e-3
    assert np.max(np.abs(c1 @ c1.T - np.eye(c1.shape[0]))) < 1e-4
