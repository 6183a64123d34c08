"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs the reference checkout, which does not
travel to the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference's own raleigh.algebra.dense_numpy.Vectors / Matrix,
raleigh.core.solver and raleigh.interfaces.pca (unmodified; the only shim is
the SciPy>=1.14 `eigh(turbo=)` keyword drop) and records

  algebra_<tag>.npz   inputs + outputs of every Vectors/Matrix method
                      (the tests_algebra.py / tests_matrix.py call list)
  solver.npz          known-answer runs of the core solver:
                      * examples/core_solver.py:65-71 doctest (58 iterations)
                      * lap3d 12^3, 6 smallest, with/without Jacobi
                      * C3-like diagonally dominant SPD, Jacobi, block 8
  pca.npz             pca() runs (interfaces/pca.py:92-133 doctest + a small twin)

tests/test_oracle.py pins oracle/ against these; tests/test_parity_gpu.py pins
the CUDA backend against the same files.
"""
import os
import sys

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

import raleigh.core.solver as rsolver  # noqa: E402


class _SlaShim:
    def __getattr__(self, name):
        return getattr(sla, name)

    def eigh(self, *a, turbo=None, **k):
        return sla.eigh(*a, **k)


rsolver.sla = _SlaShim()

from raleigh.algebra.dense_numpy import Vectors, Matrix  # noqa: E402
from raleigh.core.solver import Problem, Solver, Options, DefaultConvergenceCriteria  # noqa: E402


def algebra_case(tag, n, nv, dtype, seed):
    rng = np.random.RandomState(seed)
    u = rng.randn(nv, n).astype(dtype)
    v = rng.randn(nv, n).astype(dtype)
    k = max(1, nv // 2)
    q = rng.randn(nv, k).astype(dtype)          # multiply: (self.nvec, out.nvec)
    p = rng.randn(k, nv).astype(dtype)          # add: (other.nvec, self.nvec)
    s = rng.rand(nv).astype(dtype) + 0.5
    s0 = s.copy()
    s0[::3] = 0.0
    M = max(3, n // 3)
    A = rng.randn(M, n).astype(dtype)
    ind = rng.permutation(nv)[:k]
    out = dict(u=u, v=v, q=q, p=p, s=s, s0=s0, A=A, ind=ind)

    U = Vectors(u.copy())
    V = Vectors(v.copy())
    out['dot'] = U.dot(V)
    out['dots'] = U.dots(V)
    out['dots_t'] = U.dots(V, transp=True)
    # windows: self = u[1:1+k], other = v[2:2+k-?]
    U.select(k, 1)
    V.select(max(1, k - 1), 2) if nv >= k + 2 else V.select(max(1, k - 1), 0)
    out['win_other'] = np.array(V.selected())
    out['dot_win'] = U.dot(V)
    U.select_all()
    V.select_all()
    W = Vectors(n, k, dtype)
    U.multiply(q, W)
    out['multiply'] = W.data().copy()
    V2 = Vectors(v.copy())
    W.select(k)
    V2.add(W, -0.5, p)
    out['add_q'] = V2.data().copy()
    V2 = Vectors(v.copy())
    V2.add(U, 2.0)
    out['add_s'] = V2.data().copy()
    V2 = Vectors(v.copy())
    V2.add(U, s)
    out['add_diag'] = V2.data().copy()
    V2 = Vectors(v.copy())
    V2.scale(s0)
    out['scale_div'] = V2.data().copy()
    V2 = Vectors(v.copy())
    V2.scale(s0, multiply=True)
    out['scale_mul'] = V2.data().copy()
    V2 = Vectors(v.copy())
    V2.select(k, nv - k)
    U.copy(V2, ind)
    V2.select_all()
    out['copy_ind'] = V2.data().copy()
    V2 = Vectors(v.copy())
    U.select(k, 1)
    V2.select(k, nv - k)
    U.copy(V2)
    V2.select_all()
    U.select_all()
    out['copy_win'] = V2.data().copy()
    # svd: sigma is unique; vectors only up to sign -> store sigma and the two
    # invariants the reference's own test prints (tests_algebra.py:330-341)
    W2 = Vectors(u.copy())
    sigma, qq = W2.svd()
    out['svd_sigma'] = sigma
    recon = (qq.conj() * sigma[None, :]) @ W2.data()
    out['svd_recon_err'] = np.array(np.linalg.norm(recon - u) / np.linalg.norm(u))
    # orthogonalize against an orthonormal set
    qmat, _ = np.linalg.qr(rng.randn(n, k).astype(dtype))
    onb = np.ascontiguousarray(qmat.T)
    out['onb'] = onb
    X = Vectors(v.copy())
    Qv = X.orthogonalize(Vectors(onb.copy()))
    out['orth_q'] = Qv.data().copy()
    out['orth_x'] = X.data().copy()
    # append
    X = Vectors(u.copy())
    Y = Vectors(v.copy())
    Y.select(k, 1)
    X.append(Y)
    out['append0'] = X.data().copy()
    X = Vectors(u.copy())
    X.append(Vectors(v.copy()), axis=1)
    out['append1'] = X.data().copy()
    # reference()+zero() on the upper half (tests_algebra.py:399-408)
    X = Vectors(u.copy())
    Z = X.reference()
    Z.select(nv // 2, nv // 2)
    Z.zero()
    out['ref_zero'] = X.data().copy()
    # Matrix
    Aop = Matrix(A.copy())
    x = Vectors(u.copy())
    y = Vectors(M, nv, dtype)
    Aop.apply(x, y)
    out['apply'] = y.data().copy()
    z = Vectors(n, nv, dtype)
    Aop.apply(y, z, transp=True)
    out['apply_t'] = z.data().copy()
    out['mdots'] = Aop.dots()
    np.savez_compressed(os.path.join(HERE, 'algebra_%s.npz' % tag), **out)
    print('algebra', tag, 'ok')


class CsrOperator:
    """SciPy stand-in for sparse_mkl.SparseSymmetricMatrix.apply (MKL absent)."""

    def __init__(self, A):
        u = sp.triu(A, format='csr')
        u.sort_indices()
        self.full = (u + sp.triu(u, k=1).T).tocsr()

    def apply(self, x, y):
        y.data()[...] = (self.full @ x.data().T).T


class JacobiOperator:
    def __init__(self, A):
        self.d = 1.0 / A.diagonal()

    def apply(self, x, y):
        y.data()[...] = x.data() * self.d[None, :]


def run_solver(A_op, n, dtype, which, tol, block, T=None, crit='k eigenvector error', seed=1,
               max_iter=1000):
    np.random.seed(seed)
    opt = Options()
    opt.block_size = block
    opt.max_iter = max_iter
    opt.convergence_criteria = DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance(crit, tol)
    v = Vectors(n, data_type=dtype)
    solver = Solver(Problem(v, A_op))
    if T is not None:
        solver.set_preconditioner(T)
    status = solver.solve(v, opt, which=which)
    return status, solver.iteration, np.array(solver.eigenvalues), v.data().copy()


def spd_c3_like(n, seed=0):
    """Small twin of BASELINE config 3: banded symmetric, strictly diagonally
    dominant, strongly varying diagonal (so Jacobi is a real preconditioner)."""
    rng = np.random.default_rng(seed)
    offs = [1, 2, 3, 7, 19, 20, 21]
    diags = [-rng.uniform(0.1, 1.0, n - o) for o in offs]
    L = sp.diags(diags, offs, shape=(n, n), format='csr')
    S = L + L.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(0.01, 1.0, n) * np.linspace(1, 100, n)
    return (S + sp.diags(d)).tocsr()


def solver_cases():
    out = {}
    # (1) examples/core_solver.py doctest: diag(1..100), 6 left, tol 1e-8 on
    # 'eigenvector error', auto block
    n = 100
    a = np.arange(1, n + 1).astype(np.float64)
    st, it, lmd, _ = run_solver(Matrix(np.diag(a)), n, np.float64, (6, 0), 1e-8, -1,
                                crit='eigenvector error', max_iter=-1)
    assert it == 58, it
    out['diag_iter'] = np.array(it)
    out['diag_lmd'] = lmd
    # (2) 12^3 Laplacian (reference generator), 6 smallest
    from raleigh.examples.laplace import lap3d
    L = lap3d(12, 12, 12, 1.0, 1.0, 1.0)
    n = L.shape[0]
    st, it, lmd, x = run_solver(CsrOperator(L), n, np.float64, (6, 0), 1e-6, 8)
    out['lap_status'], out['lap_iter'], out['lap_lmd'] = np.array(st), np.array(it), lmd
    st, it, lmd, x = run_solver(CsrOperator(L.astype(np.float32)), n, np.float32, (6, 0), 1e-3, 8)
    out['lap32_status'], out['lap32_iter'], out['lap32_lmd'] = np.array(st), np.array(it), lmd
    # (3) C3-like with Jacobi
    A = spd_c3_like(3000)
    st, it, lmd, x = run_solver(CsrOperator(A), 3000, np.float64, (5, 0), 1e-6, 8, T=JacobiOperator(A))
    out['spd_status'], out['spd_iter'], out['spd_lmd'] = np.array(st), np.array(it), lmd
    res = A @ x.T - x.T * lmd[None, :]
    out['spd_resnorm'] = np.linalg.norm(res, axis=0)
    np.savez_compressed(os.path.join(HERE, 'solver.npz'), **out)
    print('solver cases ok:', {k: (v.tolist() if v.size < 8 else v.shape) for k, v in out.items()})


def pca_cases():
    from raleigh.interfaces.pca import pca, pca_error
    from raleigh.examples.pca.generate_matrix import generate
    out = {}
    # small twin (kept small enough for the CPU test-suite)
    np.random.seed(1)
    A, sigma, u, v = generate(600, 400, 200, pca=True)
    out['small_sigma'] = sigma[:64]
    mean, trans, comps = pca(A, npc=40, opt=Options())
    out['small_npc40_err'] = np.array(pca_error(A, mean, trans, comps))
    out['small_npc40_ncomp'] = np.array(comps.shape[0])
    out['small_npc40_sv'] = np.linalg.norm(trans, axis=0)
    out['small_mean'] = mean
    mean, trans, comps = pca(A, tol=0.1, opt=Options())
    out['small_tol_err'] = np.array(pca_error(A, mean, trans, comps))
    out['small_tol_ncomp'] = np.array(comps.shape[0])
    mean, trans, comps = pca(A, batch_size=200, tol=0.1, opt=Options())
    out['small_inc_err'] = np.array(pca_error(A, mean, trans, comps))
    out['small_inc_ncomp'] = np.array(comps.shape[0])
    # the doctest itself (interfaces/pca.py:92-133)
    np.random.seed(1)
    A, sigma, u, v = generate(3000, 2000, 1000, pca=True)
    mean, trans, comps = pca(A, npc=300, opt=Options())
    out['doc_npc300_err'] = np.array(pca_error(A, mean, trans, comps))
    out['doc_npc300_ncomp'] = np.array(comps.shape[0])
    mean, trans, comps = pca(A, tol=0.05, opt=Options())
    out['doc_tol_err'] = np.array(pca_error(A, mean, trans, comps))
    out['doc_tol_ncomp'] = np.array(comps.shape[0])
    mean, trans, comps = pca(A, batch_size=1000, tol=0.05, opt=Options())
    out['doc_inc_err'] = np.array(pca_error(A, mean, trans, comps))
    out['doc_inc_ncomp'] = np.array(comps.shape[0])
    np.savez_compressed(os.path.join(HERE, 'pca.npz'), **out)
    print('pca cases ok:', {k: v.tolist() for k, v in out.items() if v.size < 4})


if __name__ == '__main__':
    algebra_case('d_300x12', 300, 12, np.float64, 11)
    algebra_case('s_257x7', 257, 7, np.float32, 12)
    algebra_case('d_64x5', 64, 5, np.float64, 13)
    solver_cases()
    pca_cases()
