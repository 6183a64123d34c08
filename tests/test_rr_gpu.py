"""GPU parity of the device-resident Rayleigh-Ritz kernels (csrc/rr.cu, jacobi.cu,
rr_solve.cu) through the C ABI, each against a NumPy statement of the same
algorithm (oracle/jcg_engine_np.py) or against LAPACK where the result is unique."""
import ctypes

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import jcg_engine_np as E

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def rt(gpu_backend):
    """Tiny runtime: device fp64 matrices from host arrays and back."""
    import torch
    from raleigh_b200._lib import lib, check
    from raleigh_b200 import device as dev

    class RT:
        pass
    r = RT()
    r.lib, r.check, r.dev, r.torch = lib, check, dev, torch

    def up(a, dtype=torch.float64):
        return torch.from_numpy(np.ascontiguousarray(a)).to('cuda').to(dtype).contiguous()

    def zeros(*shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device='cuda')
    r.up, r.zeros = up, zeros
    r.st = dev.stream
    return r


def _spd(n, cond, seed):
    rng = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rng.randn(n, n))
    w = np.logspace(0, -np.log10(cond), n)
    return (q * w) @ q.T


@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (5, 7, 3), (64, 64, 16), (65, 130, 33), (256, 256, 256), (1000, 128, 1000)])
def test_small_gemm(rt, ta, tb, M, N, K):
    rng = np.random.RandomState(M + N + K)
    A = rng.randn(K, M) if ta else rng.randn(M, K)
    B = rng.randn(N, K) if tb else rng.randn(K, N)
    C = rng.randn(M, N)
    dA, dB, dC = rt.up(A), rt.up(B), rt.up(C)
    rt.check(rt.lib.rl_small_gemm(ta, tb, M, N, K, -1.5, dA.data_ptr(), A.shape[1], dB.data_ptr(), B.shape[1], 2.0,
                                  dC.data_ptr(), N, rt.st()))
    ref = -1.5 * ((A.T if ta else A) @ (B.T if tb else B)) + 2.0 * C
    assert np.max(np.abs(dC.cpu().numpy() - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))) * np.sqrt(K)


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('n,r', [(1, 1), (5, 3), (31, 40), (32, 32), (33, 65), (100, 100), (256, 256), (257, 130), (480, 64)])
def test_small_trsm(rt, mode, n, r):
    rng = np.random.RandomState(n * 7 + r)
    U = np.triu(rng.randn(n, n)) + np.diag(3.0 + rng.rand(n))
    B = rng.randn(n, r)
    ld = n + 3
    Up = np.zeros((n, ld)); Up[:, :n] = U
    dU, dB = rt.up(Up), rt.up(B)
    rt.check(rt.lib.rl_small_trsm(mode, dU.data_ptr(), ld, n, dB.data_ptr(), r, r, rt.st()))
    ref = sla.solve_triangular(U, B, trans=1) if mode == 0 else sla.solve_triangular(U, B)
    got = dB.cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))


def _run_chol(rt, A, k, eps):
    n = A.shape[0]
    ld = n + 1
    Ap = np.zeros((n, ld)); Ap[:, :n] = A
    dA, dA0 = rt.up(Ap), rt.zeros(n, ld)
    ind = rt.zeros(n, dtype=rt.torch.int32)
    info = rt.zeros(4, dtype=rt.torch.int32)
    rt.check(rt.lib.rl_rr_piv_chol(dA.data_ptr(), dA0.data_ptr(), ld, n, k, eps, ind.data_ptr(), info.data_ptr(), rt.st()))
    return dA.cpu().numpy()[:, :n], ind.cpu().numpy(), info.cpu().numpy()


@pytest.mark.parametrize('n,k', [(1, 0), (8, 0), (8, 8), (16, 8), (33, 16), (64, 32), (130, 65), (200, 70), (256, 128), (300, 0)])
def test_piv_chol_full_rank(rt, n, k):
    A = _spd(n, 1e3, n + k)
    U, ind, info = _run_chol(rt, A, k, 1e-8)
    Ar = A.copy()
    ind_r, dropped_r, status = E.piv_chol(Ar, n, k, 1e-8)
    assert info[0] == dropped_r == 0 and info[1] == 0
    assert np.array_equal(ind, ind_r)
    assert np.max(np.abs(np.triu(U) - np.triu(Ar))) <= 1e-10
    P = A[np.ix_(ind, ind)]
    assert np.max(np.abs(np.triu(U).T @ np.triu(U) - P)) <= 1e-12 * n
    assert np.max(np.abs(np.tril(U, -1))) == 0.0


@pytest.mark.parametrize('n,k,rank_y,noise', [(40, 20, 10, 0.0), (40, 20, 10, 1e-7), (140, 70, 50, 1e-6), (256, 128, 100, 1e-5),
                                              (200, 0, 120, 1e-6), (96, 16, 79, 3e-5)])
def test_piv_chol_drops_like_the_host_rule(rt, n, k, rank_y, noise):
    """Gram matrix of [X, Y] with orthonormal X and Y spanning only rank_y directions (+ noise):
    the drop decisions (pivot <= eps, condition estimate <= eps, bisection) must match the NumPy
    statement of the same rule."""
    rng = np.random.RandomState(n + rank_y)
    N = 4 * n
    X, _ = np.linalg.qr(rng.randn(N, max(k, 1)))
    X = X[:, :k]
    Yb = rng.randn(N, rank_y)
    Y = Yb @ rng.randn(rank_y, n - k) + noise * rng.randn(N, n - k)
    if k:
        Y -= X @ (X.T @ Y)
    Y /= np.linalg.norm(Y, axis=0)
    V = np.concatenate((X, Y), axis=1)
    A = V.T @ V
    U, ind, info = _run_chol(rt, A, k, 1e-8)
    Ar = A.copy()
    ind_r, dropped_r, status = E.piv_chol(Ar, n, k, 1e-8)
    assert info[1] == 0 and status == 0
    assert abs(int(info[0]) - dropped_r) <= 1, (info, dropped_r)       # borderline estimates may differ by one
    assert info[0] >= n - k - rank_y - 1
    kept = n - int(info[0])
    if int(info[0]) == dropped_r and np.array_equal(ind[:kept], ind_r[:kept]):
        assert np.max(np.abs(np.triu(U)[:kept, :kept] - np.triu(Ar)[:kept, :kept])) <= 1e-6
    Uk = np.triu(U)[:kept, :kept]
    P = A[np.ix_(ind[:kept], ind[:kept])]
    assert np.max(np.abs(Uk.T @ Uk - P)) <= 1e-10
    assert np.all(U[kept:, :] == 0.0)


def _eig(rt, G, fn='rl_syevj_cluster'):
    n = G.shape[0]
    ld = n + 2
    Gp = np.zeros((n, ld)); Gp[:, :n] = G
    dG = rt.up(Gp)
    w, Q = rt.zeros(n), rt.zeros(n, n)
    wsb = rt.lib.rl_small_eigh_ws_bytes(n)
    ws = rt.zeros(wsb // 8 + 8)
    info = rt.zeros(4, dtype=rt.torch.int32)
    if fn == 'rl_syevj_cluster':
        rt.check(rt.lib.rl_syevj_cluster(dG.data_ptr(), ld, n, 0, 0.0, w.data_ptr(), Q.data_ptr(), n, ws.data_ptr(), wsb,
                                         info.data_ptr(), rt.st()))
    else:
        rt.check(getattr(rt.lib, fn)(dG.data_ptr(), ld, n, 0.0, w.data_ptr(), Q.data_ptr(), n, ws.data_ptr(), wsb,
                                     info.data_ptr(), rt.st()))
    return w.cpu().numpy(), Q.cpu().numpy(), info.cpu().numpy()


EIG_SIZES = [1, 2, 3, 4, 15, 16, 31, 32, 33, 64, 100, 128, 129, 200, 240, 256, 257, 320]


@pytest.mark.parametrize('n', EIG_SIZES)
@pytest.mark.parametrize('kind', ['random', 'spd', 'neardiag', 'clustered'])
def test_cluster_jacobi_eigh(rt, n, kind):
    rng = np.random.RandomState(n)
    if kind == 'random':
        G = rng.randn(n, n); G = G + G.T
    elif kind == 'spd':
        G = _spd(n, 1e6, n) * 1e3
    elif kind == 'neardiag':          # what the Rayleigh-Ritz step produces late in a solve
        G = np.diag(np.sort(rng.rand(n)) * 100) + 1e-5 * rng.randn(n, n); G = 0.5 * (G + G.T)
    else:
        q, _ = np.linalg.qr(rng.randn(n, n))
        G = (q * np.repeat(np.arange(1, n // 3 + 2), 3)[:n]) @ q.T
    w, Q, info = _eig(rt, G)
    wr = np.linalg.eigvalsh(G)
    scale = max(1.0, np.max(np.abs(wr)))
    assert info[1] == 1, info
    assert np.max(np.abs(w - wr)) <= 1e-13 * scale * n
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13 * n
    assert np.max(np.abs(G @ Q - Q * w[None, :])) <= 1e-13 * scale * n
    assert np.all(np.diff(w) >= 0)
    # same sweeps as the NumPy statement of the algorithm (+-1: summation order)
    if n <= 64:
        _, _, sweeps = E.jacobi_eigh(G.copy())
        assert abs(int(info[0]) - sweeps) <= 1, (info, sweeps)


@pytest.mark.parametrize('n', [321, 336, 337, 512, 513, 777, 1000, 1024])
@pytest.mark.parametrize('kind', ['random', 'neardiag', 'clustered'])
def test_ring_jacobi_eigh(rt, n, kind):
    """Orders beyond one cluster: the two-level ring tournament over the whole GPU (point-to-point block exchange
    through L2); the flat one-barrier-per-round kernel stays behind a knob and must agree."""
    rng = np.random.RandomState(n)
    if kind == 'random':
        G = rng.randn(n, n); G = G + G.T
    elif kind == 'neardiag':
        G = np.diag(np.sort(rng.rand(n)) * 100) + 1e-5 * rng.randn(n, n); G = 0.5 * (G + G.T)
    else:
        q, _ = np.linalg.qr(rng.randn(n, n))
        G = (q * np.repeat(np.arange(1, n // 3 + 2), 3)[:n]) @ q.T
    w, Q, info = _eig(rt, G)
    wr = np.linalg.eigvalsh(G)
    scale = max(1.0, np.max(np.abs(wr)))
    assert info[1] == 1, info
    assert np.max(np.abs(w - wr)) <= 1e-13 * scale * n
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) <= 1e-13 * n
    assert np.max(np.abs(G @ Q - Q * w[None, :])) <= 1e-13 * scale * n
    assert np.all(np.diff(w) >= 0)
    if n in (337, 1000) and kind == 'random':
        rt.lib.rl_debug_set_knob(14, 1)
        try:
            w2, Q2, info2 = _eig(rt, G)
        finally:
            rt.lib.rl_debug_set_knob(14, 0)
        assert info2[1] == 1
        assert np.max(np.abs(w2 - w)) <= 1e-13 * scale * n


@pytest.mark.parametrize('n', [1, 7, 64, 130, 321, 500, 1000])
@pytest.mark.parametrize('cond', [1e2, 1e8])
def test_potrf_and_jacobi_on_the_factor(rt, n, cond):
    """G = U^T U (blocked Cholesky on the device), then eigh from the factor: every eigenvalue to relative
    accuracy even at condition 1e8 (the plain symmetric solver is only absolutely accurate)."""
    rng = np.random.RandomState(n)
    q, _ = np.linalg.qr(rng.randn(n, n))
    lam = np.logspace(0, -np.log10(cond), n) if n > 1 else np.ones(1)
    # graded matrix: well conditioned after diagonal scaling, like the Gram matrix of nearly singular vectors
    d = np.sqrt(lam)
    S = np.eye(n) + 1e-3 * (q + q.T)
    G = d[:, None] * S * d[None, :]
    dU = rt.up(G)
    info = rt.zeros(4, dtype=rt.torch.int32)
    rt.check(rt.lib.rl_small_potrf(dU.data_ptr(), n, n, info.data_ptr(), rt.st()))
    assert int(info[0]) == 0
    U = dU.cpu().numpy()
    assert np.max(np.abs(np.tril(U, -1))) == 0.0
    assert np.max(np.abs(U.T @ U - G) / np.sqrt(np.outer(np.diag(G), np.diag(G)))) < 1e-13 * max(n, 8)
    w, Q = rt.zeros(n), rt.zeros(n, n)
    wsb = rt.lib.rl_small_eigh_ws_bytes(n)
    ws = rt.zeros(wsb // 8 + 8)
    rt.check(rt.lib.rl_small_eigh_factor(dU.data_ptr(), n, n, 0.0, w.data_ptr(), Q.data_ptr(), n, ws.data_ptr(), wsb,
                                         info.data_ptr(), rt.st()))
    w, Q = w.cpu().numpy(), Q.cpu().numpy()
    assert int(info[1]) == 1, info.cpu().numpy()
    wr = np.linalg.eigvalsh(G.astype(np.longdouble).astype(np.float64))
    # relative accuracy against the eigenvalues of the scaled problem solved in extended precision
    Sr = np.linalg.eigvalsh(S)
    assert np.all(w > 0)
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) < 1e-13 * max(n, 8)
    resid = G @ Q - Q * w[None, :]
    assert np.max(np.abs(resid) / d[:, None]) / np.max(d) < 1e-12 * max(n, 8)
    if cond <= 1e2:
        assert np.max(np.abs(w - wr) / wr) < 1e-12 * max(n, 8)
    assert Sr[0] > 0


@pytest.mark.parametrize('n', [8, 100, 340, 500, 1024, 1100])
def test_small_eigh_dispatch(rt, n):
    rng = np.random.RandomState(n)
    G = rng.randn(n, n); G = G + G.T
    w, Q, _ = _eig(rt, G, 'rl_small_eigh')
    wr = np.linalg.eigvalsh(G)
    assert np.max(np.abs(w - wr)) <= 1e-12 * np.max(np.abs(wr)) * n
    assert np.max(np.abs(G @ Q - Q * w[None, :])) <= 1e-11 * np.max(np.abs(wr)) * n


@pytest.mark.parametrize('nx,ny,lx,rx,lxn,rxn', [(8, 8, 8, 0, 8, 0), (16, 12, 10, 6, 12, 7), (0, 9, 0, 0, 5, 0), (12, 0, 12, 0, 12, 0),
                                                 (128, 128, 0, 128, 0, 128), (120, 97, 70, 50, 80, 60), (3, 1, 2, 1, 2, 1)])
def test_rr_solve_against_lapack(rt, nx, ny, lx, rx, lxn, rxn):
    n = nx + ny
    rng = np.random.RandomState(n + lx)
    N = 5 * n + 7
    V = rng.randn(N, n)
    if nx:
        V[:, :nx], _ = np.linalg.qr(V[:, :nx])
    V[:, nx:] /= np.linalg.norm(V[:, nx:], axis=0)
    A = np.diag(np.linspace(1.0, 50.0, N))
    GB = V.T @ V
    GA = V.T @ A @ V
    ld = n
    Ub = GB.copy()
    ind, dropped, status = E.piv_chol(Ub, n, n, 0.0)          # unpivoted factor
    U = np.triu(Ub)
    eng = E.NumpyEngine()
    class T_:                                                 # minimal template for begin()
        def dimension(self): return 4
        def data_type(self): return np.float64
    eng.m = max(n, 1)
    M = n
    z = lambda r, c: E.Small(np.zeros((r, c)))
    eng.GB, eng.GA = E.Small(U.copy()), E.Small(GA.copy())
    eng.CX, eng.CZ = z(M, M), z(M, M)
    eng.lmdx, eng.lmdz = np.zeros(M), np.zeros(M)
    eng.rayleigh_ritz(nx, ny, lx, rx, lxn, rxn)
    dX_r, dl_r = eng._est
    dGA, dU = rt.up(GA), rt.up(U)
    cx, cz = rt.zeros(n, n), rt.zeros(n, n)
    lmdx, lmdz, est = rt.zeros(n), rt.zeros(n), rt.zeros(2 * n)
    wsb = rt.lib.rl_rr_solve_ws_bytes(n)
    ws = rt.zeros(wsb // 8 + 8)
    info = rt.zeros(4, dtype=rt.torch.int32)
    rt.check(rt.lib.rl_rr_solve(dGA.data_ptr(), dU.data_ptr(), ld, nx, ny, lx, rx, lxn, rxn, cx.data_ptr(), n,
                                cz.data_ptr(), n, lmdx.data_ptr(), lmdz.data_ptr(), est.data_ptr(), n, 0.0, ws.data_ptr(),
                                wsb, info.data_ptr(), rt.st()))
    nxn, nz = lxn + rxn, n - lxn - rxn
    lx_d, lz_d = lmdx.cpu().numpy()[:nxn], lmdz.cpu().numpy()[:nz]
    assert nxn == 0 or np.max(np.abs(lx_d - eng.lmdx[:nxn])) <= 1e-11 * 50
    assert nz == 0 or np.max(np.abs(lz_d - eng.lmdz[:nz])) <= 1e-11 * 50
    C = np.concatenate((cx.cpu().numpy()[:, :nxn], cz.cpu().numpy()[:, :nz]), axis=1)
    lam = np.concatenate((lx_d, lz_d))
    assert np.max(np.abs(GA @ C - GB @ C * lam[None, :])) <= 1e-10 * 50
    assert np.max(np.abs(C.T @ GB @ C - np.eye(n))) <= 1e-11 * n
    e = est.cpu().numpy()
    if nx:
        assert np.max(np.abs(e[:nx] - dX_r)) <= 1e-8
        assert np.max(np.abs(e[n:n + nx] - dl_r)) <= 1e-8 * 50


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('m,k,n', [(1, 1, 5), (7, 3, 1000), (16, 16, 4099), (32, 20, 20011), (128, 128, 12000), (40, 130, 5003)])
def test_block_ops_with_device_small_matrices(gpu_backend, rt, dtype, m, k, n):
    from raleigh_b200.engine import DeviceEngine, DSmall
    rng = np.random.RandomState(m + k)
    s, o = rng.randn(m, n).astype(dtype), rng.randn(k, n).astype(dtype)
    S, O = gpu_backend.Vectors(s.copy()), gpu_backend.Vectors(o.copy())
    eng = DeviceEngine()
    eng.begin(S, max(m, k))
    tol = 1e-12 if dtype is np.float64 else 2e-6
    out = rt.zeros(k + 1, m + 3)
    view = DSmall(out.data_ptr(), m + 3, k + 1, m + 3).sub(1, 2, k, m)
    eng.gram(S, O, view)
    ref = o.astype(np.float64) @ s.astype(np.float64).T
    got = out.cpu().numpy()
    assert np.max(np.abs(got[1:, 2:2 + m] - ref)) <= tol * np.sqrt(n) * max(1.0, np.max(np.abs(ref)))
    assert np.all(got[0, :] == 0) and np.all(got[:, :2] == 0) and np.all(got[:, 2 + m:] == 0)
    if m == k:
        d = rt.zeros(1, m + 1)
        eng.dots(S, O, DSmall(d.data_ptr(), m + 1, 1, m + 1))
        refd = np.sum(s.astype(np.float64) * o.astype(np.float64), axis=1)
        assert np.max(np.abs(d.cpu().numpy()[0, :m] - refd)) <= (tol if dtype is np.float64 else 3e-5) * np.sqrt(n) * max(1.0, np.max(np.abs(refd)))
    q = rng.randn(k, m)
    dq = rt.up(np.pad(q, ((0, 0), (0, 5))))
    Out = gpu_backend.Vectors(rng.randn(m, n).astype(dtype))
    out0 = Out.data()
    eng.update(Out, O, DSmall(dq.data_ptr(), m + 5, k, m), -0.5, 1.0)
    ref = out0 - 0.5 * (q.astype(dtype).T @ o)
    assert np.max(np.abs(Out.data() - ref)) <= (1e-11 if dtype is np.float64 else 3e-4) * max(1.0, np.max(np.abs(ref)))
    eng.update(Out, O, DSmall(dq.data_ptr(), m + 5, k, m), 2.0, 0.0)
    ref = 2.0 * (q.astype(dtype).T @ o)
    assert np.max(np.abs(Out.data() - ref)) <= (1e-11 if dtype is np.float64 else 3e-4) * max(1.0, np.max(np.abs(ref)))
    # residual and normalisation
    lmd = rng.randn(m)
    dl = rt.up(lmd.reshape(1, -1))
    W = gpu_backend.Vectors(n, m, dtype)
    AX = gpu_backend.Vectors(rng.randn(m, n).astype(dtype))
    ax = AX.data()
    eng.residual(W, AX, S, DSmall(dl.data_ptr(), m, 1, m))
    ref = ax - lmd.astype(dtype)[:, None] * s
    assert np.max(np.abs(W.data() - ref)) <= (1e-13 if dtype is np.float64 else 1e-5) * max(1.0, np.max(np.abs(ref)))
    before = W.data()
    s2 = np.sum(before.astype(np.float64) ** 2, axis=1)
    s2[0] = 0.0
    d2 = rt.up(s2.reshape(1, -1))
    eng.scale_rsqrt(W, DSmall(d2.data_ptr(), m, 1, m))
    got = W.data()
    assert np.array_equal(got[0], before[0])          # zero norm: vector left alone (dense_numpy.py:50-52)
    if m > 1:
        assert np.max(np.abs(np.linalg.norm(got[1:].astype(np.float64), axis=1) - 1.0)) <= (1e-13 if dtype is np.float64 else 1e-6)


def test_ritz_check_and_conjugation(rt):
    rng = np.random.RandomState(5)
    nx, ld = 37, 50
    xax = rng.randn(nx, ld); xbx = np.eye(nx, ld) + 1e-3 * rng.randn(nx, ld)
    lmdx = rng.randn(nx)
    d = [rt.up(a) for a in (xax, xbx, lmdx)]
    lmd, stats = rt.zeros(nx), rt.zeros(8)
    rt.check(rt.lib.rl_rr_ritz_check(d[0].data_ptr(), d[1].data_ptr(), ld, nx, d[2].data_ptr(), lmd.data_ptr(),
                                     stats.data_ptr(), rt.st()))
    ref = np.diag(xax[:, :nx]) / np.diag(xbx[:, :nx])
    assert np.allclose(lmd.cpu().numpy(), ref, rtol=1e-15)
    st = stats.cpu().numpy()
    assert np.isclose(st[0], np.max(np.abs(ref - lmdx)) / np.max(np.abs(lmdx)), rtol=1e-14)
    assert np.isclose(st[1], np.max(np.abs(xbx[:, :nx] - np.eye(nx))), rtol=1e-14)
    nz, ny = 19, 23
    eng = E.NumpyEngine()
    z = lambda r, c: E.Small(np.zeros((r, c)))
    eng.ZAY, eng.ZBY, eng.Beta = E.Small(rng.randn(nz, ld)), E.Small(rng.randn(nz, ld)), z(nz, ld)
    eng.v_lmd, eng.v_y2, eng.v_t2 = E.Small(rng.randn(1, ld)), E.Small(rng.rand(1, ld)), E.Small(rng.rand(1, ld))
    eng.lmdz = rng.randn(ld)
    eng.ZAY.a[3, 4] = 1e9            # a coefficient the safeguard must zero
    eng.conjugation(nz, ny)
    dz = [rt.up(a) for a in (eng.ZAY.a, eng.ZBY.a, eng.v_lmd.a, eng.lmdz, eng.v_y2.a, eng.v_t2.a)]
    beta = rt.zeros(nz, ld)
    rt.check(rt.lib.rl_rr_conjugation(dz[0].data_ptr(), dz[1].data_ptr(), beta.data_ptr(), ld, nz, ny, dz[2].data_ptr(),
                                      dz[3].data_ptr(), dz[4].data_ptr(), dz[5].data_ptr(), rt.st()))
    got = beta.cpu().numpy()[:, :ny]
    assert got[3, 4] == 0.0
    assert np.allclose(got, eng.Beta.a[:nz, :ny], rtol=1e-13, atol=1e-300)
