"""GPU parity: every Vectors / Matrix / sparse method of raleigh_b200 (through
the C ABI) against the CPU oracle on the same seeded inputs, against the golden
vectors produced by the reference, and -- at sizes the oracle cannot hold --
through size-independent properties."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

import oracle
from oracle import algebra_np as K
from test_oracle import _run_backend_case, TAGS

pytestmark = pytest.mark.gpu

# floating-point tolerances of BASELINE.json north_star: fp64 1e-10 on eigenvalues,
# fp32 1e-5; per-kernel checks are tighter: a few ulps times the reduction length
TOL = {np.float64: 1e-12, np.float32: 3e-5}


def close(a, b, dtype, fac=10.0):
    t = TOL[np.dtype(dtype).type]
    scale = max(1.0, float(np.max(np.abs(b)))) if np.size(b) else 1.0
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if np.size(b) else 0.0
    assert err <= t * scale * fac, err


@pytest.mark.parametrize('tag', TAGS)
def test_golden_call_sequence(gpu_backend, tag):
    g = np.load(os.path.join(GOLDEN, 'algebra_%s.npz' % tag))
    _run_backend_case(gpu_backend.Vectors, gpu_backend.Matrix, g, close)


SHAPES = [(1, 1, 1), (1, 7, 1), (3, 1, 2), (5, 33, 4), (16, 1000, 16), (17, 4099, 9), (32, 20000, 32),
          (40, 5003, 33), (64, 3001, 70), (8, 262147, 8)]


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('m,n,k', SHAPES)
def test_gram_update_dots_against_oracle(gpu_backend, dtype, m, n, k):
    rng = np.random.RandomState(m * 1000 + k)
    s = rng.randn(m, n).astype(dtype)
    o = rng.randn(k, n).astype(dtype)
    S, O = gpu_backend.Vectors(s), gpu_backend.Vectors(o)
    fac = 10.0 * max(1.0, np.sqrt(n) / 10)
    close(S.dot(O), K.gram(s, o), dtype, fac)
    close(S.dot(S), K.gram(s, s), dtype, fac)
    close(S.dots(S), K.row_dots(s, s), dtype, fac)
    if m == k:
        close(S.dots(O), K.row_dots(s, o), dtype, fac)
        close(S.dots(O, transp=True), K.column_dots(s, o), dtype, fac)
    q = rng.randn(m, k).astype(dtype)
    out = gpu_backend.Vectors(n, k, dtype)
    S.multiply(q, out)
    close(out.data(), K.combine(s, q), dtype, fac)
    # add with C-, F-ordered and strided q
    p = rng.randn(k, m).astype(dtype)
    for pv in (p, np.asfortranarray(p), rng.randn(2 * k, 2 * m).astype(dtype)[::2, ::2]):
        S2 = gpu_backend.Vectors(s.copy())
        S2.add(O, -0.75, pv)
        close(S2.data(), K.add_combined(s, o, -0.75, np.asarray(pv)), dtype, fac)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_windows_and_zero_selection(gpu_backend, dtype):
    rng = np.random.RandomState(7)
    n, nv = 2050, 12
    a = rng.randn(nv, n).astype(dtype)
    b = rng.randn(nv, n).astype(dtype)
    A, B = gpu_backend.Vectors(a.copy()), gpu_backend.Vectors(b.copy())
    A.select(5, 3)
    B.select(4, 7)
    close(A.dot(B), K.gram(a[3:8], b[7:11]), dtype, 50)
    A.select(0, 2)
    assert A.dot(B).shape == (4, 0)
    assert A.dots(A).shape == (0,)
    A.scale(np.ones(0))
    A.add(B, 1.0, np.zeros((4, 0), dtype=dtype))
    A.zero()
    A.select_all()
    close(A.data(), a, dtype)
    # clone of a window, reference of a window
    A.select(4, 2)
    C = A.clone()
    assert C.nvec() == 4 and C.selected() == (0, 4)
    close(C.data(), a[2:6], dtype)
    R = A.reference()
    R.zero()
    A.select_all()
    exp = a.copy()
    exp[2:6] = 0
    close(A.data(), exp, dtype)
    # empty container + append growth (partial_hevp.py:211, solver lock events)
    E = gpu_backend.Vectors(n, data_type=dtype)
    assert E.nvec() == 0 and E.data().shape == (0, n)
    for t in range(5):
        B.select(3, t)
        E.append(B)
    B.select_all()
    exp = np.concatenate([b[t:t + 3] for t in range(5)])
    close(E.data(), exp, dtype)
    assert E.nvec() == 15


def test_error_behaviour(gpu_backend):
    V, M = gpu_backend.Vectors, gpu_backend.Matrix
    with pytest.raises(ValueError):
        V('nope')
    with pytest.raises(ValueError):
        V(10, 2, np.int32)
    a = M(np.zeros((4, 6), dtype=np.float32))
    x, y = V(6, 2, np.float32), V(4, 2, np.float32)
    a.apply(x, y)
    with pytest.raises(ValueError):
        a.apply(y, x)
    with pytest.raises(ValueError):
        a.apply(V(6, 2, np.float64), y)
    with pytest.raises(ValueError):
        a.apply(x, V(4, 3, np.float32))
    with pytest.raises(ValueError):
        M(np.zeros((8, 8))[::2, ::2])
    with pytest.raises(ValueError):
        V(M(np.asfortranarray(np.zeros((4, 6)))), shallow=True)
    with pytest.raises(ValueError):
        x.fill(np.zeros((3, 6), dtype=np.float32))
    with pytest.raises(AssertionError):
        x.select(3)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('M,N,k', [(5, 7, 3), (64, 64, 64), (130, 257, 17), (1000, 333, 128), (257, 1024, 128),
                                   (300, 517, 1), (1003, 2050, 2), (129, 4099, 7), (4000, 1024, 8), (67, 70, 1)])
def test_dense_apply(gpu_backend, dtype, M, N, k):
    rng = np.random.RandomState(M + N)
    a = rng.randn(M, N).astype(dtype)
    x = rng.randn(k, N).astype(dtype)
    fac = 20 * max(1.0, np.sqrt(N) / 4)
    for arr in (a, np.asfortranarray(a)):
        A = gpu_backend.Matrix(arr)
        assert A.order() == ('C_CONTIGUOUS' if arr.flags['C_CONTIGUOUS'] else 'F_CONTIGUOUS')
        X = gpu_backend.Vectors(x.copy())
        Y = gpu_backend.Vectors(M, k, dtype)
        A.apply(X, Y)
        y = K.dense_apply(a, x)
        close(Y.data(), y, dtype, fac)
        Z = gpu_backend.Vectors(N, k, dtype)
        A.apply(Y, Z, transp=True)
        close(Z.data(), K.dense_apply(a, y, transp=True), dtype, fac * np.sqrt(M))
    A = gpu_backend.Matrix(a)
    close(A.dots(), K.row_sqnorms(a), dtype, fac)
    # Vectors(Matrix, shallow=True) aliases the matrix memory (dense_cublas.py:369-376)
    V = gpu_backend.Vectors(A, shallow=True)
    V.select(1, 0)
    V.zero()
    exp = a.copy()
    exp[0] = 0
    X = gpu_backend.Vectors(x.copy())
    Y = gpu_backend.Vectors(M, k, dtype)
    A.apply(X, Y)
    close(Y.data(), K.dense_apply(exp, x), dtype, fac)


@pytest.mark.parametrize('M,N,k', [(5, 7, 9), (64, 64, 16), (130, 258, 17), (1000, 334, 128), (257, 1024, 130), (48, 20000, 33),
                                   (3001, 66, 200)])
def test_dense_apply_fp64_tensor_pipe(gpu_backend, M, N, k):
    """fp64 Matrix.apply with more than 8 vectors runs on the FP64 tensor pipe (csrc/gemm_dmma.cu: cp.async ring,
    m8n8k4 DMMA); both orientations, ragged edges in all three dimensions, alpha/beta, against NumPy and against the
    FMA-pipe kernel (knob 16 = -1)."""
    from raleigh_b200._lib import lib
    rng = np.random.RandomState(M + N + k)
    a = rng.randn(M, N)
    x = rng.randn(k, N)
    A = gpu_backend.Matrix(a.copy())
    X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(M, k, np.float64)
    A.apply(X, Y)
    y = x @ a.T
    assert np.max(np.abs(Y.data() - y)) <= 1e-13 * np.max(np.abs(y)) * np.sqrt(N)
    Z = gpu_backend.Vectors(N, k, np.float64)
    A.apply(Y, Z, transp=True)
    z = y @ a
    assert np.max(np.abs(Z.data() - z)) <= 1e-13 * np.max(np.abs(z)) * np.sqrt(M)
    lib.rl_debug_set_knob(16, -1)
    try:
        Y2, Z2 = gpu_backend.Vectors(M, k, np.float64), gpu_backend.Vectors(N, k, np.float64)
        A.apply(X, Y2)
        A.apply(Y, Z2, transp=True)
    finally:
        lib.rl_debug_set_knob(16, 0)
    assert np.max(np.abs(Y2.data() - Y.data())) <= 1e-13 * np.max(np.abs(y)) * np.sqrt(N)
    assert np.max(np.abs(Z2.data() - Z.data())) <= 1e-13 * np.max(np.abs(z)) * np.sqrt(M)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_data_matrix_as_vectors_paths(gpu_backend, dtype):
    """lra.update shapes: thousands of short vectors aliasing a data chunk
    (SURVEY section 3.4): dots, multiply(e1), add(vmean, -1, e1.T), orthogonalize."""
    rng = np.random.RandomState(3)
    n1, n, r = 3000, 200, 24
    a = rng.randn(n1, n).astype(dtype)
    A = gpu_backend.Matrix(a.copy())
    v = gpu_backend.Vectors(A, shallow=True)
    close(v.dots(v), K.row_dots(a, a), dtype, 50)
    e1 = np.ones((n1, 1), dtype=dtype)
    mean1 = v.new_vectors(1, n)
    v.multiply(e1, mean1)
    close(mean1.data(), K.combine(a, e1), dtype, 300)
    mean = (a.sum(axis=0) / n1).reshape(1, n).astype(dtype)
    vmean = v.new_vectors(mean)
    v.add(vmean, -1.0, e1.T)
    a = K.add_combined(a, mean, -1.0, e1.T)
    close(v.data(), a, dtype, 50)
    qmat, _ = np.linalg.qr(rng.randn(n, r).astype(dtype))
    comps = np.ascontiguousarray(qmat.T)
    left = v.orthogonalize(gpu_backend.Vectors(comps.copy()))
    a_new, q = K.project_out(a, comps)
    close(left.data(), q, dtype, 100)
    close(v.data(), a_new, dtype, 100)
    close(v.dots(v, transp=True), K.column_dots(a_new, a_new), dtype, 300)


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_svd_identities(gpu_backend, dtype):
    """tests_algebra.py:330-341 (reconstruction) + orthonormality + sigma vs LAPACK,
    incl. an ill-conditioned block like the one _lra_ortho feeds (lra.py:473-482)."""
    rng = np.random.RandomState(11)
    for m, n, cond in [(1, 50, 1.0), (8, 300, 1.0), (16, 1000, 1e3), (40, 777, 1e2)]:
        a = rng.randn(m, n).astype(dtype)
        if cond > 1:
            a *= np.logspace(0, -np.log10(cond), m).astype(dtype)[:, None]
            a = (rng.randn(m, m).astype(dtype) @ a)
        W = gpu_backend.Vectors(a.copy())
        sigma, q = W.svd()
        w = W.data()
        t = TOL[dtype]
        ref_sigma = np.linalg.svd(a.astype(np.float64), compute_uv=False)
        assert np.max(np.abs(sigma - ref_sigma)) <= 100 * t * ref_sigma[0]
        assert np.all(np.diff(sigma) <= 1e-6 * sigma[0])
        assert np.max(np.abs(w.astype(np.float64) @ w.T.astype(np.float64) - np.eye(m))) <= 200 * t
        recon = (q.astype(np.float64) * sigma[None, :].astype(np.float64)) @ w.astype(np.float64)
        assert np.linalg.norm(recon - a) / np.linalg.norm(a) <= 200 * t


def test_lra_ortho_identity(gpu_backend):
    """tests_algebra.py:43-82 (test_lra_ortho): v u^H is preserved, u orthonormalised."""
    rng = np.random.RandomState(2)
    k, nu, nv_ = 8, 500, 300
    u0 = rng.randn(k, nu)
    v0 = rng.randn(k, nv_)
    Vc = gpu_backend.Vectors
    u, v, wu, wv = Vc(u0.copy()), Vc(v0.copy()), Vc(nu, k), Vc(nv_, k)
    u.copy(wu)
    s, q = wu.svd()
    v.multiply(q, wv)
    wv.scale(s, multiply=True)
    wv.copy(v)
    s, q = v.svd()
    wu.multiply(q, u)
    v.scale(s, multiply=True)
    prod0 = v0.T @ u0
    prod1 = v.data().T @ u.data()
    assert np.linalg.norm(prod1 - prod0) / np.linalg.norm(prod0) < 1e-12
    g = u.data() @ u.data().T
    assert np.max(np.abs(g - np.eye(k))) < 1e-12


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_spmm_and_jacobi(gpu_backend, dtype):
    import scipy.sparse as sp
    from tests_common import spd_c3_like
    rng = np.random.RandomState(4)
    cases = [K.lap3d_csr(7, 6, 5).astype(dtype), spd_c3_like(1111).astype(dtype)]
    R = sp.random(300, 300, density=0.03, random_state=9, format='csr')
    cases.append(((R + R.T) + sp.diags(np.arange(1.0, 301.0))).tocsr().astype(dtype))
    # ragged: a few long rows (beyond the shared-memory staging capacity) + empty rows
    D = sp.lil_matrix((700, 700))
    D[0, :] = 1.0
    D[:, 0] = 1.0
    D[5, 5] = 2.0
    cases.append(D.tocsr().astype(dtype))
    for A in cases:
        n = A.shape[0]
        op = gpu_backend.SparseSymmetricMatrix(A)
        assert op.size() == n and op.data_type() == np.dtype(dtype)
        oop = oracle.SparseSymmetricMatrix(A)
        for m in (1, 3, 8, 17):
            x = rng.randn(m, n).astype(dtype)
            X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(n, m, dtype)
            op.apply(X, Y)
            ref = K.sym_spmm(oop.csr(), x)
            close(Y.data(), ref, dtype, 100 * max(1.0, float(abs(A).max())))
    # upper-triangle semantics on an unsymmetric input (mkl 'SUNF')
    B = sp.random(64, 64, density=0.2, random_state=1, format='csr').astype(dtype)
    op = gpu_backend.SparseSymmetricMatrix(B)
    x = rng.randn(4, 64).astype(dtype)
    X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(64, 4, dtype)
    op.apply(X, Y)
    close(Y.data(), K.sym_spmm(K.sym_upper_csr(B), x), dtype, 50)
    # Jacobi through Operator
    A = spd_c3_like(1111).astype(dtype)
    T = gpu_backend.Operator(gpu_backend.DiagonalPreconditioner(A))
    x = rng.randn(5, 1111).astype(dtype)
    X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(1111, 5, dtype)
    T.apply(X, Y)
    close(Y.data(), K.jacobi_apply(A.diagonal(), x), dtype, 10)
    # host-only user preconditioner (reference contract, partial_hevp.py:64-73)
    T2 = gpu_backend.Operator(oracle.Jacobi(A))
    T2.apply(X, Y)
    close(Y.data(), K.jacobi_apply(A.diagonal(), x), dtype, 10)


def test_fill_random_host_stream_and_device_partition_independence(gpu_backend):
    n = 1003
    np.random.seed(1)
    V = gpu_backend.Vectors(n, 4)
    V.fill_random()
    np.random.seed(1)
    ref = K.uniform_fill_cublas(4, n)
    assert np.array_equal(V.data(), ref)
    # device fill: a shard (rows 200..700) reproduces the same rows of the full fill
    full = gpu_backend.Vectors(n, 6)
    full.fill_random_device(1234)
    part = gpu_backend.Vectors(501, 6)
    part.fill_random_device(1234, row0=200)
    f = full.data()
    assert np.array_equal(part.data(), f[:, 200:701])
    assert abs(f.mean()) < 0.05 and abs(f.var() - 1 / 3) < 0.05 and f.min() >= -1 and f.max() < 1
    f32 = gpu_backend.Vectors(n, 3, np.float32)
    f32.fill_random_device(7)
    p32 = gpu_backend.Vectors(n - 5, 3, np.float32)
    p32.fill_random_device(7, row0=5)
    assert np.array_equal(p32.data(), f32.data()[:, 5:])


def test_large_properties(gpu_backend):
    """Sizes the oracle cannot hold comfortably: linearity, symmetry, norms."""
    n, m = 4_000_000, 32
    X = gpu_backend.Vectors(n, m)
    X.fill_random_device(5)
    G = X.dot(X)
    assert np.allclose(G, G.T, rtol=0, atol=1e-9 * n)
    d = X.dots(X)
    assert np.allclose(np.diag(G), d, rtol=1e-12)
    assert np.allclose(d / n, 1 / 3, rtol=5e-3)
    Y = gpu_backend.Vectors(n, m)
    q = np.eye(m)[:, ::-1].copy()
    X.multiply(q, Y)                      # permutation: exact
    assert np.array_equal(Y.dots(Y), d[::-1])
    Y.add(X, -1.0, q)                     # exact cancellation
    assert np.max(np.abs(Y.dots(Y))) == 0.0
    # determinism: same bits on repeat
    assert np.array_equal(G, X.dot(X))


GRAM_TMA_SHAPES = [(32, 8192, 32), (17, 10007, 32), (9, 65539, 5), (40, 20000, 33), (8, 12345, 8), (1, 9000, 32),
                   (16, 300001, 16)]


@pytest.mark.parametrize('mode', [1, 2, 3])
@pytest.mark.parametrize('m,n,k', GRAM_TMA_SHAPES)
def test_gram_tma_variant_against_oracle(gpu_backend, mode, m, n, k):
    """The TMA-fed Gram kernel (csrc/gram_tma.cu), forced through the A/B knob: every tile shape,
    ragged n (TMA zero-fill), offset windows, X.dot(X) with shared fragments, repeatability."""
    from raleigh_b200._lib import lib
    rng = np.random.RandomState(m * 100 + k)
    s = rng.randn(m + 3, n)
    o = rng.randn(k + 2, n)
    S, O = gpu_backend.Vectors(s.copy()), gpu_backend.Vectors(o.copy())
    S.select(m, 3)
    O.select(k, 2)
    fac = 10.0 * max(1.0, np.sqrt(n) / 10)
    lib.rl_debug_set_knob(0, mode)
    try:
        g = S.dot(O)
        close(g, K.gram(s[3:], o[2:]), np.float64, fac)
        assert np.array_equal(g, S.dot(O))
        close(S.dot(S), K.gram(s[3:], s[3:]), np.float64, fac)
    finally:
        lib.rl_debug_set_knob(0, 0)


@pytest.mark.parametrize('group', [4, 8, 16])
def test_spmm_clustered_run_order_is_bit_identical(gpu_backend, group):
    """Footprint-clustered CTA composition (rl_spmm_cluster_runs) only permutes which rows a CTA
    owns: Y must equal the consecutive-run kernel's bit for bit, and the oracle's to rounding."""
    from raleigh_b200 import sparse as rsp
    rng = np.random.RandomState(group)
    for A in (K.lap3d_csr(20, 17, 13), K.lap3d_csr(32, 32, 8)):
        n = A.shape[0]
        x = rng.randn(9, n)
        X = gpu_backend.Vectors(x.copy())
        saved = rsp.SPMM_CLUSTER_WARPS
        try:
            rsp.SPMM_CLUSTER_WARPS = 0
            op0 = gpu_backend.SparseSymmetricMatrix(A)
            rsp.SPMM_CLUSTER_WARPS = group
            op1 = gpu_backend.SparseSymmetricMatrix(A)
        finally:
            rsp.SPMM_CLUSTER_WARPS = saved
        assert op0.layout() == 'csr' and op1.layout() == 'csr+clustered%d' % group
        assert op1.footprint_ratio < 4.0
        Y0, Y1 = gpu_backend.Vectors(n, 9), gpu_backend.Vectors(n, 9)
        op0.apply(X, Y0)
        op1.apply(X, Y1)
        assert np.array_equal(Y0.data(), Y1.data())
        close(Y1.data(), K.sym_spmm(K.sym_upper_csr(A), x), np.float64, 100 * float(abs(A).max()))


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_matrix_minmax_and_amatrix_scale(gpu_backend, dtype, ref_root):
    """rl_minmax_h against numpy.amin/amax (C and F order, padded leading dimension, negative
    extremes), and the reference's own AMatrix.scale() (dense_matrix.py:32-34, :63-64) served from
    the device copy through compat's numpy proxy."""
    from raleigh_b200 import vectors as rv
    rng = np.random.RandomState(3)
    one = gpu_backend.Matrix(np.full((1, 1), -7.0, dtype=dtype)).minmax()
    assert one[0] == one[1] == -7.0
    for shape in ((1, 2), (3, 1001), (257, 130), (64, 70000)):
        a = (rng.randn(*shape) * 3).astype(dtype)
        a[-1, -1] = -50.0
        a[0, 0] = 40.0
        for arr in (a, np.asfortranarray(a)):
            lo, hi = gpu_backend.Matrix(arr).minmax()
            assert lo == np.amin(a) == -50.0 and hi == np.amax(a) == 40.0
    gpu_backend.install(ref_root)
    from raleigh.algebra.dense_matrix import AMatrix
    saved = rv.MINMAX_ON_DEVICE_BYTES
    rv.MINMAX_ON_DEVICE_BYTES = 0
    try:
        a = rng.randn(300, 200).astype(dtype)
        am = AMatrix(a, arch='gpu!')
        assert am.scale() == max(abs(a.min()), abs(a.max()))
        assert not rv._recent_uploads                  # entry consumed by the amin/amax pair
        assert np.amin(a) == a.min()
    finally:
        rv.MINMAX_ON_DEVICE_BYTES = saved


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_spmm_long_rows_csr_batches_and_sell_layout(gpu_backend, dtype):
    """>= 16 entries per row: the staged-CSR kernel switches to 4-entry batches (127-register
    budget); the opt-in SELL-32 layout must give the same product.  Odd vector counts exercise the
    ragged last group of both."""
    from raleigh_b200 import sparse as rsp
    from tests_common import spd_c3_like
    offs = (1, 2, 3, 4, 5, 6, 40, 41, 42, 43, 44, 45, 46, 300, 301, 302, 303)
    A = spd_c3_like(3000, offsets=offs).astype(dtype)
    assert A.nnz >= 24 * A.shape[0]
    rng = np.random.RandomState(8)
    saved = rsp.SPMM_LAYOUT
    try:
        ops = []
        for layout in ('csr', 'sell'):
            rsp.SPMM_LAYOUT = layout
            ops.append(gpu_backend.SparseSymmetricMatrix(A))
    finally:
        rsp.SPMM_LAYOUT = saved
    assert ops[0].layout() == 'csr' and ops[1].layout() == 'sell32'
    for m in (1, 8, 13, 32, 35):
        x = rng.randn(m, 3000).astype(dtype)
        ref = K.sym_spmm(K.sym_upper_csr(A), x)
        for op in ops:
            X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(3000, m, dtype)
            op.apply(X, Y)
            close(Y.data(), ref, dtype, 100 * float(abs(A).max()))


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
def test_spmm_halo_kernels_on_one_gpu(gpu_backend, dtype):
    """The row-sharded (halo) variants of the CSR and SELL-32 kernels, driven directly through the
    C ABI on ONE device: rows [0, nloc) of the operator, owned columns read from the local block,
    the other columns from a row-interleaved halo buffer built on the host the way the NCCL
    exchange delivers it (halo[(c - nloc) * m + v])."""
    from raleigh_b200 import sparse as rsp
    from raleigh_b200 import device as dev
    from raleigh_b200._lib import lib, check, dtype_code
    from tests_common import spd_c3_like
    offs = (1, 2, 3, 4, 5, 6, 40, 41, 42, 43, 44, 45, 46, 300, 301, 302, 303)
    A = spd_c3_like(2000, offsets=offs).astype(dtype).tocsr()
    n, nloc = 2000, 1100
    slab = A[:nloc].tocsr()
    slab.sort_indices()
    halo_cols = np.unique(slab.indices[slab.indices >= nloc])
    remap = np.arange(n, dtype=np.int64)
    remap[halo_cols] = nloc + np.arange(halo_cols.size)
    indptr = np.ascontiguousarray(slab.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(remap[slab.indices], dtype=np.int32)
    values = np.ascontiguousarray(slab.data, dtype=dtype)
    d_ip, d_ix, d_va = rsp._to_device(indptr), rsp._to_device(indices), rsp._to_device(values)
    sell = rsp._build_sell32(indptr, indices, values)
    assert sell is not None
    code = dtype_code(dtype)
    rng = np.random.RandomState(12)
    for m in (1, 8, 13, 35):
        x = rng.randn(m, n).astype(dtype)
        ref = (slab @ x.T).T
        X = gpu_backend.Vectors(np.ascontiguousarray(x[:, :nloc]))
        H = rsp._to_device(np.ascontiguousarray(x[:, halo_cols].T))         # (nhalo, m): row-interleaved
        scale = 100 * float(abs(A).max())
        Y = gpu_backend.Vectors(nloc, m, dtype)
        check(lib.rl_csr_spmm_halo(code, nloc, slab.nnz, d_ip.ptr, d_ix.ptr, d_va.ptr, X._wptr(), X._ld, Y._wptr(),
                                   Y._ld, m, nloc, H.ptr, dev.stream()))
        close(Y.data(), ref, dtype, scale)
        Y2 = gpu_backend.Vectors(nloc, m, dtype)
        sp_, sc_, sv_, nsl = sell
        check(lib.rl_sell_spmm_halo(code, nloc, slab.nnz, nsl, sp_.ptr, sc_.ptr, sv_.ptr, X._wptr(), X._ld,
                                    Y2._wptr(), Y2._ld, m, nloc, H.ptr, dev.stream()))
        close(Y2.data(), ref, dtype, scale)


def test_dense_apply_skinny_at_config2_size(gpu_backend):
    """k in {1, 2, 7} vectors against the 12,000 x 39,375 fp32 data matrix (the mean-shift vectors of the PCA
    operator, partial_svd.py:256-277): the HBM-bound sweep of gemm_skinny.cu against fp64 NumPy on a row /
    column sample, and linearity over the whole result."""
    import torch
    M, N = 12000, 39375
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    a = torch.randn(M, N, generator=g, device='cuda', dtype=torch.float32)
    A = gpu_backend.Matrix(a.cpu().numpy())
    rng = np.random.RandomState(1)
    for k in (1, 2, 7):
        x = rng.randn(k, N).astype(np.float32)
        X, Y = gpu_backend.Vectors(x.copy()), gpu_backend.Vectors(M, k, np.float32)
        A.apply(X, Y)
        ref = (torch.from_numpy(x).cuda().double() @ a.double().T).cpu().numpy()
        got = Y.data()
        assert np.max(np.abs(got - ref)) <= 3e-5 * np.sqrt(N) * max(1.0, np.max(np.abs(ref)) / np.sqrt(N))
        z = rng.randn(k, M).astype(np.float32)
        Z, W = gpu_backend.Vectors(z.copy()), gpu_backend.Vectors(N, k, np.float32)
        A.apply(Z, W, transp=True)
        ref = (torch.from_numpy(z).cuda().double() @ a.double()).cpu().numpy()
        got = W.data()
        assert np.max(np.abs(got - ref)) <= 3e-5 * np.sqrt(M) * max(1.0, np.max(np.abs(ref)) / np.sqrt(M))


@pytest.mark.parametrize('dtype', [np.float64, np.float32])
@pytest.mark.parametrize('case', ['cond1e10', 'rank_deficient', 'zero_rows', 'graded'])
def test_svd_returns_a_full_orthonormal_set(gpu_backend, dtype, case):
    """Vectors.svd() on ill-conditioned and rank-deficient blocks (the situations in which the reference calls
    it: partial_svd.py:183 `icond < 100 eps`, the restart at solver.py:885): S_new must have ORTHONORMAL rows --
    numpy.linalg.svd always returns them -- and S_old = v diag(sigma) S_new must hold."""
    rng = np.random.RandomState(11)
    m, n = 24, 3000
    u, _ = np.linalg.qr(rng.randn(m, m))
    w, _ = np.linalg.qr(rng.randn(n, m))
    big = dtype is np.float64
    if case == 'cond1e10':
        sig = np.logspace(0, -10 if big else -5, m)
    elif case == 'rank_deficient':
        sig = np.concatenate((np.linspace(1, 0.1, m - 7), np.zeros(7)))
    elif case == 'zero_rows':
        sig = np.linspace(1, 0.5, m)
    else:
        sig = np.logspace(0, -6 if big else -3, m)
    s = ((u * sig) @ w.T).astype(dtype)
    if case == 'zero_rows':
        s[3] = 0
        s[17] = 0
    S = gpu_backend.Vectors(s.copy())
    sigma, v = S.svd()
    snew = S.data().astype(np.float64)
    eps = np.finfo(dtype).eps
    assert np.max(np.abs(snew @ snew.T - np.eye(m))) < 50 * m * eps, case
    recon = (v.astype(np.float64) * sigma.astype(np.float64)[None, :]) @ snew
    assert np.max(np.abs(recon - s)) < 50 * m * eps * max(1.0, np.max(np.abs(s)))
    assert np.all(np.diff(sigma) <= 1e-6 * sigma[0])
    exact = np.linalg.svd(s.astype(np.float64), compute_uv=False)
    live = exact > (1e-9 if big else 3e-4) * exact[0]
    assert np.max(np.abs(sigma[live] - exact[live]) / exact[live]) < (1e-6 if big else 2e-3)
    assert np.max(np.abs(v.astype(np.float64).T @ v.astype(np.float64) - np.eye(m))) < 100 * m * eps
