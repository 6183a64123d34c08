"""Synthetic inputs shared by the CPU and GPU test suites and bench.py."""
import numpy as np
import scipy.sparse as sp


def spd_c3_like(n, seed=0, offsets=(1, 2, 3, 7, 19, 20, 21)):
    """Small twin of BASELINE config 3: banded symmetric, strictly diagonally
    dominant SPD matrix with a strongly varying diagonal (Jacobi is a real
    preconditioner).  Must stay identical to tests/golden/make_golden.py."""
    rng = np.random.default_rng(seed)
    diags = [-rng.uniform(0.1, 1.0, n - o) for o in offsets]
    L = sp.diags(diags, list(offsets), shape=(n, n), format='csr')
    S = L + L.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(0.01, 1.0, n) * np.linspace(1, 100, n)
    return (S + sp.diags(d)).tocsr()
