"""CPU-side checks of the SpMM set-up code that runs on the host (C++ in the library, no device
work): rl_spmm_cluster_runs must return a permutation of the 32-row runs whatever the matrix,
and on grid stencils it must cut the per-CTA column footprint the way DESIGN.md claims."""
import ctypes

import numpy as np
import scipy.sparse as sp

from oracle import algebra_np as K
from tests_common import spd_c3_like


def _cluster(A, group):
    from raleigh_b200._lib import lib
    A = A.tocsr()
    A.sort_indices()
    n = A.shape[0]
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    order = np.full((n + 31) // 32, -1, dtype=np.int32)
    ratio = ctypes.c_double()
    rc = lib.rl_spmm_cluster_runs(n, indptr.ctypes.data, indices.ctypes.data, group, order.ctypes.data,
                                  ctypes.byref(ratio))
    assert rc == 0
    return order, ratio.value


def _footprint(A, order, group):
    """Distinct 32-column segments gathered per CTA (all entries counted), per run."""
    A = A.tocsr()
    n = A.shape[0]
    nruns = (n + 31) // 32
    total = 0
    for c0 in range(0, nruns, group):
        segs = set()
        for a in order[c0:c0 + group]:
            r0, r1 = a * 32, min(a * 32 + 32, n)
            segs.update((A.indices[A.indptr[r0]:A.indptr[r1]] >> 5).tolist())
        total += len(segs)
    return total / nruns


def test_cluster_is_a_permutation_on_ragged_and_empty_rows():
    rng = np.random.default_rng(3)
    for n in (1, 31, 32, 33, 1000, 4097):
        A = sp.random(n, n, density=min(1.0, 6.0 / n), random_state=rng, format='csr')
        A = (A + A.T).tocsr()
        A[n // 2, :] = 0                      # an empty row
        A.eliminate_zeros()
        for group in (1, 4, 8, 16):
            order, _ = _cluster(A, group)
            assert np.array_equal(np.sort(order), np.arange((n + 31) // 32))


def test_cluster_cuts_the_stencil_footprint():
    L = K.lap3d_csr(32, 32, 32)
    ident = np.arange(L.shape[0] // 32, dtype=np.int32)
    base = _footprint(L, ident, 4)
    for group, bound in ((4, 3.2), (8, 2.6), (16, 2.2)):
        order, ratio = _cluster(L, group)
        fp = _footprint(L, order, group)
        assert fp < bound < base, (group, fp, base)
        assert ratio <= fp + 1e-12           # the library counts only segments with >= 8 entries


def test_clustered_spmm_is_the_same_product():
    """Emulate the kernel's slot -> run mapping on the host: any permutation of the runs gives
    the same Y (rows are independent), so clustering can never change results."""
    A = spd_c3_like(1500)
    order, _ = _cluster(A, 4)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 1500))
    y = np.zeros_like(x)
    for slot, a in enumerate(order):
        r0, r1 = a * 32, min(a * 32 + 32, 1500)
        y[:, r0:r1] = (A[r0:r1] @ x.T).T
    assert np.allclose(y, (A @ x.T).T, rtol=0, atol=0)
