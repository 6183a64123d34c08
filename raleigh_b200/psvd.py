"""Device-resident post-processing of the partial SVD.

Replaces `PartialSVD._finalize_svd` (raleigh/interfaces/partial_svd.py:163-235),
which takes the nsv computed right singular vectors v and their images Av and
returns (u, sigma, v') with orthonormal u and A v' = u diag(sigma).  The
reference does it on the host with nsv x nsv LAPACK calls: `eigh(Gram, Diag)` for
a conditioning test, `cholesky`, `svd`, `inv` (0.17 s at nsv = 1000, a quarter
of a config-2 solve once the algebra runs on the GPU).

The same mathematics with one symmetric eigen-decomposition on the device:
with Gram = Av^T Av = U^T U and svd(U) = p Sigma q^T,

    Gram = q Sigma^2 q^T          and          Av inv(U) p = Av q Sigma^-1,

so u = Av q Sigma^-1 and v' = v q come from eigh(Gram) alone.  The conditioning
test needs lambda_min / lambda_max of the diagonally scaled Gram matrix S: a
Gershgorin bound computed on the device settles it whenever it already
exceeds the threshold (the usual case: S = I + O(svtol)); otherwise the
eigenvalues of S are computed on the device as well.  The optional re-
orthogonalisation sweeps (:207-231) are the same step applied to u Sigma.
"""
import numpy

from ._lib import lib, check
from . import device as dev
from .vectors import Vectors, block_gemm


class _Work:
    """fp64 device workspace for one nsv x nsv factorisation."""

    def __init__(self, n, dtype):
        self.n = n
        # Jacobi stopping tolerance: working precision for fp64 data; fp32 Gram matrices carry 1e-7 noise
        self.tol = 1e-9 if numpy.dtype(dtype) == numpy.float32 else 0.0
        ews = lib.rl_small_eigh_ws_bytes(n)
        mat = (n * n * 8 + 255) & ~255                 # every piece 256-byte aligned (128-bit accesses in the kernels)
        vec = (n * 8 + 255) & ~255
        self.buf = dev.Buffer(5 * mat + 4 * vec + ews + 4096, zero=True)
        p = (self.buf.ptr + 255) & ~255
        self.G, self.S, self.Q, self.q, self.cs = (p + i * mat for i in range(5))
        p += 5 * mat
        self.w, self.sigma, self.ger = p, p + vec, p + 2 * vec
        self.info = p + 3 * vec
        self.ews, self.ews_bytes = p + 4 * vec, ews


def _gram64(block, work, exact=False):
    """work.G = block block^T in fp64.  fp32 blocks: tensor-core product (3xTF32, fp32 accumulation: what the
    reference's sgemm delivers), widened -- unless `exact`, which accumulates the exact fp32 products in fp64
    (Vectors.svd needs singular values far below sqrt(eps_32) sigma_max)."""
    n = work.n
    st = dev.stream()
    if block._code == 0 and n >= 48 and not exact:
        t = Vectors._local(n, n, block.data_type())
        block_gemm(block._code, block._wptr(), block._ld, n, block._n, block._wptr(), block._ld, t._wptr(), t._ld, n, 0)
        if block._shard is not None:
            block._reduce_device(t._buf, n * t._ld, block.data_type())
        check(lib.rl_block_to_small(block._code, t._wptr(), t._ld, n, n, work.G, n, st))
        return
    check(lib.rl_gram_dev(block._code, block._wptr(), block._ld, n, block._wptr(), block._ld, n, block._n, work.G, n, st))
    if block._shard is not None:
        import torch
        off = work.G - work.buf.ptr
        block._shard[0].allreduce_(work.buf.tensor[off:off + n * n * 8].view(torch.float64))


def _eigh(work, src):
    check(lib.rl_small_eigh(src, work.n, work.n, work.tol, work.w, work.Q, work.n, work.ews, work.ews_bytes,
                            work.info, dev.stream()))


def _factor(work):
    """work.S <- upper Cholesky factor of work.G; False when G is not numerically positive definite."""
    n, st = work.n, dev.stream()
    if n > lib.rl_syevj_grid_max_n():
        return False
    check(lib.rl_small_copy(work.G, n, work.S, n, n, n, st))
    check(lib.rl_small_potrf(work.S, n, n, work.info, st))
    flag = numpy.zeros(1, dtype=numpy.int32)
    check(lib.rl_d2h(dev.host_ptr(flag), work.info, 4, st))
    return flag[0] == 0


def _eigh_gram(work, factored=None):
    """Eigen-decomposition of the Gram matrix work.G -> work.w (ascending), work.Q.  Through its Cholesky
    factor when that exists (Jacobi on the factor keeps the relative accuracy of the small singular values,
    like the reference's cholesky + svd, partial_svd.py:192-193); plain symmetric Jacobi otherwise."""
    n, st = work.n, dev.stream()
    if factored is None:
        factored = _factor(work)
    if factored:
        check(lib.rl_small_eigh_factor(work.S, n, n, work.tol, work.w, work.Q, n, work.ews, work.ews_bytes, work.info,
                                       st))
    else:
        _eigh(work, work.G)


def _fetch(ptr, count):
    h = numpy.empty(count, dtype=numpy.float64)
    check(lib.rl_d2h(dev.host_ptr(h), ptr, count * 8, dev.stream()))
    return h


def _rotate(block, coeff_ptr, n, tmp):
    """block <- coeff^T-combination of its vectors: new_j = sum_i coeff[i, j] block_i."""
    ct = Vectors._local(n, n, block.data_type())
    check(lib.rl_small_to_block(block._code, coeff_ptr, n, n, n, 1, ct._wptr(), ct._ld, dev.stream()))
    tmp.select(n)
    tmp._touch()
    block_gemm(block._code, block._wptr(), block._ld, n, block._n, ct._wptr(), ct._ld, tmp._wptr(), tmp._ld, n, 1)
    tmp.copy(block)


def _orthonormalise(block, work, tmp):
    """block <- block q Sigma^-1 with Gram(block) = q Sigma^2 q^T; leaves q in work.q, sigma in work.sigma."""
    _gram64(block, work)
    _eigh_gram(work)
    check(lib.rl_psvd_coeffs(work.Q, work.n, work.w, work.n, work.q, work.cs, work.n, work.sigma, dev.stream()))
    _rotate(block, work.cs, work.n, tmp)


def finalize_svd(v, Av, eps):
    """Drop-in for PartialSVD._finalize_svd(v, Av, eps) on raleigh_b200 vectors."""
    nsv = v.nvec()
    dtype = v.data_type()
    work = _Work(nsv, dtype)
    st = dev.stream
    _gram64(Av, work)

    # conditioning of the Gram matrix (partial_svd.py:171-182): icond = lambda_min / lambda_max of the
    # diagonally scaled S = D^-1/2 G D^-1/2, needed only to be compared with delta.  Certified lower bounds
    # first -- Gershgorin discs, then 1 / (||D^1/2 U^-1||_F^2 ||S||_inf) from the Cholesky factor G = U^T U that
    # the fast route needs anyway -- and the eigenvalues of S on the device only if neither settles it.
    check(lib.rl_psvd_gershgorin(work.G, nsv, nsv, work.ger, st()))
    radius, dmin, dmax = _fetch(work.ger, 3)
    delta = 100 * float(numpy.finfo(dtype).eps)
    factored = None
    if not dmin > 0.0:
        icond = 0.0
    elif radius < 1.0 and (1.0 - radius) / (1.0 + radius) >= delta:
        icond = (1.0 - radius) / (1.0 + radius)
    else:
        icond = -1.0
        factored = _factor(work)
        if factored:
            check(lib.rl_small_set_identity(work.Q, nsv, nsv, st()))
            check(lib.rl_small_trsm(1, work.S, nsv, nsv, work.Q, nsv, nsv, st()))          # U^-1
            check(lib.rl_psvd_invbound(work.Q, nsv, work.G, nsv, nsv, work.ger, st()))
            bound = 1.0 / (float(_fetch(work.ger, 1)[0]) * (1.0 + radius))
            if bound >= delta:
                icond = bound
        if icond < 0.0:
            check(lib.rl_psvd_scale(work.G, nsv, nsv, work.q, nsv, st()))
            _eigh(work, work.q)
            lmd = _fetch(work.w, nsv)
            icond = lmd[0] / lmd[-1]
    if icond < delta:          # Av too ill-conditioned for the Gram route: SVD of Av itself (:183-189)
        sigma, q = Av.svd()
        w = v.new_vectors(nsv)
        v.multiply(q, w)
        w.copy(v)
        return Av, sigma, v

    # A v = (Av q Sigma^-1) Sigma q^T (:191-197)
    tmp = Av.new_vectors(nsv)
    _eigh_gram(work, factored)
    check(lib.rl_psvd_coeffs(work.Q, nsv, work.w, nsv, work.q, work.cs, nsv, work.sigma, st()))
    _rotate(Av, work.cs, nsv, tmp)
    u = Av
    qtot = dev.Buffer(nsv * nsv * 8)
    check(lib.rl_small_copy(work.q, nsv, qtot.ptr, nsv, nsv, nsv, st()))

    # orthonormality of the trailing vectors decides whether refinement is needed (:199-210)
    nv = int(min(32, nsv // 2))
    no_max = 0.0
    if nv > 0:
        tail = u.reference()
        tail.select(nv, nsv - nv)
        g = tail.dot(tail)
        no_max = float(numpy.amax(abs(g - numpy.eye(nv, dtype=g.dtype))))
    if no_max >= eps:
        it = 0
        while it < 2:
            # one more pass on u Sigma: Gram = Sigma (u^T u) Sigma = qh^T Sigma'^2 qh (:216-231)
            check(lib.rl_scale(u._code, u._wptr(), u._ld, nsv, u._n, _typed(work.sigma, nsv, u), 1, st()))
            _orthonormalise(u, work, tmp)
            prod = dev.Buffer(nsv * nsv * 8)
            check(lib.rl_small_gemm(0, 0, nsv, nsv, nsv, 1.0, qtot.ptr, nsv, work.q, nsv, 0.0, prod.ptr, nsv, st()))
            qtot = prod
            it += 1
            _gram64(u, work)
            gram = _fetch(work.G, nsv * nsv).reshape(nsv, nsv)
            if numpy.amax(gram - numpy.eye(nsv)) <= eps:
                break
    sigma = _fetch(work.sigma, nsv).astype(dtype)
    w = v.new_vectors(nsv)
    _rotate(v, qtot.ptr, nsv, w)
    return u, sigma, v


def _typed(ptr64, n, like):
    """Device pointer to n coefficients in the block's dtype (fp64 source)."""
    if like._code == 1:
        return ptr64
    t = Vectors._local(n, 1, like.data_type())
    check(lib.rl_small_to_block(like._code, ptr64, n, 1, n, 0, t._wptr(), t._ld, dev.stream()))
    _typed.keep = t            # lives until the next call (stream order makes that safe)
    return t._wptr()
