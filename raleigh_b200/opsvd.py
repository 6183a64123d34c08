"""Device-side `_OperatorSVD.apply` (raleigh/interfaces/partial_svd.py:258-291).

The reference forms y = (A - e a)(A - e a)^T x (or the transposed product) from two products with the data matrix
and rank-one corrections whose coefficients it pulls to the host: `s = x.dot(ones)`, `s = z.dot(aves)` -- two
blocking device-to-host copies -- and it ends with a device synchronise for its timer.  Three pipeline drains per
operator application, ~50 per config-2 solve.  Same arithmetic here with the (1 x k) coefficient rows kept in
device memory (rl_gram_dev -> rl_update_dev); `time` then accumulates the host time of the launches only.
"""
import time

from ._lib import lib, check
from . import device as dev


def _coeff(op_svd, k):
    buf = getattr(op_svd, '_rl_coeff', None)
    if buf is None or buf.nbytes < k * 8 + 256:
        buf = dev.Buffer(max(k, 256) * 8 + 256, zero=True)
        op_svd._rl_coeff = buf
    return (buf.ptr + 255) & ~255


def _dot_row(v, one, out_ptr, k):
    """out (1 x k, device fp64) = <one, v_j>: what `v.dot(one)` returns."""
    check(lib.rl_gram_dev(v._code, v._wptr(), v._ld, k, one._wptr(), one._ld, 1, v._n, out_ptr, k, dev.stream()))


def _add_rank_one(v, one, alpha, coeff_ptr, k):
    """v_j += alpha coeff[j] one: what `v.add(one, alpha, s)` does with s on the host."""
    v._touch()
    check(lib.rl_update_dev(v._code, v._wptr(), v._ld, k, one._wptr(), one._ld, 1, coeff_ptr, k, float(alpha), 1.0,
                            v._n, dev.stream()))


def supported(op_svd, x, y):
    if not (hasattr(x, '_rl_device_block') and hasattr(y, '_rl_device_block')):
        return False
    if x.is_sharded() or y.is_sharded() or getattr(op_svd.op, '_mshard', None) is not None:
        return False            # sharded products end with an all-reduce of the host result: reference route
    return x.nvec() > 0


def apply(op_svd, x, y):
    m, n = op_svd.op.shape()
    k = x.nvec()
    start = time.time()
    c = _coeff(op_svd, k)
    if op_svd.transp:
        if op_svd.w.nvec() < k:
            op_svd.w = x.new_vectors(k, n)
        z = op_svd.w
        z.select(k)
        op_svd.op.apply(x, z, transp=True)
        if op_svd.shift:
            _dot_row(x, op_svd.ones, c, k)
            _add_rank_one(z, op_svd.aves, -1.0, c, k)
        op_svd.op.apply(z, y)
        if op_svd.shift:
            _dot_row(z, op_svd.aves, c, k)
            _add_rank_one(y, op_svd.ones, -1.0, c, k)
    else:
        if op_svd.w.nvec() < k:
            op_svd.w = x.new_vectors(k, m)
        z = op_svd.w
        z.select(k)
        op_svd.op.apply(x, z)
        if op_svd.shift:
            for _ in range(2):          # "accurate orthogonalization needed!" (partial_svd.py:284)
                _dot_row(z, op_svd.ones, c, k)
                _add_rank_one(z, op_svd.ones, -1.0 / m, c, k)
        op_svd.op.apply(z, y, transp=True)
    op_svd.time += time.time() - start
