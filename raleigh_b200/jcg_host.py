"""Host-side bookkeeping of the device-resident block Jacobi-CG driver (jcg.py).

Everything here works on a handful of length-m host arrays (m = block size);
no vector of the eigenproblem's dimension is ever touched.  The decisions are
those of the reference's main loop -- convergence history and kinematic error
estimates (solver.py:922-1007), Lehmann / Davis-Kahan residual estimates
(:1009-1049), cluster detection (:1066-1098), per-side convergence and
stagnation tests (:1100-1195), the choice of the next block layout
(:1495-1541) and the re-indexing of the history when the block window moves
(:1543-1587) -- restated over whole arrays instead of per-element Python loops,
so that the few hundred microseconds they cost do not stall the GPU queue.
"""
import math

import numpy

RECORDS = 100          # length of the eigenvalue-decrement history (solver.py:122)


class BlockLayout:
    """Where the active iterates sit inside the block of m slots.

    Slots [0, left_block) belong to the left margin of the spectrum, the rest
    to the right one; the active X vectors occupy slots [ix, ix + nx), the
    first leftX of them being left iterates (solver.py:760-766)."""

    def __init__(self, m, left_block):
        self.m = m
        self.left_block = left_block
        self.leftX = left_block
        self.rightX = m - left_block
        self.ix = 0
        self.nx = m

    def copy(self):
        other = BlockLayout(self.m, self.left_block)
        other.leftX, other.rightX, other.ix, other.nx = self.leftX, self.rightX, self.ix, self.nx
        return other


def initial_split(m, left, right, largest):
    """Share of the block given to the left margin (solver.py:605-624)."""
    if left == 0 and not largest:
        return 0.0, 1
    if right == 0:
        return 1.0, m - 1
    if left > 0 and right > 0:
        ratio = left / (left + 1.0 * right)
        l = int(round(ratio * m))
        return ratio, min(max(l, 2), m - 2)
    return 0.5, m // 2


class History:
    """Per-slot convergence data of the block (all arrays have m entries)."""

    def __init__(self, m, epsilon):
        self.m = m
        self.epsilon = epsilon
        self.cnv = numpy.zeros((m,), dtype=numpy.int32)
        self.lmd = numpy.zeros((m,), dtype=numpy.float64)
        self.res = -numpy.ones((m,), dtype=numpy.float32)
        self.err_lmd = -numpy.ones((2, m), dtype=numpy.float32)
        self.err_X = -numpy.ones((2, m), dtype=numpy.float32)
        self.iterations = numpy.zeros((m,), dtype=numpy.int32)
        self.dlmd = numpy.zeros((m, RECORDS), dtype=numpy.float32)
        self.dX = numpy.ones((m,), dtype=numpy.float32)
        self.acf = numpy.ones((2, m), dtype=numpy.float32)
        self.cluster = numpy.zeros((2, m), dtype=numpy.int32)
        self.rec = 0
        self.dlmd_floor = [0.0, 0.0]          # stagnation floors frozen at iteration 2 (left, right)
        self._floor_now = [0.0, 0.0]

    # -- solver.py:922-940 ---------------------------------------------------------
    def record_ritz_values(self, ix, new_lmd):
        nx = new_lmd.shape[0]
        s = slice(ix, ix + nx)
        self.iterations[s] += 1
        if self.rec > 0:
            old = self.lmd[s]
            delta = old - new_lmd
            thr = math.sqrt(self.epsilon) * numpy.maximum(abs(old), abs(new_lmd))
            col = self.dlmd[s, self.rec - 1]
            numpy.copyto(col, delta.astype(numpy.float32), where=abs(delta) > thr)
        self.lmd[s] = new_lmd

    # -- solver.py:976-1007 --------------------------------------------------------
    def kinematic_estimates(self, ix, nx):
        rec = self.rec
        if rec <= 3:
            return
        s = slice(ix, ix + nx)
        far = self.dX[s] > 0.01
        self.err_X[0, s][far] = -1.0
        depth = rec // 3 + 1                                   # records rec-1 ... rec-depth
        tail = abs(self.dlmd[s, rec - depth:rec][:, ::-1])     # most recent first
        alive = numpy.cumprod(tail != 0, axis=1).astype(bool)  # stop at the first zero decrement
        k = alive.sum(axis=1)
        sums = numpy.cumsum(numpy.where(alive, tail, numpy.float32(0)), axis=1, dtype=numpy.float32)[:, -1]
        last = tail[:, 0]
        ok = (~far) & (k >= 2) & (sums != 0)
        with numpy.errstate(divide='ignore', invalid='ignore'):
            q = numpy.where(ok, last / numpy.where(ok, sums, 1), 0).astype(numpy.float32)
            ok &= q > 0
            expo = (1.0 / numpy.maximum(k - 1, 1)).astype(numpy.float32)
            q = numpy.where(ok, q ** expo, q).astype(numpy.float32)
            idx = numpy.nonzero(ok)[0] + ix
            self.acf[1, idx] = self.acf[0, idx]
            self.acf[0, idx] = q[ok]
            conv = ok & (q < 1.0)
            theta = q / (1 - q)
            d = theta * self.dlmd[s, rec - 1]
            qx = numpy.sqrt(q)
            ex = self.dX[s] * qx / (1 - qx)
        idx = numpy.nonzero(conv)[0] + ix
        self.err_lmd[0, idx] = abs(d[conv])
        self.err_X[0, idx] = ex[conv]

    # -- solver.py:1009-1049 (standard problems only) ------------------------------
    def residual_estimates(self, lay):
        ix, nx = lay.ix, lay.nx
        lmd, res, dX = self.lmd, self.res, self.dX
        # left margin: the farthest iterate separated from its left neighbour by more than its
        # residual serves as the pole of the Lehmann / Davis-Kahan bounds for those before it
        if lay.leftX > 1:
            ks = numpy.arange(1, lay.leftX)
            i = ix + ks
            stop = numpy.nonzero(dX[i] > 0.01)[0]
            upto = stop[0] if stop.size else ks.size
            good = numpy.nonzero((lmd[i] - lmd[i - 1] > res[i])[:upto])[0]
            if good.size:
                l = int(ks[good[-1]])
                t = lmd[ix + l]
                j = slice(ix, ix + l)
                gap = t - lmd[j]
                with numpy.errstate(divide='ignore', invalid='ignore'):
                    self.err_lmd[1, j] = res[j] * res[j] / gap
                    self.err_X[1, j] = res[j] / gap
        if lay.rightX > 1:
            ks = numpy.arange(1, lay.rightX)
            i = ix + nx - ks - 1
            stop = numpy.nonzero(dX[i] > 0.01)[0]
            upto = stop[0] if stop.size else ks.size
            good = numpy.nonzero((lmd[i + 1] - lmd[i] > res[i])[:upto])[0]
            if good.size:
                l = int(ks[good[-1]])
                t = lmd[ix + nx - l - 1]
                j = slice(ix + nx - l, ix + nx)
                gap = lmd[j] - t
                with numpy.errstate(divide='ignore', invalid='ignore'):
                    self.err_lmd[1, j] = res[j] * res[j] / gap
                    self.err_X[1, j] = res[j] / gap

    # -- solver.py:1066-1098 -------------------------------------------------------
    def update_floors_and_clusters(self, lay, iteration):
        eps = self.epsilon ** 0.67
        lbs, m = lay.left_block, self.m
        last = self.dlmd[:, self.rec - 1]          # rec == 0 reads the (still zero) final record, as the reference does
        if lbs > 0:
            self._floor_now[0] = eps * numpy.amax(abs(last[:lbs]))
        if lbs < m:
            self._floor_now[1] = eps * numpy.amax(abs(last[lbs:]))
        if iteration == 2:
            self.dlmd_floor = list(self._floor_now)
        if iteration < 2:
            return
        cl = self.cluster
        cl[:, :] = 0
        count = 0
        lmd = self.lmd
        for i in range(lbs - 1):
            if abs(lmd[i + 1] - lmd[i]) <= self._floor_now[0]:
                if cl[0, i] == 0:
                    count += 1
                    cl[0, i] = count
                    cl[1, i] = 1
                cl[0, i + 1] = cl[0, i]
                cl[1, i + 1] = cl[1, i] + 1
        for j in range(m - lbs - 1):
            i = m - j - 1
            if abs(lmd[i - 1] - lmd[i]) <= self._floor_now[1]:
                if cl[0, i] == 0:
                    count += 1
                    cl[0, i] = count
                    cl[1, i] = 1
                cl[0, i - 1] = cl[0, i]
                cl[1, i - 1] = cl[1, i] + 1

    # -- solver.py:1100-1195 -------------------------------------------------------
    def count_converged(self, solver, lay, criteria, opts):
        """How many iterates at the left / right end of the active window are done
        (converged by the user's criteria, or stagnated)."""
        left, right, largest, sigma, min_iter, detect, iteration = opts
        ix, nx, rec = lay.ix, lay.nx, self.rec
        cnv, lmd = self.cnv, self.lmd
        lcon = 0
        if left != 0:
            for i in range(lay.leftX - lay.leftX // 4):
                k = ix + i
                if sigma is not None and lmd[k] > 0:
                    break
                it = self.iterations[k]
                if it < min_iter:
                    break
                d1 = abs(self.dlmd[k, max(0, rec - 1)])
                d2 = abs(self.dlmd[k, max(0, rec - 3)])
                if criteria.satisfied(solver, k):
                    lcon += 1
                    cnv[k] = iteration + 1
                elif detect and it > 2 and d1 <= self.dlmd_floor[0] and (d1 > d2 or d1 == 0.0):
                    lcon += 1
                    cnv[k] = -iteration - 1
                else:
                    if self.cluster[0, k] > 0:       # a cluster is locked as a whole or not at all
                        for l in range(k - 1, k - self.cluster[1, k], -1):
                            if cnv[l] == -iteration - 1:
                                cnv[l] = 0
                                lcon -= 1
                    break
        rcon = 0
        if right != 0:
            for i in range(lay.rightX - lay.rightX // 4):
                k = ix + nx - i - 1
                if sigma is not None and lmd[k] < 0:
                    break
                it = self.iterations[k]
                if it < min_iter:
                    break
                d1 = abs(self.dlmd[k, max(0, rec - 1)])
                d2 = abs(self.dlmd[k, max(0, rec - 3)])
                if criteria.satisfied(solver, k):
                    rcon += 1
                    cnv[k] = iteration + 1
                elif detect and it > 2 and d1 <= self.dlmd_floor[1] and (d1 > d2 or d1 == 0.0):
                    rcon += 1
                    cnv[k] = -iteration - 1
                else:
                    if self.cluster[0, k] > 0:
                        for l in range(k + 1, k + self.cluster[1, k]):
                            if cnv[l] == -iteration - 1:
                                cnv[l] = 0
                                rcon -= 1
                    break
        if largest:          # the largest in magnitude must lock first (solver.py:1181-1195)
            if lcon > 0:
                i = ix + lcon - 1
                j = ix + nx - rcon - 1
                while lcon > 0 and abs(lmd[i]) < abs(lmd[j]):
                    cnv[i] = 0
                    lcon -= 1
                    i -= 1
            if rcon > 0:
                i = ix + lcon
                j = ix + nx - rcon
                while rcon > 0 and abs(lmd[i]) > abs(lmd[j]):
                    cnv[j] = 0
                    rcon -= 1
                    j += 1
        return lcon, rcon

    # -- solver.py:1488-1493 -------------------------------------------------------
    def push_record(self, ix, nx, predicted, change):
        if self.rec == RECORDS:
            self.dlmd[:, :-1] = self.dlmd[:, 1:].copy()
        else:
            self.rec += 1
        self.dX[ix:ix + nx] = change
        self.dlmd[ix:ix + nx, self.rec - 1] = predicted

    # -- solver.py:1543-1587 -------------------------------------------------------
    def _move(self, dst, src):
        for a in (self.cnv, self.lmd, self.res, self.dX, self.iterations):
            a[dst] = a[src].copy()
        for a in (self.acf, self.err_lmd, self.err_X):
            a[:, dst] = a[:, src].copy()
        self.dlmd[dst, :] = self.dlmd[src, :].copy()

    def _reset(self, sl):
        self.cnv[sl] = 0
        self.res[sl] = -1.0
        self.acf[:, sl] = 1.0
        self.err_lmd[:, sl] = -1.0
        self.dlmd[sl, :] = 0
        self.err_X[:, sl] = -1.0
        self.dX[sl] = 1.0
        self.iterations[sl] = 0

    def shift(self, left_block, new_left_block, shift_left, shift_right):
        """Slide the per-slot data towards the ends vacated by locked pairs and
        clear the slots that will receive fresh iterates."""
        m, l, nl = self.m, left_block, new_left_block
        if shift_left > 0 and l - shift_left > 0:
            self._move(slice(0, l - shift_left), slice(shift_left, l))
        if shift_left >= 0:
            if nl > l - shift_left:
                self._reset(slice(max(l - shift_left, 0), nl))
        else:
            self._reset(slice(0, l))
        if shift_right > 0 and m - (l + shift_right) > 0:
            self._move(slice(l + shift_right, m), slice(l, m - shift_right))
        if shift_right >= 0:
            if l + shift_right > nl:
                self._reset(slice(nl, l + shift_right))
        else:
            self._reset(slice(l, m))


def next_layout(lay, ny, nxy, lcon, rcon, tot_lcon, tot_rcon, left, right, left_total, right_total,
                left_ratio):
    """Numbers of left and right iterates of the next block (solver.py:1495-1541).
    Returns (new layout, shift_left, shift_right, left_ratio)."""
    m, ix, nx, leftX, rightX = lay.m, lay.ix, lay.nx, lay.leftX, lay.rightX
    if left < 0:
        shift_left = ix
    elif lcon > 0:
        shift_left = min(max(0, left_total - tot_lcon - leftX), ix)
    else:
        shift_left = 0
    if right < 0:
        shift_right = m - ix - nx
    elif rcon > 0:
        shift_right = min(max(0, right_total - tot_rcon - rightX), m - ix - nx)
    else:
        shift_right = 0
    if shift_left + shift_right > ny:
        shift_left = min(shift_left, int(round(left_ratio * ny)))
        shift_right = min(shift_right, ny - shift_left)
    new = lay.copy()
    if left > 0 and lcon > 0 and tot_lcon >= left:          # left margin finished
        l = lay.left_block
        new.leftX = 0
        new.rightX = min(nxy, l + rightX + shift_right)
        new.left_block = l + rightX + shift_right - new.rightX
        shift_left = -leftX - lcon
        left_ratio = 0.0
        new.ix = new.left_block
    elif right > 0 and rcon > 0 and tot_rcon >= right:      # right margin finished
        new.ix = ix - shift_left
        new.leftX = min(nxy, m - new.ix)
        new.rightX = 0
        shift_right = -rightX - rcon
        new.left_block = new.ix + new.leftX
        left_ratio = 1.0
    else:
        new.leftX = leftX + shift_left
        new.rightX = rightX + shift_right
        new.ix = ix - shift_left
    new.nx = new.leftX + new.rightX
    return new, shift_left, shift_right, left_ratio
