"""Device-resident block Jacobi-conjugated-gradient driver.

The reference's main loop (raleigh/core/solver.py:838-1663) alternates a dozen
block-vector operations with O(m^3) host LAPACK on 2m x 2m matrices -- pivoted
Cholesky (`_piv_chol`, :1749-1826), two triangular transforms (`_transform`,
:1685-1688), two symmetric eigenproblems (:1459, :1470) -- and pulls every Gram
matrix back to the host to feed them (about 14 synchronous round trips per
iteration).  With the algebra on a B200 that host work IS the run time.

Here the same iteration (Appendix B of SURVEY.md) keeps every Gram matrix and
every coefficient matrix in device memory:

  * Gram products write fp64 results straight into device-resident small
    matrices (rl_gram_dev), block updates read their coefficients from there
    (rl_update_dev);
  * the pivoted Cholesky factorisation with the reference's drop rule, the
    conjugation coefficients, the Rayleigh-Ritz reduction, both symmetric
    eigenproblems and the back-transformation are hand-written kernels
    (csrc/rr.cu, csrc/jacobi.cu) -- no host LAPACK anywhere;
  * the host sees three small packets per iteration (Ritz values + residual
    norms; the number of dropped search directions; the change estimates) and
    does the convergence bookkeeping on length-m arrays (jcg_host.py);
  * the block buffers are rotated instead of copied: the reference moves every
    updated block through the W workspace and copies it back (:1610-1656).

Supported: standard problems A x = lambda x and generalised problems
A x = lambda B x and A B x = lambda x (B symmetric positive definite: the
B-images BX, BY, BZ of the blocks are carried along, Gram matrices are B-Gram
matrices; the product form applies A to the images and measures residuals in
the B-norm), optional preconditioner, optional previously computed
eigenvectors, real float32/float64.

The driver talks to an `engine` (engine.py: DeviceEngine, ctypes over the C
ABI).  The tests substitute a NumPy engine to check the control flow against
the reference solver on the CPU; the product path never does.
"""
import math

import numpy

from .jcg_host import BlockLayout, History, initial_split, next_layout


class _Fatal(Exception):
    pass


def supported(solver, eigenvectors):
    """True when the device-resident driver can run this problem."""
    problem = solver.problem()
    if problem.type() not in ('s', 'g', 'p'):
        return False
    return hasattr(eigenvectors, '_rl_device_block')


def _parse_which(which):
    try:
        if len(which) != 2:
            raise ValueError('which must be either integer or tuple of 2 integers')
        return False, int(which[0]), int(which[1])
    except TypeError:
        return True, which, which


class _Pool:
    """Block buffers of the iteration; roles are rotated instead of copying."""

    def __init__(self, engine, count):
        self._free = [engine.new_block() for _ in range(count)]

    def take(self, nv):
        blk = self._free.pop()
        blk.select(nv)
        return blk

    def give(self, *blocks):
        for b in blocks:
            if b is not None:
                self._free.append(b)


def solve(solver, eigenvectors, options, which, extra, init, engine):
    """Drop-in for Solver._solve (solver.py:587-1665): same arguments, same
    attributes set on `solver`, same return codes."""
    verb = options.verbosity
    sigma = options.sigma
    largest, left, right = _parse_which(which)
    m = solver.block_size
    left_ratio, left_block = initial_split(m, left, right, largest)
    lay = BlockLayout(m, left_block)

    extra_left, extra_right = int(extra[0]), int(extra[1])
    left_total = right_total = 0
    if left >= 0:
        left_total = left + extra_left if extra_left > 0 else max(left + 1, left_block)
    if right >= 0:
        right_total = right + extra_right if extra_right > 0 else max(right + 1, m - left_block)

    problem = solver.problem()
    vector = problem.vector()
    data_type = vector.data_type()
    epsilon = float(numpy.finfo(data_type).eps)
    single = data_type in (numpy.float32, numpy.complex64)
    chol_eps = 1e-3 if single else 1e-8          # solver.py:1404-1407

    hist = History(m, epsilon)
    solver.cnv, solver.lmd, solver.res = hist.cnv, hist.lmd, hist.res
    solver.err_lmd, solver.err_X = hist.err_lmd, hist.err_X
    criteria = options.convergence_criteria
    if criteria is None:
        criteria = _default_criteria()
    view = _SolverView(solver)

    opA = problem.A()
    gen = problem.type() == 'g'                   # A x = lambda B x
    pro = problem.type() == 'p'                   # A B x = lambda x
    hasB = gen or pro
    opB = problem.B() if hasB else None
    opP = solver.preconditioner()
    eng = engine
    eng.begin(vector, m)
    pool = _Pool(eng, 10 if hasB else 7)          # with B: the B-images of X, Y, Z as well

    # ---- initial block (solver.py:676-723) ---------------------------------------
    X = pool.take(m)
    X.fill_random()                      # host RNG stream: seeded runs start like the reference's
    l = left_block
    init_left = 0
    if init[0] is not None:
        init_left = min(l, init[0].nvec())
        X.select(init_left)
        init[0].select(init_left)
        init[0].copy(X)
    if init[1] is not None:
        init_right = min(m - l, init[1].nvec())
        X.select(init_right, init_left)
        init[1].select(init_right)
        init[1].copy(X)
    X.select(m)
    s = X.dots(X)
    for i in numpy.nonzero(s == 0.0)[0]:
        if verb > -1:
            print('Zero initial guess, replacing with random')
        X.select(1, int(i))
        X.fill_random()
    X.select(m)
    eng.dots(X, X, eng.v_s2)
    eng.scale_rsqrt(X, eng.v_s2)

    # ---- constraints: eigenvectors already in the container (solver.py:743-775) ---
    solver.eigenvectors = eigenvectors
    Xc = eigenvectors
    nc = Xc.nvec()
    if hasB:
        BXc = eigenvectors.clone()
        if nc > 0:
            opB.apply(Xc, BXc)
        solver.eigenvectors_im = BXc
    else:
        BXc = Xc
    eng.reserve_constraints(nc + m)
    if nc > 0:
        eng.gram(BXc, Xc, eng.Gc.sub(0, 0, nc, nc))
        _project_out(eng, X, BXc, Xc, nc, m)

    # ---- drop linearly dependent initial vectors (solver.py:779-813) -------------
    nx = m
    BX = X
    if hasB:
        BX = pool.take(m)
        opB.apply(X, BX)
    eng.gram(BX, X, eng.GB.sub(0, 0, m, m))
    eng.piv_chol(eng.GB, m, 0, 1e-2)
    dropped, ind = eng.fetch_chol(m)
    if dropped > 0:
        if verb > 0:
            print('dropped %d initial vectors out of %d' % (dropped, nx))
        nx -= dropped
        T = pool.take(m)
        if nx > 0:
            T.select(nx)
            eng.gather(X, ind[:nx], T)
        T.select(dropped, nx)
        T.fill_random()
        if nc > 0:
            _project_out(eng, T, BXc, Xc, nc, dropped)
        T.select(m)
        pool.give(X)
        X = T
        nx = m
        if hasB:
            opB.apply(X, BX)      # the reference gathers the kept images and projects the new ones: same block
        else:
            BX = X

    # ---- Rayleigh-Ritz in the initial space (solver.py:815-830) ------------------
    AX = pool.take(m)
    opA.apply(BX if pro else X, AX)
    eng.gram(BX, X, eng.GB.sub(0, 0, m, m))
    eng.gram(AX, BX if pro else X, eng.GA.sub(0, 0, m, m))
    eng.ritz_initial(m)                   # generalised m x m problem -> coefficients CX, Ritz values lmdx
    if hasB:
        X, AX, BX = _rotate(eng, pool, (X, AX, BX), m, m)
    else:
        X, AX = _rotate(eng, pool, (X, AX), m, m)
        BX = X

    # ---- main loop ------------------------------------------------------------------
    max_iter = options.max_iter
    min_iter = options.min_iter
    if max_iter < 0:
        max_iter = 100
    solver.iteration = 0
    Z = AZ = BZ = None
    nz = 0
    W = None

    while True:
        maxit = 0
        if left != 0 and lay.left_block > 0:
            maxit = numpy.amax(hist.iterations[:lay.left_block])
        if right != 0 and lay.left_block < m:
            maxit = max(maxit, numpy.amax(hist.iterations[lay.left_block:]))
        if maxit >= max_iter:
            if verb > -1:
                print('iterations limit of %d exceeded, terminating' % max_iter)
            break
        if verb > 0:
            print('------------- iteration %d' % solver.iteration)

        nx, ix = lay.nx, lay.ix
        X.select(nx)
        AX.select(nx)
        BX.select(nx)

        # Rayleigh quotients, orthonormality check, residuals (solver.py:854-974)
        eng.gram(AX, BX if pro else X, eng.XAX.sub(0, 0, nx, nx))
        eng.gram(BX, X, eng.XBX.sub(0, 0, nx, nx))
        eng.ritz_check(nx)                                   # -> v_lmd, rv_err, rv_no
        W = pool.take(nx)
        BW = pool.take(nx) if pro else None                  # product form: B-image of the residuals
        _residuals(eng, W, BW, X, BX, AX, Xc, BXc, nc, nx, gen, opB if pro else None)
        # Search directions for the whole (old) window -- preconditioning and conjugation to the previous directions
        # (solver.py:1315-1351) -- depend on nothing the host is about to decide: they are queued BEFORE the host
        # reads the Ritz packet, so the device works on them while the host does its convergence bookkeeping.
        Y = _directions(eng, pool, opP, pro, hasB, W, BW, Z, AZ, BZ, nz, nx)
        new_lmd, res2, rv_err, rv_no = eng.fetch_ritz(nx)
        if verb > 2:
            print('Ritz values error: %.1e' % rv_err)
            print('Ritz vectors non-orthonormality: %.1e' % rv_no)
        if max(rv_err, rv_no) > math.sqrt(epsilon) or not numpy.all(numpy.isfinite(new_lmd)):
            if verb > 0:
                print('restarting...')
            hist.rec = 0
            nz = 0
            X, AX, BX = _restart(eng, pool, opA, opB, X, AX, BX, nx, pro)
            eng.gram(AX, BX if pro else X, eng.XAX.sub(0, 0, nx, nx))
            eng.gram(BX, X, eng.XBX.sub(0, 0, nx, nx))
            eng.ritz_check(nx)
            W.select(nx)
            _residuals(eng, W, BW, X, BX, AX, Xc, BXc, nc, nx, gen, opB if pro else None)
            if Y is not W:
                pool.give(Y)
            Y = _directions(eng, pool, opP, pro, hasB, W, BW, None, None, None, 0, nx)      # nz = 0: no conjugation
            new_lmd, res2, rv_err, rv_no = eng.fetch_ritz(nx)

        hist.record_ritz_values(ix, new_lmd)
        hist.res[ix:ix + nx] = numpy.sqrt(abs(res2))
        hist.kinematic_estimates(ix, nx)
        if not gen:                          # Lehmann / Davis-Kahan bounds: not valid for A x = lambda B x (solver.py:1009)
            hist.residual_estimates(lay)
        hist.update_floors_and_clusters(lay, solver.iteration)
        if verb > 1:
            _print_table(solver, hist, m)

        view.refresh()
        lcon, rcon = hist.count_converged(
            view, lay, criteria,
            (left, right, largest, sigma, min_iter, options.detect_stagnation, solver.iteration))

        # lock converged pairs (solver.py:1197-1270)
        if lcon > 0:
            _record_converged(solver, hist, ix, ix + lcon)
            nc = _lock(eng, X, BX, Xc, BXc, nc, 0, lcon, hasB)
        if rcon > 0:
            _record_converged(solver, hist, ix + nx - rcon, ix + nx)
            nc = _lock(eng, X, BX, Xc, BXc, nc, nx - rcon, rcon, hasB)
        solver.lcon += lcon
        solver.rcon += rcon

        # stopping tests (solver.py:1272-1298)
        if options.stopping_criteria is not None and options.stopping_criteria.satisfied(solver):
            return 0
        if largest and right > 0 and solver.lcon + solver.rcon >= right:
            return 0
        left_done = left >= 0 and solver.lcon >= left
        right_done = right >= 0 and solver.rcon >= right
        if left_done and right_done:
            return 0
        if sigma is not None:
            lmd, err_lmd = hist.lmd, hist.err_lmd
            if right_done:
                i = ix + lcon
                if lmd[i] > 0 and err_lmd[0, i] != -1.0 and err_lmd[0, i] < lmd[i] / 4:
                    return 4
            if left_done:
                i = ix + nx - rcon - 1
                if lmd[i] < 0 and err_lmd[0, i] != -1.0 and err_lmd[0, i] < -lmd[i] / 4:
                    return 5
        if eigenvectors.nvec() > options.max_quota * eigenvectors.dimension():
            return 1

        # shrink the active window (solver.py:1300-1313)
        iy, ny = ix, nx                      # the residual block still has one vector per OLD iterate
        lay.leftX -= lcon
        lay.rightX -= rcon
        lay.ix += lcon
        lay.nx -= lcon + rcon
        ix, nx = lay.ix, lay.nx
        x0 = lcon                            # first active vector inside the (compact) device blocks
        X.select(nx, x0)
        AX.select(nx, x0)
        BX.select(nx, x0)

        # search directions: computed above, one per OLD iterate (solver.py:1315-1351)
        if Y is not W:
            pool.give(W)
        W = None
        Y.select(ny)

        # orthogonalise to X and to the locked vectors, normalise (solver.py:1360-1381)
        if pro:
            BW.select(ny)
        if nx > 0:
            eng.gram(Y, BX, eng.T1.sub(0, 0, nx, ny))
            eng.update(Y, X, eng.T1.sub(0, 0, nx, ny), -1.0, 1.0)
            if pro:
                eng.update(BW, BX, eng.T1.sub(0, 0, nx, ny), -1.0, 1.0)
        if nc > 0:
            _project_out(eng, Y, BXc, Xc, nc, ny)
            if pro:                              # same coefficients, applied to the image block
                BXc.select(nc)
                eng.update(BW, BXc, eng.QC.sub(0, 0, nc, ny), -1.0, 1.0)
        if pro:
            BY = BW
            BW = None
            eng.dots(BY, Y, eng.v_s2)
            eng.scale_rsqrt(Y, eng.v_s2)
            eng.scale_rsqrt(BY, eng.v_s2)
        elif gen:
            BY = pool.take(ny)
            opB.apply(Y, BY)
            eng.dots(BY, Y, eng.v_s2)
            eng.scale_rsqrt(Y, eng.v_s2)
            eng.scale_rsqrt(BY, eng.v_s2)
        else:
            BY = Y
            eng.dots(Y, Y, eng.v_s2)
            eng.scale_rsqrt(Y, eng.v_s2)

        # Gram matrix of (X, Y) and its pivoted Cholesky factor (solver.py:1375-1435)
        nxy = nx + ny
        if nx > 0:
            eng.copy_small(eng.XBX.sub(x0, x0, nx, nx), eng.GB.sub(0, 0, nx, nx))
            eng.gram(BY, X, eng.GB.sub(0, nx, nx, ny))
        eng.gram(BY, Y, eng.GB.sub(nx, nx, ny, ny))
        eng.mirror_upper(eng.GB, nx, ny)
        eng.piv_chol(eng.GB, nxy, nx, chol_eps)
        # The operator is applied to ALL candidate directions before the host learns which of them survive: the
        # device works through A Y while the host waits for the pivot packet, instead of idling until the host has
        # read it and launched the gather.  Column j of A Y depends on column j of Y only, so gathering A Y by the
        # pivot order afterwards gives the same block as applying A to the gathered Y (solver.py:1424-1440).
        AYf = pool.take(ny)
        opA.apply(BY if pro else Y, AYf)
        dropped, ind = eng.fetch_chol(nxy)
        if dropped > 0 and verb > 0:
            print('dropped %d search directions out of %d' % (dropped, ny))
        ny -= dropped
        if ny < 1:
            if verb > -1:
                print('no search directions left, terminating')
            return 3
        nxy = nx + ny
        Yp = pool.take(ny)
        eng.gather(Y, ind[nx:nxy] - nx, Yp)
        pool.give(Y)
        Y = Yp
        if hasB:
            BYp = pool.take(ny)
            eng.gather(BY, ind[nx:nxy] - nx, BYp)
            pool.give(BY)
            BY = BYp
        else:
            BY = Y

        AY = pool.take(ny)
        AYf.select(ny + dropped)
        eng.gather(AYf, ind[nx:nxy] - nx, AY)
        pool.give(AYf)

        # A-Gram matrix of (X, Y) (solver.py:1437-1454)
        if nx > 0:
            eng.copy_small(eng.XAX.sub(x0, x0, nx, nx), eng.GA.sub(0, 0, nx, nx))
            eng.gram(AY, BX if pro else X, eng.GA.sub(0, nx, nx, ny))
        eng.gram(AY, BY if pro else Y, eng.GA.sub(nx, nx, ny, ny))
        eng.mirror_upper(eng.GA, nx, ny)

        # next block layout: integers only, known before the Ritz problem is solved
        new, shift_left, shift_right, left_ratio = next_layout(
            lay, ny, nxy, lcon, rcon, solver.lcon, solver.rcon, left, right, left_total, right_total,
            left_ratio)
        if verb > 2:
            print('left X: was %d, now %d' % (lay.leftX, new.leftX))
            print('right X: was %d, now %d' % (lay.rightX, new.rightX))
            print('new ix %d, new nx %d, nxy %d' % (new.ix, new.nx, nxy))
        nz_new = nxy - new.leftX - new.rightX

        # Rayleigh-Ritz (solver.py:1456-1493, 1589-1607), all on the device
        eng.rayleigh_ritz(nx, ny, lay.leftX, lay.rightX, new.leftX, new.rightX)

        # new X, Z and their images (solver.py:1609-1656); buffers rotate, nothing is copied back
        nxn = new.nx
        Xn = pool.take(nxn)
        _combine(eng, Xn, X, Y, eng.CX, nx, ny, nxn)
        if nz_new > 0:
            if Z is None:
                Z, AZ = pool.take(nz_new), pool.take(nz_new)
            Z.select(nz_new)
            AZ.select(nz_new)
            _combine(eng, Z, X, Y, eng.CZ, nx, ny, nz_new)
        pool.give(X, Y)
        AXn = pool.take(nxn)
        _combine(eng, AXn, AX, AY, eng.CX, nx, ny, nxn)
        if nz_new > 0:
            _combine(eng, AZ, AX, AY, eng.CZ, nx, ny, nz_new)
        pool.give(AX, AY)
        if hasB:
            BXn = pool.take(nxn)
            _combine(eng, BXn, BX, BY, eng.CX, nx, ny, nxn)
            if nz_new > 0:
                if BZ is None:
                    BZ = pool.take(nz_new)
                BZ.select(nz_new)
                _combine(eng, BZ, BX, BY, eng.CZ, nx, ny, nz_new)
            pool.give(BX, BY)
            BX = BXn
        else:
            BX = Xn
        X, AX = Xn, AXn
        # the change estimates of the Ritz step are host bookkeeping only: read them AFTER the block updates are
        # queued (coefficients and layout are already on the device / known), so the wait overlaps those kernels
        change, predicted = eng.fetch_estimates(nx)
        hist.push_record(ix, nx, predicted, change)
        hist.shift(lay.left_block, new.left_block, shift_left, shift_right)
        nz = nz_new
        lay = new
        solver.iteration += 1

    return 2


# ---------------------------------------------------------------------------------------
class _SolverView:
    """What the convergence criteria see as `solver`: the Solver object itself, with `convergence_data`
    (solver.py:333-387) answered from values cached once per iteration.  The reference's method re-scans
    `lmd` and all converged eigenvalues and re-parses its `what` string on every call -- three calls per tested
    pair from lra.py:458-463, 22 ms of host time per config-2 solve with the GPU idle.  Same values, same
    semantics; every other attribute is the Solver's own."""

    _KINDS = {}

    def __init__(self, solver):
        object.__setattr__(self, '_solver', solver)
        object.__setattr__(self, '_max_lmd', None)

    def __getattr__(self, name):
        return getattr(self._solver, name)

    def __setattr__(self, name, value):
        setattr(self._solver, name, value)

    def refresh(self):
        s = self._solver
        m = numpy.amax(abs(s.lmd))
        if s.lcon + s.rcon > 0:
            m = max(m, numpy.amax(abs(s.eigenvalues)))
        object.__setattr__(self, '_max_lmd', m)

    @classmethod
    def _kind(cls, what):
        k = cls._KINDS.get(what)
        if k is None:
            # the decision tree of solver.py:343-387, evaluated once per distinct string
            if what.find('block') > -1:
                k = 'block'
            elif what.find('res') > -1 and what.find('vec') == -1:
                k = 'res'
            elif what.find('val') > -1:
                if what.find('max') > -1:
                    k = 'max'
                elif what.find('err') > -1:
                    k = 'val_err_k' if what.find('k') else 'val_err_r'      # sic: `if what.find('k')` (solver.py:364)
                else:
                    k = 'val'
            elif what.find('vec') > -1:
                k = 'vec_k' if what.find('k') > -1 else 'vec_r'
            else:
                k = 'unknown'
            cls._KINDS[what] = k
        return k

    def convergence_data(self, what='residual', which=0):
        s = self._solver
        k = self._kind(what)
        if k == 'res':
            return s.res[which] / self._max_lmd
        if k == 'val':
            return s.lmd[which]
        if k == 'max':
            return self._max_lmd
        if k == 'vec_k':
            return s.err_X[0, which]
        if k == 'vec_r':
            return s.err_X[1, which]
        if k == 'val_err_k':
            return s.err_lmd[0, which]
        if k == 'val_err_r':
            return s.err_lmd[1, which]
        if k == 'block':
            return s.block_size
        raise ValueError('convergence data %s not found' % what)


def _default_criteria():
    class _Kinematic:
        tolerance = 1e-3
        error = 'kinematic eigenvector error'

        def satisfied(self, solver, i):
            err = solver.convergence_data(self.error, i)
            return err >= 0 and err <= self.tolerance
    return _Kinematic()


def _project_out(eng, V, Against, Sub, nc, nv):
    """V <- V - Sub (2I - Gc) (Against^T V)  (solver.py:774-775, 962-966, 1370-1371).  Standard problem:
    Against = Sub = Xc.  Generalised: iterates and search directions are projected with Against = B Xc,
    Sub = Xc; residuals with Against = Xc, Sub = B Xc."""
    Against.select(nc)
    Sub.select(nc)
    T = eng.TC.sub(0, 0, nc, nv)
    Q = eng.QC.sub(0, 0, nc, nv)
    eng.gram(V, Against, T)
    eng.constraint_coeffs(nc, nv)
    eng.update(V, Sub, Q, -1.0, 1.0)


def _residuals(eng, W, BW, X, BX, AX, Xc, BXc, nc, nx, gen, opB_pro):
    """Residuals projected off the locked vectors, squared norms -> v_s2 (solver.py:942-974).
    std: W = A X - X lmd, W -= Xc Gci Xc^T W.   gen: W = A X - B X lmd, W -= B Xc Gci Xc^T W.
    pro: W = A B X - X lmd, W -= Xc Gci (B Xc)^T W, BW = B W, norms in the B inner product."""
    eng.residual(W, AX, BX if gen else X, eng.v_lmd)
    if nc > 0:
        if opB_pro is not None:
            _project_out(eng, W, BXc, Xc, nc, nx)
        else:
            _project_out(eng, W, Xc, BXc if gen else Xc, nc, nx)
    if opB_pro is not None:
        BW.select(nx)
        opB_pro.apply(W, BW)
        eng.dots(BW, W, eng.v_s2)
    else:
        eng.dots(W, W, eng.v_s2)


def _directions(eng, pool, opP, pro, hasB, W, BW, Z, AZ, BZ, nz, ny):
    """Y = T W (or W itself), conjugated to the previous directions Z: Y -= Z Beta with Beta from
    (Z^T A Y - Z^T B Y lmd) / (lmdz - lmd) (solver.py:1315-1351; rl_rr_conjugation).  Product form: no
    preconditioner, A-products taken with the image block BW, which is updated alongside."""
    if opP is None or pro:
        Y = W
    else:
        Y = pool.take(ny)
        W.select(ny)
        opP.apply(W, Y)
    Y.select(ny)
    if nz > 0:
        Z.select(nz)
        AZ.select(nz)
        if hasB:
            BZ.select(nz)
        if pro:
            BW.select(ny)
        eng.gram(BW if pro else Y, AZ, eng.ZAY.sub(0, 0, nz, ny))
        eng.gram(Y, BZ if hasB else Z, eng.ZBY.sub(0, 0, nz, ny))
        eng.dots(Y, Y, eng.v_y2)              # not v_s2: the residual norms in the Ritz packet are still unread
        eng.dots(Z, Z, eng.v_t2)
        eng.conjugation(nz, ny)               # uses the Ritz values of the OLD window, v_lmd[0:ny]
        eng.update(Y, Z, eng.Beta.sub(0, 0, nz, ny), -1.0, 1.0)
        if pro:
            eng.update(BW, BZ, eng.Beta.sub(0, 0, nz, ny), -1.0, 1.0)
    return Y


def _rotate(eng, pool, blocks, k, mout):
    """blocks <- blocks . CX[:k, :mout] through fresh buffers (no copy back)."""
    out = []
    for blk in blocks:
        new = pool.take(mout)
        blk.select(k)
        eng.update(new, blk, eng.CX.sub(0, 0, k, mout), 1.0, 0.0)
        pool.give(blk)
        out.append(new)
    return out


def _combine(eng, out, X, Y, C, nx, ny, mout):
    """out = X . C[:nx] + Y . C[nx:nx+ny]."""
    out.select(mout)
    if nx > 0:
        eng.update(out, X, C.sub(0, 0, nx, mout), 1.0, 0.0)
        eng.update(out, Y, C.sub(nx, 0, ny, mout), 1.0, 1.0)
    else:
        eng.update(out, Y, C.sub(0, 0, ny, mout), 1.0, 0.0)


def _restart(eng, pool, opA, opB, X, AX, BX, nx, pro=False):
    """Loss of orthonormality among the iterates (solver.py:877-920): orthonormalise X
    by its SVD, recompute AX (and BX) and redo the Rayleigh-Ritz procedure in span(X)."""
    X.select(nx)
    X.svd()
    AX.select(nx)
    if opB is not None:
        BX.select(nx)
        opB.apply(X, BX)
        eng.gram(BX, X, eng.GB.sub(0, 0, nx, nx))
    else:
        eng.gram(X, X, eng.GB.sub(0, 0, nx, nx))
    opA.apply(BX if pro else X, AX)
    eng.gram(AX, BX if pro else X, eng.GA.sub(0, 0, nx, nx))
    eng.ritz_initial(nx)
    if opB is not None:
        return _rotate(eng, pool, (X, AX, BX), nx, nx)
    X, AX = _rotate(eng, pool, (X, AX), nx, nx)
    return X, AX, X


def _record_converged(solver, hist, i0, i1):
    solver.eigenvalues = numpy.concatenate((solver.eigenvalues, hist.lmd[i0:i1]))
    solver.eigenvalue_errors.append(hist.err_lmd[:, i0:i1])
    solver.eigenvector_errors.append(hist.err_X[:, i0:i1])
    solver.residual_norms = numpy.concatenate((solver.residual_norms, hist.res[i0:i1]))
    solver.convergence_status = numpy.concatenate((solver.convergence_status, hist.cnv[i0:i1]))


def _lock(eng, X, BX, Xc, BXc, nc, first, count, gen):
    """Append `count` iterates starting at `first` to the locked set (and their B-images to its image) and
    extend the B-Gram matrix of the set by the new row and column blocks (solver.py:1208-1230)."""
    eng.reserve_constraints(nc + count)
    X.select(count, first)
    Xc.select(nc)
    if gen:
        BXc.select(nc)
    if nc > 0:
        eng.gram(X, BXc, eng.Gc.sub(0, nc, nc, count))
    Xc.append(X)
    nc_new = nc + count
    Xc.select(nc_new)
    if gen:
        BX.select(count, first)
        BXc.append(BX)
        BXc.select(nc_new)
    eng.gram(BXc, X, eng.Gc.sub(nc, 0, count, nc_new))
    return nc_new


def _print_table(solver, hist, m):
    print('  eigenvalue   residual   estimated errors (kinematic/residual)      a.c.f.')
    print('                             eigenvalue            eigenvector ')
    for i in range(m):
        print('%14e %8.1e  %8.1e / %8.1e    %.1e / %.1e  %.3e  %d' %
              (hist.lmd[i], hist.res[i], hist.err_lmd[0, i], hist.err_lmd[1, i],
               abs(hist.err_X[0, i]), abs(hist.err_X[1, i]), hist.acf[0, i], hist.cnv[i]))
