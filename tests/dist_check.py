"""Multi-rank check of the row-sharded path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank builds the same global problem from a fixed seed, keeps its row slab,
and compares the sharded results with the global NumPy oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    tdist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import raleigh_b200 as rb
    from raleigh_b200 import dist
    from oracle import algebra_np as K

    ctx = dist.enable()
    rng = np.random.RandomState(3)
    M, N, k = 1003, 517, 24
    a = rng.randn(M, N).astype(np.float32)
    row0, mloc = dist.partition(M, world, rank)
    A = rb.Matrix(np.ascontiguousarray(a[row0:row0 + mloc]))
    assert A.shape() == (M, N), A.shape()
    x = rng.randn(k, N).astype(np.float32)
    X = A.new_vectors(N, k)
    assert not X.is_sharded()
    X.fill(x)
    Y = A.new_vectors(M, k)
    assert Y.is_sharded() and Y.local_dimension() == mloc and Y.dimension() == M
    A.apply(X, Y)
    y = K.dense_apply(a, x)
    assert np.allclose(Y.data(), y, rtol=1e-4, atol=1e-3), 'apply'
    Z = A.new_vectors(N, k)
    A.apply(Y, Z, transp=True)
    z = K.dense_apply(a, y, transp=True)
    assert np.allclose(Z.data(), z, rtol=1e-4, atol=2e-2), 'apply transp (all-reduce)'
    # reductions on sharded vectors
    g = Y.dot(Y)
    assert np.allclose(g, K.gram(y, y), rtol=1e-4, atol=1e-1), 'sharded Gram'
    d = Y.dots(Y)
    assert np.allclose(d, K.row_dots(y, y), rtol=1e-4), 'sharded dots'
    # bitwise identical on all ranks (identical control flow in the replicated solver)
    t = torch.from_numpy(g.copy()).cuda()
    tmax, tmin = t.clone(), t.clone()
    tdist.all_reduce(tmax, op=tdist.ReduceOp.MAX)
    tdist.all_reduce(tmin, op=tdist.ReduceOp.MIN)
    assert torch.equal(tmax, tmin), 'Gram differs between ranks'
    # svd of a sharded block
    W = Y.clone()
    sigma, q = W.svd()
    w = W.data().astype(np.float64)
    assert np.max(np.abs(w @ w.T - np.eye(k))) < 1e-4, 'svd orthonormality'
    recon = (q.astype(np.float64) * sigma[None, :]) @ w
    assert np.linalg.norm(recon - y) / np.linalg.norm(y) < 1e-4, 'svd reconstruction'
    # device fill is partition independent
    F = A.new_vectors(M, 3)
    np.random.seed(5)
    F.fill_random()
    dist.disable()
    np.random.seed(5)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    G = rb.Vectors(M, 3, np.float32)
    G.fill_random_device(seed)
    assert np.array_equal(F.data(), G.data()), 'sharded fill_random'
    ctx = dist.enable()

    # end to end: the reference's pca on the row-sharded matrix vs the golden CPU run
    if rb.find_reference() is not None:
        rb.install()
        from raleigh.interfaces.pca import pca, pca_error
        from raleigh.examples.pca.generate_matrix import generate
        from raleigh.core.solver import Options
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'pca.npz'))
        np.random.seed(1)
        Afull, sig, u, v = generate(600, 400, 200, pca=True)
        r0, ml = dist.partition(600, world, rank)
        np.random.seed(1)
        mean, trans, comps = pca(np.ascontiguousarray(Afull[r0:r0 + ml]), npc=40, arch='gpu!', opt=Options())
        assert trans.shape == (600, comps.shape[0]) and comps.shape[1] == 400
        em, ef = pca_error(Afull, mean, trans, comps)
        assert abs(em - g['small_npc40_err'][0]) < 2e-3 and abs(ef - g['small_npc40_err'][1]) < 2e-3, (em, ef)
        sv = np.linalg.norm(trans, axis=0)
        # sharded blocks start from the device RNG (partition independent), not the host
        # stream of the golden run: agreement is to the solver tolerance (svtol 1e-3)
        exact = np.linalg.svd(Afull.astype(np.float64) - Afull.mean(axis=0, dtype=np.float64)[None, :],
                              compute_uv=False)
        dev_sv = np.max(np.abs(sv[:20] - exact[:20]) / exact[:20])
        dev_gold = np.max(np.abs(g['small_npc40_sv'][:20] - exact[:20]) / exact[:20])
        # the reference's own CPU run is this far from the exact values: (dev_gold);
        # the sharded run must be at least as good up to the solver tolerance (svtol 1e-3)
        assert dev_sv < max(2 * dev_gold, 2e-3), (dev_sv, dev_gold)
        if rank == 0:
            print('sharded pca ok: components %d, pca_error %.3e %.3e, leading singular values within %.1e of exact (reference CPU run: %.1e), '
                  'all-reduces %d (%.1f MB)' % (comps.shape[0], em, ef, dev_sv, dev_gold, ctx.allreduce_calls,
                                                ctx.allreduce_bytes / 1e6))
    tdist.barrier()
    if rank == 0:
        print('DIST_CHECK_OK world=%d' % world)
    tdist.destroy_process_group()


if __name__ == '__main__':
    main()
