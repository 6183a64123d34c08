"""Host-side logic of the multi-GPU path with world_size 2 on the gloo backend
(CPU): partitioning, the sharded-dimension registry and the host collectives the
Vectors layer uses.  The CUDA kernels themselves are row-local and are covered by
tests/dist_check.py on >= 2 GPUs."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    import torch.distributed as tdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    tdist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from raleigh_b200 import dist
    ctx = dist.enable()
    assert ctx.world == world and not ctx.on_device
    n = 1003
    row0, nloc = ctx.register_even(n)
    assert ctx.lookup(n) == (row0, nloc) and ctx.lookup(77) is None
    with pytest.raises(ValueError):
        ctx.register(n, row0 + 1, nloc)
    rng = np.random.RandomState(0)
    x = rng.randn(5, n)
    y = rng.randn(7, n)
    xl, yl = x[:, row0:row0 + nloc], y[:, row0:row0 + nloc]
    g = ctx.allreduce_host(yl @ xl.T)                 # what Vectors.dot does for a sharded block
    assert np.allclose(g, y @ x.T, atol=1e-10)
    counts = ctx.allgather_counts(nloc)
    assert sum(counts) == n and counts[rank] == nloc
    full = ctx.allgather_columns(xl, counts)          # Vectors.data()
    assert np.array_equal(full, x)
    # sample-partitioned dense operator: partial products + all-reduce == global product
    a = rng.randn(n, 31)
    al = a[row0:row0 + nloc]
    z = ctx.allreduce_host(xl @ al)
    assert np.allclose(z, x @ a, atol=1e-9)
    # halo plan of a row-partitioned sparse operator: pack -> exchange -> local product
    import scipy.sparse as sp
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle import algebra_np as K
    L = K.lap3d_csr(7, 6, 5)
    R = sp.random(210, 210, density=0.03, random_state=4, format='csr')
    for A in (L, (L + R + R.T).tocsr()):
        ng = A.shape[0]
        r0, nl = dist.partition(ng, world, rank)
        slab = A[r0:r0 + nl].tocsr()
        slab.sort_indices()
        plan = dist.HaloPlan(ctx, slab.indptr, slab.indices, r0, nl, ng)
        assert sum(plan.recv_counts) == plan.nhalo and plan.recv_counts[rank] == 0 and plan.send_counts[rank] == 0
        xg = rng.randn(3, ng)
        xloc = xg[:, r0:r0 + nl]
        m = 3
        send = torch.from_numpy(np.ascontiguousarray(xloc[:, plan.send_idx].T).reshape(-1))   # row-interleaved pack
        recv = torch.empty(plan.nhalo * m, dtype=torch.float64)
        plan.exchange(send, recv, m)
        halo = recv.numpy().reshape(plan.nhalo, m).T                                          # (m, nhalo)
        xext = np.concatenate([xloc, halo], axis=1)
        Aloc = sp.csr_matrix((slab.data, plan.local_indices, plan.indptr), shape=(nl, nl + plan.nhalo))
        yloc = (Aloc @ xext.T).T
        assert np.allclose(yloc, (A @ xg.T).T[:, r0:r0 + nl], atol=1e-9)
    # append(axis=1) of two row-sharded blocks (vectors.py: append) registers the joined dimension process-major:
    # rows of process 0 of both blocks, then process 1, ... -- a contiguous partition again, and data() of the
    # joined block is the process-major concatenation
    n1 = 400
    r1, l1 = ctx.register_even(n1)
    ctx.register(n + n1, row0 + r1, nloc + l1)
    joined_counts = ctx.allgather_counts(nloc + l1)
    assert sum(joined_counts) == n + n1 and sum(joined_counts[:rank]) == row0 + r1
    xb = rng.randn(5, n1)
    mine = np.hstack((xl, xb[:, r1:r1 + l1]))
    gathered = ctx.allgather_columns(mine, joined_counts)
    expect = np.hstack([np.hstack((x[:, a:a + b], xb[:, c:c + d]))
                        for (a, b), (c, d) in ((dist.partition(n, world, r), dist.partition(n1, world, r)) for r in range(world))])
    assert np.array_equal(gathered, expect)
    # Matrix Market file -> per-process row slab -> halo plan (io.read_matrix_market; nobody holds the whole matrix)
    import scipy.io
    from raleigh_b200 import io as rio
    path = os.path.join(out['tmp'], 'lap.mtx')
    if rank == 0:
        scipy.io.mmwrite(path, sp.triu(L).tocoo(), symmetry='general')     # upper triangle on disk ...
        with open(path) as f:
            text = f.read().replace('general', 'symmetric', 1)             # ... declared symmetric
        with open(path, 'w') as f:
            f.write(text)
    tdist.barrier()
    ng = L.shape[0]
    r0, nl = dist.partition(ng, world, rank)
    slab = rio.read_matrix_market(path, r0, nl, block_lines=50)
    assert abs(slab - L[r0:r0 + nl]).max() == 0.0
    plan = dist.HaloPlan(ctx, slab.indptr, slab.indices, r0, nl, ng)
    assert sum(plan.recv_counts) == plan.nhalo
    out[rank] = 1
    tdist.barrier()
    tdist.destroy_process_group()


def test_partition_covers_rows():
    from raleigh_b200.dist import partition
    for n, w in ((10, 3), (16777216, 8), (5, 8), (12000, 7)):
        parts = [partition(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
        for a, b in zip(parts, parts[1:]):
            assert a[0] + a[1] == b[0]


def test_shard_context_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    import tempfile
    out['tmp'] = tempfile.mkdtemp(prefix='rl_dist_')
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(k for k in out.keys() if k != 'tmp') == [0, 1]
