"""Row-sharding of block vectors across the GPUs of one node (one process per
GPU, torch.distributed over NCCL/NVLink; gloo on CPU for the host-side logic).

The reference has no distributed layer at all; its solver
only needs the reductions of the Vectors contract to be global.  A Vectors
object whose LOGICAL dimension is registered here as sharded holds rows
[row0, row0 + nloc) of every vector; `dot`, `dots` and the Gram matrices behind
`svd`/`orthogonalize` are combined with an all-reduce(sum), which returns the
same bits on every rank, so the replicated host logic of the solver takes
identical decisions everywhere.  Everything else (copy, scale, add, multiply,
SpMM interior) is row-local.
"""
import numpy
import torch
import torch.distributed as tdist

_current = None


def partition(n_global, world, rank):
    """Contiguous block partition: the first n % world ranks own one extra row."""
    base, rem = divmod(int(n_global), int(world))
    row0 = rank * base + min(rank, rem)
    return row0, base + (1 if rank < rem else 0)


class ShardContext:
    def __init__(self, group=None, shard_matrices=True):
        self.shard_matrices = shard_matrices   # Matrix(ndarray) = this rank's ROWS of a row-sharded matrix
        if not tdist.is_initialized():
            raise RuntimeError('torch.distributed is not initialised')
        self.group = group
        self.rank = tdist.get_rank(group)
        self.world = tdist.get_world_size(group)
        self.backend = tdist.get_backend(group)
        self.on_device = self.backend == 'nccl'
        self.sharded_dims = {}        # logical (global) dimension -> (row0, nloc)
        self.allreduce_calls = 0
        self.allreduce_bytes = 0

    # ---- registry -----------------------------------------------------------------
    def register(self, n_global, row0, nloc):
        prev = self.sharded_dims.get(int(n_global))
        if prev is not None and prev != (int(row0), int(nloc)):
            raise ValueError('dimension %d is already sharded differently' % n_global)
        self.sharded_dims[int(n_global)] = (int(row0), int(nloc))

    def register_even(self, n_global):
        row0, nloc = partition(n_global, self.world, self.rank)
        self.register(n_global, row0, nloc)
        return row0, nloc

    def lookup(self, n_global):
        return self.sharded_dims.get(int(n_global))

    # ---- collectives --------------------------------------------------------------
    def allreduce_(self, tensor):
        """In-place sum over ranks of a torch tensor (device tensor with NCCL)."""
        self.allreduce_calls += 1
        self.allreduce_bytes += tensor.numel() * tensor.element_size()
        tdist.all_reduce(tensor, op=tdist.ReduceOp.SUM, group=self.group)
        return tensor

    def allreduce_host(self, array):
        """Sum over ranks of a host ndarray (goes through the device for NCCL)."""
        t = torch.from_numpy(numpy.ascontiguousarray(array))
        if self.on_device:
            t = t.cuda()
        self.allreduce_(t)
        return t.cpu().numpy()

    def allgather_counts(self, value):
        t = torch.tensor([int(value)], dtype=torch.int64)
        if self.on_device:
            t = t.cuda()
        out = [torch.zeros_like(t) for _ in range(self.world)]
        tdist.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    def allgather_columns(self, local, counts):
        """Concatenate along axis 1 the (m, nloc_r) host blocks of all ranks."""
        m = local.shape[0]
        width = max(counts) if counts else 0
        pad = numpy.zeros((m, width), dtype=local.dtype)
        pad[:, :local.shape[1]] = local
        t = torch.from_numpy(pad)
        if self.on_device:
            t = t.cuda()
        out = [torch.empty_like(t) for _ in range(self.world)]
        tdist.all_gather(out, t, group=self.group)
        return numpy.concatenate([o.cpu().numpy()[:, :c] for o, c in zip(out, counts)], axis=1)

    def allgather_columns_device(self, ptr, ld_elems, m, nloc, np_dtype, counts):
        """Same result as allgather_columns, starting from a pitched DEVICE block (m rows of nloc elements, row
        pitch ld_elems): pack on the device, NCCL all-gather, join on the device, ONE download -- instead of
        download, pad, upload, all-gather, world downloads and a host concatenation."""
        from ._lib import lib, check
        from . import device as dev
        dt = numpy.dtype(np_dtype)
        tdt = torch.float32 if dt == numpy.float32 else torch.float64
        code = 0 if dt == numpy.float32 else 1
        width = max(counts) if counts else 0
        total = int(sum(counts))
        if m < 1 or total < 1:
            return numpy.zeros((m, total), dtype=dt)
        mine = torch.zeros((m, width), dtype=tdt, device='cuda')
        if nloc > 0:
            check(lib.rl_copy(code, mine.data_ptr(), width, ptr, ld_elems, m, nloc, dev.stream()))
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        tdist.all_gather(parts, mine, group=self.group)
        joined = torch.cat([p[:, :c] for p, c in zip(parts, counts)], dim=1).contiguous()
        return dev.download_2d(joined.data_ptr(), total * dt.itemsize, m, total, dt.type)

    def barrier(self):
        tdist.barrier(group=self.group)


class HaloPlan:
    """Communication plan of a row-partitioned sparse operator.

    Input: this rank's rows of the full operator as CSR with GLOBAL column indices and
    the contiguous row partition.  Output: the local matrix with columns renumbered
    [owned 0..nloc) | halo nloc..nloc+nhalo), halo columns ordered by owner rank and
    then by global index; for every peer the local rows it needs from us (`send_idx`,
    concatenated in peer order, `send_counts`) and how many halo rows arrive from each
    peer (`recv_counts`).  Per SpMM only those boundary rows cross NVLink."""

    def __init__(self, ctx, indptr, indices, row0, nloc, n_global):
        self.ctx = ctx
        world, rank = ctx.world, ctx.rank
        starts = numpy.array([partition(n_global, world, r)[0] for r in range(world)] + [n_global], dtype=numpy.int64)
        assert starts[rank] == row0 and starts[rank + 1] - row0 == nloc
        indices = numpy.asarray(indices, dtype=numpy.int64)
        owned = (indices >= row0) & (indices < row0 + nloc)
        halo_cols = numpy.unique(indices[~owned])                    # sorted global ids => grouped by owner
        owner = numpy.searchsorted(starts, halo_cols, side='right') - 1
        self.recv_counts = [int(numpy.count_nonzero(owner == r)) for r in range(world)]
        self.nhalo = int(halo_cols.shape[0])
        self.nloc = int(nloc)
        local = numpy.empty_like(indices)
        local[owned] = indices[owned] - row0
        local[~owned] = nloc + numpy.searchsorted(halo_cols, indices[~owned])
        self.local_indices = local.astype(numpy.int32)
        self.indptr = numpy.asarray(indptr, dtype=numpy.int64)
        # tell every owner which of its rows we need
        need = [halo_cols[owner == r] - starts[r] for r in range(world)]
        wanted = self._exchange_lists(need)
        self.send_counts = [int(w.shape[0]) for w in wanted]
        self.send_idx = numpy.concatenate(wanted).astype(numpy.int64) if world > 0 else numpy.zeros(0, numpy.int64)

    def _exchange_lists(self, need):
        world = self.ctx.world
        gathered = [None] * world
        tdist.all_gather_object(gathered, [numpy.asarray(a, dtype=numpy.int64) for a in need], group=self.ctx.group)
        return [numpy.asarray(gathered[src][self.ctx.rank], dtype=numpy.int64) for src in range(world)]

    def exchange(self, send, recv, m):
        """send: torch tensor (sum(send_counts) * m,), recv: (nhalo * m,), both grouped by peer."""
        ctx = self.ctx
        ins = [c * m for c in self.send_counts]
        outs = [c * m for c in self.recv_counts]
        if ctx.on_device:
            tdist.all_to_all_single(recv, send, output_split_sizes=outs, input_split_sizes=ins, group=ctx.group)
            return
        ops, so, ro = [], 0, 0
        for peer in range(ctx.world):
            if ins[peer]:
                ops.append(tdist.P2POp(tdist.isend, send[so:so + ins[peer]], peer, group=ctx.group))
            if outs[peer]:
                ops.append(tdist.P2POp(tdist.irecv, recv[ro:ro + outs[peer]], peer, group=ctx.group))
            so += ins[peer]
            ro += outs[peer]
        if ops:
            for req in tdist.batch_isend_irecv(ops):
                req.wait()


def enable(group=None, shard_matrices=True):
    """Activate sharding for objects created from now on in this process."""
    global _current
    _current = ShardContext(group, shard_matrices)
    return _current


def disable():
    global _current
    _current = None


def current():
    return _current
