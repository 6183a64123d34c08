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

    def barrier(self):
        tdist.barrier(group=self.group)


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
