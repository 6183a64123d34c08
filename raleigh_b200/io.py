"""Input helpers for the two file formats the reference's examples read:

* `.npy` data matrices (examples/pca/incremental_pca.py:43: `numpy.load(path, mmap_mode='r')`): `open_npy` returns
  the memory map; `pca(data, batch_size=..., arch='gpu!')` then streams it chunk by chunk -- the Matrix constructor
  uploads each chunk through the staged copy path (page faults included) and, inside lra.icompute, prefetches the
  next one in the background (vectors._ChunkPrefetch works on slices of a memory map like on slices of an array).
  `npy_rows(path, rank, world)` is the slab of a sample-partitioned run.

* Matrix Market coordinate files (examples/sparse_evp.py:68-70: `mmread(matrix).tocsr()`): `read_matrix_market`
  parses the file in blocks of lines and keeps only the rows of one slab [row0, row0 + rows) -- for a symmetric
  file both the stored entry and its mirror image are considered -- so that every process of a row-partitioned run
  builds its own part of the operator, `SparseSymmetricMatrix(slab, local_rows=(row0, n))`, without any process
  ever holding the whole matrix.
"""
import numpy
import scipy.sparse as sp


def open_npy(path):
    """Memory map of a 2-D `.npy` file (C order), read-only."""
    a = numpy.load(path, mmap_mode='r')
    if a.ndim != 2:
        raise ValueError('%s does not hold a 2-D array' % path)
    return a


def npy_rows(path, rank, world):
    """Rows of the `.npy` matrix owned by process `rank` of `world` (contiguous block partition, dist.partition)."""
    from .dist import partition
    a = open_npy(path)
    row0, rows = partition(a.shape[0], world, rank)
    return a[row0:row0 + rows], row0


def read_matrix_market(path, row0=0, rows=None, block_lines=1 << 20, dtype=numpy.float64):
    """Rows [row0, row0 + rows) of a real Matrix Market coordinate matrix as CSR with GLOBAL column indices,
    shape (rows, ncols).  `symmetric` files are expanded (entry (i, j) also gives (j, i)); `general` files are
    taken as they are.  Duplicate entries are summed, like scipy.io.mmread does."""
    with open(path, 'r') as f:
        header = f.readline().split()
        if len(header) < 5 or header[0] != '%%MatrixMarket' or header[1].lower() != 'matrix':
            raise ValueError('%s is not a Matrix Market matrix file' % path)
        fmt, field, symmetry = header[2].lower(), header[3].lower(), header[4].lower()
        if fmt != 'coordinate' or field not in ('real', 'integer', 'double'):
            raise ValueError('only real coordinate Matrix Market files are supported (%s %s)' % (fmt, field))
        if symmetry not in ('general', 'symmetric'):
            raise ValueError('unsupported symmetry %s' % symmetry)
        line = f.readline()
        while line.startswith('%') or not line.strip():
            line = f.readline()
        nrows, ncols, nnz = (int(t) for t in line.split()[:3])
        if rows is None:
            rows = nrows - row0
        if row0 < 0 or rows < 0 or row0 + rows > nrows:
            raise ValueError('rows [%d, %d) outside the matrix (%d rows)' % (row0, row0 + rows, nrows))
        ri, ci, vv = [], [], []
        left = nnz
        while left > 0:
            take = min(left, block_lines)
            blk = numpy.loadtxt(f, dtype=numpy.float64, max_rows=take, ndmin=2)
            if blk.shape[0] == 0:
                break
            left -= blk.shape[0]
            i = blk[:, 0].astype(numpy.int64) - 1
            j = blk[:, 1].astype(numpy.int64) - 1
            v = blk[:, 2]
            keep = (i >= row0) & (i < row0 + rows)
            ri.append(i[keep] - row0); ci.append(j[keep]); vv.append(v[keep])
            if symmetry == 'symmetric':
                keep = (j >= row0) & (j < row0 + rows) & (i != j)
                ri.append(j[keep] - row0); ci.append(i[keep]); vv.append(v[keep])
        if left != 0:
            raise ValueError('%s ends before its %d entries' % (path, nnz))
    if ri:
        ri, ci, vv = numpy.concatenate(ri), numpy.concatenate(ci), numpy.concatenate(vv)
    else:
        ri = ci = numpy.zeros(0, numpy.int64); vv = numpy.zeros(0)
    slab = sp.coo_matrix((vv.astype(dtype), (ri, ci)), shape=(rows, ncols)).tocsr()
    slab.sum_duplicates()
    slab.sort_indices()
    return slab
