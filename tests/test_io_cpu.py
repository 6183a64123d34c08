"""Input helpers (raleigh_b200/io.py): Matrix Market row slabs against scipy.io.mmread, .npy memory maps."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _io():
    # load the module by path: importing the package needs the CUDA library
    spec = importlib.util.spec_from_file_location('_rl_io', os.path.join(ROOT, 'raleigh_b200', 'io.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('symmetric', [True, False])
def test_matrix_market_row_slabs(tmp_path, symmetric):
    io = _io()
    rng = np.random.RandomState(3)
    n = 57
    A = sp.random(n, n, density=0.1, random_state=rng, format='coo')
    if symmetric:
        A = (A + A.T + sp.diags(rng.rand(n) + 1.0)).tocoo()
    path = str(tmp_path / ('s.mtx' if symmetric else 'g.mtx'))
    scipy.io.mmwrite(path, A, symmetry='symmetric' if symmetric else 'general')
    full = scipy.io.mmread(path).tocsr()
    whole = io.read_matrix_market(path)
    assert whole.shape == full.shape and abs(whole - full).max() == 0.0
    # slabs of a 3-way partition, parsed in tiny blocks of lines
    starts = [0, 19, 38, 57]
    for p in range(3):
        slab = io.read_matrix_market(path, starts[p], starts[p + 1] - starts[p], block_lines=7)
        assert slab.shape == (starts[p + 1] - starts[p], n)
        assert abs(slab - full[starts[p]:starts[p + 1]]).max() == 0.0
        assert slab.has_sorted_indices
    empty = io.read_matrix_market(path, 5, 0)
    assert empty.shape == (0, n) and empty.nnz == 0
    with pytest.raises(ValueError):
        io.read_matrix_market(path, 50, 20)


def test_matrix_market_rejects_other_formats(tmp_path):
    io = _io()
    path = str(tmp_path / 'a.mtx')
    scipy.io.mmwrite(path, np.arange(6.0).reshape(2, 3))          # array format
    with pytest.raises(ValueError):
        io.read_matrix_market(path)


def test_open_npy(tmp_path):
    io = _io()
    a = np.random.RandomState(0).rand(40, 7).astype(np.float32)
    path = str(tmp_path / 'a.npy')
    np.save(path, a)
    m = io.open_npy(path)
    assert isinstance(m, np.memmap) and m.shape == a.shape and np.array_equal(m[10:20], a[10:20])
    chunk = m[10:20]
    assert chunk.flags['C_CONTIGUOUS'] and isinstance(chunk.base, np.ndarray)      # what the chunk prefetch relies on
    assert chunk.ctypes.data - chunk.base.ctypes.data == 10 * 7 * 4
    np.save(path, np.zeros(5))
    with pytest.raises(ValueError):
        io.open_npy(path)
