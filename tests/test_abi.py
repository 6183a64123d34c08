"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/raleigh_b200.h declares, and the ctypes table mirrors it."""
import ctypes
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'raleigh_b200.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rl_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from raleigh_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) > 30
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from raleigh_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()


def test_no_compute_without_gpu_but_errors_decode():
    from raleigh_b200 import _lib
    assert _lib.lib.rl_version() >= 100
    assert b'dtype' in _lib.lib.rl_error_string(-1)
    assert _lib.lib.rl_gram_ws_bytes(1, 0, 0, 0) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'raleigh_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_missing_device_fails_loudly():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip('box has a GPU')
    import raleigh_b200
    with pytest.raises(RuntimeError):
        raleigh_b200.Vectors(8, 2)
