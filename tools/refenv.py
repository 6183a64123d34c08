"""Locating and shimming the UNMODIFIED reference package for harnesses that must not load the
product: bench.py's `--impl reference` arm and its cpu_baseline leg import only this module (plus
NumPy / SciPy), never raleigh_b200 -- so no CUDA library, no kernels and no GPU are involved in the
reference's numbers.

The shims are version fixes the reference needs on ANY backend with current SciPy / NumPy / Python 3
and touch no algebra: `scipy.linalg.eigh(..., turbo=...)` (solver.py:578,822,899,1470) lost its keyword
in SciPy 1.14; `nv = min(32, nsv/2)` (partial_svd.py:201) is Python-2 integer division.
"""
import builtins
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    """Directory containing the reference's `raleigh` package, or None.
    Order: already importable, $RALEIGH_REFERENCE, <repo>/baseline/_ref."""
    try:
        spec = importlib.util.find_spec('raleigh')
        if spec is not None and spec.submodule_search_locations:
            return os.path.dirname(list(spec.submodule_search_locations)[0])
    except (ImportError, ValueError):
        pass
    for cand in (os.environ.get('RALEIGH_REFERENCE'), os.path.join(ROOT, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'raleigh')):
            return cand
    return None


class _SlaProxy:
    def __init__(self, sla):
        self._sla = sla

    def __getattr__(self, name):
        return getattr(self._sla, name)

    def eigh(self, *args, turbo=None, **kwargs):
        return self._sla.eigh(*args, **kwargs)


class _NumpyProxy:
    def __init__(self, np):
        self._np = np

    def __getattr__(self, name):
        return getattr(self._np, name)

    def eye(self, n, *args, **kwargs):
        return self._np.eye(int(n), *args, **kwargs)


def _int_min(*args, **kwargs):
    r = builtins.min(*args, **kwargs)
    return int(r) if isinstance(r, float) else r


def load_reference():
    """Put the reference on sys.path and apply the version shims.  Returns its path or None."""
    path = find_reference()
    if path is None:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    import numpy
    import scipy.linalg as sla
    import raleigh.core.solver as rsolver
    if type(rsolver.sla).__name__ != '_SlaProxy':
        rsolver.sla = _SlaProxy(sla)
    import raleigh.interfaces.partial_svd as psvd
    if type(psvd.numpy).__name__ != '_NumpyProxy':
        psvd.numpy = _NumpyProxy(numpy)
    psvd.min = _int_min
    return path
