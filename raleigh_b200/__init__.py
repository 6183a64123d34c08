"""raleigh_b200: B200-native (sm_100a) backend for RALEIGH's abstract-vectors algebra.

    import raleigh_b200
    raleigh_b200.install()          # alias into raleigh.algebra.{dense_cublas,cuda_wrap,dense_cblas,sparse_mkl}

Public classes mirror raleigh/algebra/dense_cublas.py and sparse_mkl.py.
Importing the package loads libraleigh_b200.so and fails loudly if it is missing.
"""
from ._lib import lib, LIB_PATH  # noqa: F401  (raises if the CUDA library is absent)
from .vectors import Vectors, Matrix  # noqa: F401
from .sparse import SparseSymmetricMatrix, Operator, DiagonalPreconditioner  # noqa: F401
from .compat import install, find_reference, use_device_solver  # noqa: F401
from .device import synchronize  # noqa: F401

__version__ = '0.1.0'
