"""CPU oracle for the RALEIGH abstract-vectors algebra (TEST INFRASTRUCTURE ONLY).

This package is a NumPy/SciPy restatement of the reference's CPU algebra
(raleigh/algebra/dense_numpy.py, dense_ndarray.py, sparse_mkl.py + the
MKL csrmm semantics of mkl_wrap.py).  It is the checker for the CUDA path,
never the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  raleigh_b200 itself never
imports anything from here and raises if its CUDA library is missing.

Parity pin: tests/golden/*.npz were produced by importing the *reference's own*
dense_numpy.Vectors / Matrix and core solver in the build container
(tests/golden/make_golden.py, committed); tests/test_oracle.py checks every
oracle function against those vectors, and against the reference's two
known-answer doctests (core_solver.py:65-71, the tests_algebra.py identities).
The sparse operator has no reference-side test at all (MKL is
absent and tests_algebra.py never touches sparse_mkl), so the SpMM oracle is
pinned only on the definition  A_sym = U + triu(U,1)^H  (sparse_mkl.py:18-31,
mkl_wrap.py:264-276) and on the analytic spectrum of the 7-point Laplacian.
"""
from .algebra_np import *  # noqa: F401,F403
from .host_backend import Vectors, Matrix, SparseSymmetricMatrix, Operator, Jacobi  # noqa: F401
