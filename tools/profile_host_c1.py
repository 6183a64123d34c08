"""cProfile of config 1 (32^3 Laplacian, 10 eigenpairs, tol 1e-6) on the device-resident driver: host time per call site."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import raleigh_b200 as rb
rb.install()
import raleigh.core.solver as rs
from raleigh.examples.laplace import lap3d
from raleigh_b200 import profile
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = lap3d(n1, n1, n1, 1.0, 1.0, 1.0)
n = L.shape[0]
A = rb.SparseSymmetricMatrix(L)
def solve():
    np.random.seed(1)
    opt = rs.Options(); opt.block_size = -1; opt.max_iter = 1000
    opt.convergence_criteria = rs.DefaultConvergenceCriteria()
    opt.convergence_criteria.set_error_tolerance('k eigenvector error', 1e-6)
    v = rb.Vectors(n, data_type=np.float64)
    s = rs.Solver(rs.Problem(v, A))
    s.solve(v, opt, which=(10, 0))
    return s
solve()
torch.cuda.synchronize(); t0 = time.time(); s = solve(); torch.cuda.synchronize(); dt = time.time() - t0
profile.reset(); profile.enable(True); solve(); torch.cuda.synchronize(); profile.enable(False)
rep = profile.report()
print('solve s %.4f iterations %d device ms %.1f launches/calls %s' % (dt, s.iteration, sum(v['ms'] for v in rep.values()), {k: v['count'] for k, v in rep.items()}))
pr = cProfile.Profile(); pr.enable(); solve(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(22)
