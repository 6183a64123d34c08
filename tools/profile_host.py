"""cProfile of one resident C2 solve on the GPU backend: where the host time goes."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from threadpoolctl import threadpool_limits
import raleigh_b200 as rb
from bench import generate_shard as generate_c2
rb.install()
from raleigh.interfaces.lra import LowerRankApproximation
from raleigh.algebra.dense_matrix import AMatrix
from raleigh.core.solver import Options
rows, cols, npc = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (12000, 39375, 1000)))
a = generate_c2(rows, cols, 2000, 'cuda').cpu().numpy()
matrix = AMatrix(a, arch='gpu!')
def solve():
    np.random.seed(1)
    lra = LowerRankApproximation(); lra.ortho = 1e-3
    lra.compute(matrix, opt=Options(), rank=npc, tol=0, norm='f', max_rank=-1, svtol=1e-3, shift=True, verb=0)
    return lra
with threadpool_limits(limits=1):
    solve()
    torch.cuda.synchronize(); t0 = time.time(); solve(); torch.cuda.synchronize(); print('solve s', time.time() - t0)
    pr = cProfile.Profile(); pr.enable(); solve(); pr.disable()
st = pstats.Stats(pr); st.sort_stats('tottime').print_stats(28)
st.sort_stats('cumulative').print_stats(30)
