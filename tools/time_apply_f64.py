"""fp64 Matrix.apply: FP64-tensor-pipe kernel (gemm_dmma.cu) against the FMA-pipe kernel (gemm_simt.cu, knob 16 = -1)
and torch.matmul (cuBLAS DGEMM, the reference's route) at the config-2 shape in double precision.
    python tools/time_apply_f64.py [--rows 12000 --cols 39375]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raleigh_b200._lib import lib, check
from raleigh_b200 import device as dev

ap = argparse.ArgumentParser()
ap.add_argument('--rows', type=int, default=12000)
ap.add_argument('--cols', type=int, default=39375)
args = ap.parse_args()
M, N = args.rows, args.cols
lda = (N + 15) // 16 * 16
A = torch.randn(M, lda, dtype=torch.float64, device='cuda')
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for k in (16, 32, 128):
    ldn, ldm = (N + 15) // 16 * 16, (M + 15) // 16 * 16
    Xn = torch.randn(k, ldn, dtype=torch.float64, device='cuda')
    Ym = torch.zeros(k, ldm, dtype=torch.float64, device='cuda')
    Yn = torch.zeros(k, ldn, dtype=torch.float64, device='cuda')
    out = {'A': '%dx%d fp64' % (M, N), 'k': k}
    st = dev.stream()
    for knob, name in ((0, 'dmma'), (-1, 'fma')):
        lib.rl_debug_set_knob(16, knob)
        ms = timeit(lambda: check(lib.rl_dense_apply(1, A.data_ptr(), lda, M, N, Xn.data_ptr(), ldn, Ym.data_ptr(), ldm, k, 0, 1.0, 0.0, st)))
        out['apply_%s_ms' % name] = round(ms, 3); out['apply_%s_TFLOPs' % name] = round(2.0 * M * N * k / ms / 1e9, 2)
        ref = Xn[:, :N] @ A[:, :N].T
        out['apply_%s_err' % name] = float((Ym[:, :M] - ref).abs().max() / ref.abs().max())
        ms = timeit(lambda: check(lib.rl_dense_apply(1, A.data_ptr(), lda, M, N, Ym.data_ptr(), ldm, Yn.data_ptr(), ldn, k, 1, 1.0, 0.0, st)))
        out['apply_t_%s_ms' % name] = round(ms, 3); out['apply_t_%s_TFLOPs' % name] = round(2.0 * M * N * k / ms / 1e9, 2)
        ref = Ym[:, :M] @ A[:, :N]
        out['apply_t_%s_err' % name] = float((Yn[:, :N] - ref).abs().max() / ref.abs().max())
    lib.rl_debug_set_knob(16, 0)
    ms = timeit(lambda: torch.matmul(Xn[:, :N], A[:, :N].T))
    out['cublas_dgemm_ms'] = round(ms, 3); out['cublas_dgemm_TFLOPs'] = round(2.0 * M * N * k / ms / 1e9, 2)
    print(json.dumps(out), flush=True)
