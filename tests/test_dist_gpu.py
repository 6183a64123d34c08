"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): launches
tests/dist_check.py under torchrun with one rank per GPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_row_sharded_path_on_two_gpus():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tests', 'dist_check.py')]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and 'DIST_CHECK_OK' in r.stdout, r.stdout[-3000:]
