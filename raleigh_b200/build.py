"""Build libraleigh_b200.so in-tree with nvcc for sm_100a.

    python raleigh_b200/build.py [--force] [--verbose]

The library is a plain C-ABI shared object (include/raleigh_b200.h); it is
compiled once per source change (content hash) and shipped to the GPU box with
the repo snapshot -- nothing is JIT-compiled at run time.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, 'libraleigh_b200.so')
STAMP = os.path.join(HERE, 'build', 'stamp.txt')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
    '--expt-relaxed-constexpr',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build libraleigh_b200.so')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    files.append(os.path.join(ROOT, 'include', 'raleigh_b200.h'))
    for f in files:
        h.update(f.encode())
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return False
    try:
        with open(STAMP) as fh:
            return fh.read().strip() == _digest()
    except OSError:
        return False


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library (objects in parallel)."""
    if not force and is_current():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for %s:\n%s\n' % (src, out))
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    with open(STAMP, 'w') as fh:
        fh.write(_digest())
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
