"""Build libminer_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

    python -m miner_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB = os.path.join(PKG, 'libminer_b200.so')
SOURCES = ['api.cu', 'gather.cu', 'sgemm.cu', 'poly.cu', 'target.cu', 'bias.cu', 'metrics.cu', 'auc.cu', 'loss.cu', 'train.cu', 'tc/tc_gemm.cu', 'tc/tc_gemm_tn.cu', 'tc/tmap.cu', 'tc/hist_kernel2.cu', 'tc/cand_kernel.cu', 'tc/table_project.cu', 'tc/tscore_kernel.cu', 'tc/tscore_x_kernel.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
              '-Xptxas', '-v', '--expt-relaxed-constexpr']


def _nvcc() -> str:
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError('nvcc not found')


def _deps_mtime() -> float:
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    m = max(m, os.path.getmtime(os.path.join(os.path.dirname(PKG), 'include', 'miner_b200.h')))
    return m


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = None) -> str:
    """`extra_flags` / `out`: instrumented variants (e.g. -DMINER_HIST_PROF -> libminer_b200_prof.so, objects in build_<name>/)."""
    global OBJ, LIB
    if out:
        LIB = os.path.join(PKG, out)
        OBJ = os.path.join(PKG, 'build_' + os.path.splitext(out)[0])
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace('/', '_') + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = r.stderr
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    with open(os.path.join(OBJ, 'ptxas.log'), 'w') as f:
        for src in SOURCES:
            f.write(f'==== {src}\n{logs[src]}\n')
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    if verbose:
        for src in SOURCES:
            print(f'==== {src}\n{logs[src]}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, extra_flags=[a for a in sys.argv[1:] if a.startswith('-D')],
                out=next((a.split('=', 1)[1] for a in sys.argv[1:] if a.startswith('--out=')), None)))
