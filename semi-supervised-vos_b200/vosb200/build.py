"""In-tree build of libvosprop.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent.parent / 'csrc'
LIB = CSRC / 'libvosprop.so'
SOURCES = ['vos_prop.cu']
HEADERS = ['kernels.cuh', 'affinity_idx.cuh', 'affinity_topk.cuh', 'ptx.cuh', 'decompose.h', '../../include/vos_prop.h']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static']


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).is_file():
            return cand
    raise RuntimeError('nvcc not found (set NVCC=...)')


def is_stale() -> bool:
    if not LIB.is_file():
        return True
    t = LIB.stat().st_mtime
    return any((CSRC / f).resolve().stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *os.environ.get('VOS_NVCC_DEFS', '').split(), '-o', str(LIB), *SOURCES]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
