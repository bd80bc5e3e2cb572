"""In-tree build of libvosprop.so for sm_100a (nvcc cross-compiles without a GPU).

The templated kernels are instantiated in several translation units (one per class capacity and family), compiled in
parallel into csrc/build/*.o and linked into csrc/libvosprop.so.  Environment:
  NVCC           path of nvcc
  VOS_NVCC_DEFS  extra flags for every translation unit (e.g. -DVOS_KERNEL_DEBUG for tools/epilogue_ablation.py)
  VOS_LIB_NAME   output name instead of libvosprop.so (variant builds side by side; objects go to build/<name>/)
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent.parent / 'csrc'
LIB = CSRC / 'libvosprop.so'
CLASS_CAPS = (2, 3, 4, 6, 8, 11, 14)
HEADERS = ['kernels.cuh', 'side_kernels.cuh', 'affinity_idx.cuh', 'affinity_topk.cuh', 'affinity_prob.cuh', 'topk_params.h', 'launch.h', 'ptx.cuh',
           'decompose.h', '../../include/vos_prop.h', '../../include/vos_jpeg.h']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def units():
    """(object name, source, extra defines)"""
    out = [('vos_prop', 'vos_prop.cu', []), ('inst_dispatch', 'inst_dispatch.cu', []), ('inst_topk', 'inst_topk.cu', []), ('jpeg', 'jpeg.cu', [])]
    for d in CLASS_CAPS + (24,):
        out.append((f'inst_idx_{d}', 'inst_idx.cu', [f'-DVOS_INST_D={d}']))
    for d in CLASS_CAPS:
        out.append((f'inst_dense_{d}', 'inst_dense.cu', [f'-DVOS_INST_D={d}']))
    return out


def find_nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and Path(cand).is_file():
            return cand
    raise RuntimeError('nvcc not found (set NVCC=...)')


def lib_path() -> Path:
    name = os.environ.get('VOS_LIB_NAME')
    return CSRC / name if name else LIB


def _newest_header() -> float:
    return max((CSRC / f).resolve().stat().st_mtime for f in HEADERS)


def is_stale() -> bool:
    lib = lib_path()
    if not lib.is_file():
        return True
    t = lib.stat().st_mtime
    srcs = {src for _, src, _ in units()}
    return _newest_header() > t or any((CSRC / s).stat().st_mtime > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> Path:
    lib = lib_path()
    if not force and not is_stale():
        return lib
    nvcc = find_nvcc()
    extra = os.environ.get('VOS_NVCC_DEFS', '').split()
    objdir = CSRC / 'build' / (lib.stem + ('_' + '_'.join(extra).replace('-', '').replace('=', '') if extra else ''))
    objdir.mkdir(parents=True, exist_ok=True)
    hdr_t = _newest_header()
    logs = []

    def compile_one(unit):
        name, src, defs = unit
        obj = objdir / f'{name}.o'
        if not force and obj.is_file() and obj.stat().st_mtime > max(hdr_t, (CSRC / src).stat().st_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *extra, *defs, '-c', '-o', str(obj), src]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src} {defs}:\n' + res.stdout + res.stderr)
        logs.append(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, units()))
    cmd = [nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-cudart', 'static', '-o', str(lib), *map(str, objs)]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    if verbose:
        print('\n'.join(logs))
    return lib


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
