"""Build libblp.so (the sm_100a CUDA library behind include/blp.h) in-tree with nvcc.

The shared object lands next to its sources (simple_mip_solver_b200/csrc/libblp.so); it is
git-ignored but travels to the GPU box with the repo snapshot. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / 'csrc'
LIB = CSRC / 'libblp.so'
SOURCES = [CSRC / 'blp.cu', CSRC / 'blp_mps.cpp']
DEPS = [CSRC / 'blp_kernels.cuh', CSRC / 'blp_simplex.cuh', CSRC / 'blp_prep.hpp', CSRC.parent.parent / 'include' / 'blp.h']


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: libblp.so cannot be built')
    return exe


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + DEPS)


def build_extension(force: bool = False, verbose: bool = False, tuning: bool = False, out: Path = None) -> Path:
    """``tuning=True`` adds -DBLP_TUNING: the library then reads the BLP_* sweep variables (tools/).
    ``out``: build to another path (e.g. a tuning library next to the production one; select it with
    the BLP_LIB environment variable)."""
    if out is None and not force and not is_stale():
        return LIB
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
           '-shared', '-Xcompiler', '-fPIC', '-o', str(out or LIB)] + [str(s) for s in SOURCES]
    if tuning:
        cmd.insert(1, '-DBLP_TUNING')
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return Path(out) if out is not None else LIB


if __name__ == '__main__':
    import sys
    print(build_extension(force=True, verbose=True, tuning='--tuning' in sys.argv))
