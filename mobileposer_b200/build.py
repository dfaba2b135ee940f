"""Build libmobileposer_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m mobileposer_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libmobileposer_b200.so')
SOURCES = ['gemm.cu', 'gemm_tc.cu', 'gemm_f16.cu', 'lstm_rec.cu', 'lstm_rec_tc.cu', 'lstm_rec_f16.cu', 'lstm_rec_f16w.cu', 'kinematics.cu', 'physics.cu', 'inputs.cu', 'evaluate.cu', 'train.cu', 'api.cu']
HEADERS = [os.path.join(CSRC, 'mp_common.cuh'), os.path.join(CSRC, 'mp_constants.cuh'),
           os.path.join(os.path.dirname(PKG), 'include', 'mobileposer_b200.h')]
OBJ_DIR = os.path.join(LIB_DIR, 'obj')
NVCC_FLAGS = ['-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libmobileposer_b200.so')
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(src: str, force: bool, verbose: bool):
    """One translation unit -> lib/obj/<name>.o (skipped when newer than the source and every header)."""
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + '.o')
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj):
        t = os.path.getmtime(obj)
        if all(os.path.getmtime(d) <= t for d in [path] + HEADERS):
            return obj, ''
    cmd = [_nvcc(), *NVCC_FLAGS, '-c', '-o', obj, path]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n' + res.stdout + res.stderr)
    return obj, res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (one nvcc per file, in parallel) and link one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(lambda s: _compile_one(s, force, verbose), SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    cmd = [_nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH, *[o for o, _ in results]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc link failed:\n' + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
